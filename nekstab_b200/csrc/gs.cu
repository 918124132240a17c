// gs.cu -- gather-scatter (direct-stiffness summation) over a precomputed index map.
//
// Replaces Nek5000's dssum/opdssum -> gslib gs_op(add) [UPSTREAM dssum.f, gslib gs.c], reached from every
// Helmholtz/pressure iteration inside `nek_advance` (core/matvec.f:222) and from add_noise (core/utils.f:391-403).
// Setup sorts the local dofs by global node number (the `glo_num` Nek's setvert produces) into CSR segments
// = equivalence classes of coincident GLL nodes.  The kernel is an atomics-free segmented
// sum: one thread owns one segment, adds its copies in a fixed order and writes the sum back to every copy,
// for up to 3 fields per launch (opdssum).  Across GPUs the partial sums of interface nodes are packed per
// neighbour, exchanged with grouped ncclSend/ncclRecv over NVLink, and added in ascending rank order (own
// partial at its rank position) so every rank computes bit-identical sums.
#include <algorithm>
#include <numeric>

#include "nsb_internal.h"

static int* d_send_base = nullptr;   // [nshared]
static int* d_send_cnt = nullptr;    // [nshared]
static int* d_rseg_cnt = nullptr;    // parallel to rseg_pos

template <int NF>
__global__ void k_gs_pack(int nshared, const int* __restrict__ send_seg, const int* __restrict__ send_base,
                          const int* __restrict__ send_cnt, const int* __restrict__ seg_off,
                          const int* __restrict__ seg_idx, const double* __restrict__ u, long long stride,
                          double* __restrict__ sendbuf, const CGState* skip) {
  if (skip && skip->done) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nshared) return;
  int seg = send_seg[s];
  int a = seg_off[seg], b = seg_off[seg + 1];
  double acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = 0.0;
  for (int j = a; j < b; ++j) {
    int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] += u[(long long)f * stride + idx];
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) sendbuf[send_base[s] + f * send_cnt[s]] = acc[f];
}

template <int NF, bool HALO>
__global__ void k_gs_sum(int nseg, const int* __restrict__ seg_off, const int* __restrict__ seg_idx,
                         const int* __restrict__ rseg_off, const int* __restrict__ rseg_pos,
                         const int* __restrict__ rseg_cnt, const int* __restrict__ rseg_nbefore,
                         const double* __restrict__ recvbuf, double* __restrict__ u, long long stride,
                         const CGState* skip) {
  if (skip && skip->done) return;
  int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= nseg) return;
  int a = seg_off[seg], b = seg_off[seg + 1];
  double acc[NF], loc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) { acc[f] = 0.0; loc[f] = 0.0; }
  for (int j = a; j < b; ++j) {
    int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) loc[f] += u[(long long)f * stride + idx];
  }
  if (HALO) {
    int ra = rseg_off[seg], rb = rseg_off[seg + 1], nb = rseg_nbefore[seg];
    bool any_before = nb > 0;
    for (int j = ra; j < ra + nb; ++j) {
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] += recvbuf[rseg_pos[j] + f * rseg_cnt[j]];
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = any_before ? acc[f] + loc[f] : loc[f];
    for (int j = ra + nb; j < rb; ++j) {
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] += recvbuf[rseg_pos[j] + f * rseg_cnt[j]];
    }
  } else {
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = loc[f];
  }
  for (int j = a; j < b; ++j) {
    int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) u[(long long)f * stride + idx] = acc[f];
  }
}

template <class T>
static int upload(T** dptr, const std::vector<T>& h) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  NSB_CUDA(cudaMalloc((void**)dptr, bytes));
  if (!h.empty()) NSB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int gs_free(Ctx* c) {
  GSMap& m = c->gs;
  cudaFree(m.seg_off); cudaFree(m.seg_idx); cudaFree(m.send_seg); cudaFree(m.rseg_off); cudaFree(m.rseg_pos);
  cudaFree(m.rseg_nbefore); cudaFree(m.sendbuf); cudaFree(m.recvbuf);
  cudaFree(d_send_base); cudaFree(d_send_cnt); cudaFree(d_rseg_cnt);
  cudaFree(m.surf_pts); cudaFree(m.nb_off); cudaFree(m.nb_idx);
  d_send_base = d_send_cnt = d_rseg_cnt = nullptr;
  m = GSMap();
  return 0;
}

int gs_setup(Ctx* c, const long long* glo) {
  GSMap& m = c->gs;
  const long long n = c->n;
  if (n >= (1LL << 31)) { nsb_set_error("gs_setup: more than 2^31 local dofs"); return 1; }
  // ---- sort local dofs by global id (stable in local index => fixed summation order)
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return glo[a] != glo[b] ? glo[a] < glo[b] : a < b; });
  // unique ids with their runs
  std::vector<long long> uid;
  std::vector<int> ustart;
  for (long long i = 0; i < n; ++i)
    if (i == 0 || glo[order[i]] != glo[order[i - 1]]) { uid.push_back(glo[order[i]]); ustart.push_back((int)i); }
  ustart.push_back((int)n);
  const int nu = (int)uid.size();

  // ---- which unique ids are shared with other ranks?  (candidates: nodes on element surfaces)
  std::vector<std::vector<int>> shared_u(c->nranks);   // per neighbour rank: indices into uid, ascending id
  if (c->nranks > 1) {
    const int N = c->lx1, D = c->ldim, np = c->np1;
    std::vector<long long> cand;
    for (int k = 0; k < nu; ++k) {
      int p = order[ustart[k]] % np;
      int i = p % N, j = (p / N) % N, kk = (D == 3) ? p / (N * N) : 1;
      bool surf = (i == 0 || i == N - 1 || j == 0 || j == N - 1 || (D == 3 && (kk == 0 || kk == N - 1)));
      if (surf) cand.push_back(uid[k]);
    }
    // allgather counts then ids (padded) through NCCL
    long long mycnt = (long long)cand.size();
    long long* d_cnt = nullptr;
    NSB_CUDA(cudaMalloc(&d_cnt, sizeof(long long) * (c->nranks + 1)));
    NSB_CUDA(cudaMemcpy(d_cnt + c->nranks, &mycnt, sizeof(long long), cudaMemcpyHostToDevice));
    NSB_NCCL(ncclAllGather(d_cnt + c->nranks, d_cnt, 1, ncclInt64, c->comm, c->stream));
    NSB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<long long> cnts(c->nranks);
    NSB_CUDA(cudaMemcpy(cnts.data(), d_cnt, sizeof(long long) * c->nranks, cudaMemcpyDeviceToHost));
    cudaFree(d_cnt);
    long long mx = *std::max_element(cnts.begin(), cnts.end());
    if (mx == 0) mx = 1;
    long long *d_my = nullptr, *d_all = nullptr;
    NSB_CUDA(cudaMalloc(&d_my, sizeof(long long) * mx));
    NSB_CUDA(cudaMalloc(&d_all, sizeof(long long) * mx * c->nranks));
    NSB_CUDA(cudaMemset(d_my, 0xff, sizeof(long long) * mx));
    NSB_CUDA(cudaMemcpy(d_my, cand.data(), sizeof(long long) * cand.size(), cudaMemcpyHostToDevice));
    NSB_NCCL(ncclAllGather(d_my, d_all, mx, ncclInt64, c->comm, c->stream));
    NSB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<long long> all(mx * c->nranks);
    NSB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(long long) * all.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_my); cudaFree(d_all);
    for (int r = 0; r < c->nranks; ++r) {
      if (r == c->rank) continue;
      const long long* other = all.data() + (size_t)r * mx;
      long long no = cnts[r];
      // both lists ascending: two-pointer intersection against uid (ascending)
      long long a = 0;
      int k = 0;
      while (a < no && k < nu) {
        if (other[a] < uid[k]) ++a;
        else if (other[a] > uid[k]) ++k;
        else { shared_u[r].push_back(k); ++a; ++k; }
      }
    }
  }

  // ---- segments: unique ids with local multiplicity > 1 or shared remotely
  std::vector<char> is_shared(nu, 0);
  for (int r = 0; r < c->nranks; ++r)
    for (int k : shared_u[r]) is_shared[k] = 1;
  std::vector<int> seg_of_u(nu, -1);
  std::vector<int> seg_off(1, 0), seg_idx;
  {
    // order the segments by their smallest local dof: neighbouring threads then touch neighbouring memory
    // (r1a ncu: with global-id order the 3-field dssum moved 2x the algorithmic bytes at 20 % of the HBM roofline)
    std::vector<int> segk;
    for (int k = 0; k < nu; ++k)
      if (ustart[k + 1] - ustart[k] > 1 || is_shared[k]) segk.push_back(k);
    std::sort(segk.begin(), segk.end(), [&](int a, int b) { return order[ustart[a]] < order[ustart[b]]; });
    for (int k : segk) {
      seg_of_u[k] = (int)seg_off.size() - 1;
      for (int j = ustart[k]; j < ustart[k + 1]; ++j) seg_idx.push_back(order[j]);
      seg_off.push_back((int)seg_idx.size());
    }
  }
  m.nseg = (int)seg_off.size() - 1;

  // ---- halo lists
  std::vector<int> send_seg, send_base, send_cnt;
  std::vector<std::vector<std::pair<int, int>>> rlist(m.nseg);   // per segment: (pos, cnt) in ascending rank
  std::vector<int> nbefore(m.nseg, 0);
  m.nbr_rank.clear(); m.nbr_off.clear();
  int off = 0;
  for (int r = 0; r < c->nranks; ++r) {
    if (shared_u[r].empty()) continue;
    int cnt = (int)shared_u[r].size();
    m.nbr_rank.push_back(r);
    m.nbr_off.push_back(off);
    for (int j = 0; j < cnt; ++j) {
      int seg = seg_of_u[shared_u[r][j]];
      send_seg.push_back(seg);
      send_base.push_back(3 * off + j);
      send_cnt.push_back(cnt);
      rlist[seg].push_back({3 * off + j, cnt});
      if (r < c->rank) nbefore[seg]++;
    }
    off += cnt;
  }
  m.nbr_off.push_back(off);
  m.nnbr = (int)m.nbr_rank.size();
  m.nshared = off;
  std::vector<int> rseg_off(1, 0), rseg_pos, rseg_cnt;
  for (int s = 0; s < m.nseg; ++s) {
    for (auto& pr : rlist[s]) { rseg_pos.push_back(pr.first); rseg_cnt.push_back(pr.second); }
    rseg_off.push_back((int)rseg_pos.size());
  }
  // ---- per-element gather table (fused direct-stiffness sum; single rank only: no halo entries)
  if (c->nranks == 1) {
    const int N = c->lx1, D = c->ldim, np = c->np1;
    std::vector<int> surf;
    for (int p = 0; p < np; ++p) {
      int i = p % N, j = (p / N) % N, kk = (D == 3) ? p / (N * N) : 1;
      if (i == 0 || i == N - 1 || j == 0 || j == N - 1 || (D == 3 && (kk == 0 || kk == N - 1))) surf.push_back(p);
    }
    const int ns = (int)surf.size();
    std::vector<int> uof(n);                       // dof -> index of its unique id
    for (int k = 0; k < nu; ++k)
      for (int j = ustart[k]; j < ustart[k + 1]; ++j) uof[order[j]] = k;
    std::vector<int> nb_off((size_t)c->nel * (ns + 1)), nb_idx;
    nb_idx.reserve((size_t)c->nel * ns * 2);
    bool overflow = false;
    for (int e = 0; e < c->nel; ++e) {
      for (int s = 0; s < ns; ++s) {
        nb_off[(size_t)e * (ns + 1) + s] = (int)nb_idx.size();
        const int k = uof[(size_t)e * np + surf[s]];
        for (int j = ustart[k]; j < ustart[k + 1]; ++j) nb_idx.push_back(order[j]);   // ascending dof order
        if (nb_idx.size() > (size_t)2000000000) overflow = true;
      }
      nb_off[(size_t)e * (ns + 1) + ns] = (int)nb_idx.size();
    }
    if (!overflow) {
      m.ns = ns;
      NSB_TRY(upload(&m.surf_pts, surf));
      NSB_TRY(upload(&m.nb_off, nb_off));
      NSB_TRY(upload(&m.nb_idx, nb_idx));
    }
  }
  NSB_TRY(upload(&m.seg_off, seg_off));
  NSB_TRY(upload(&m.seg_idx, seg_idx));
  NSB_TRY(upload(&m.send_seg, send_seg));
  NSB_TRY(upload(&d_send_base, send_base));
  NSB_TRY(upload(&d_send_cnt, send_cnt));
  NSB_TRY(upload(&m.rseg_off, rseg_off));
  NSB_TRY(upload(&m.rseg_pos, rseg_pos));
  NSB_TRY(upload(&d_rseg_cnt, rseg_cnt));
  NSB_TRY(upload(&m.rseg_nbefore, nbefore));
  size_t hb = std::max<size_t>((size_t)3 * m.nshared, 1) * sizeof(double);
  NSB_CUDA(cudaMalloc(&m.sendbuf, hb));
  NSB_CUDA(cudaMalloc(&m.recvbuf, hb));
  NSB_CUDA(cudaMemset(m.sendbuf, 0, hb));
  NSB_CUDA(cudaMemset(m.recvbuf, 0, hb));
  return 0;
}

template <int NF>
static int dssum_nf(Ctx* c, double* u, long long stride, const CGState* skip) {
  GSMap& m = c->gs;
  const int T = 128;
  if (m.nshared > 0) {
    k_gs_pack<NF><<<(m.nshared + T - 1) / T, T, 0, c->stream>>>(m.nshared, m.send_seg, d_send_base, d_send_cnt, m.seg_off,
                                                               m.seg_idx, u, stride, m.sendbuf, skip);
    nsb_count_launch();
    NSB_NCCL(ncclGroupStart());
    for (int i = 0; i < m.nnbr; ++i) {
      int cnt = m.nbr_off[i + 1] - m.nbr_off[i];
      NSB_NCCL(ncclSend(m.sendbuf + 3 * m.nbr_off[i], (size_t)NF * cnt, ncclDouble, m.nbr_rank[i], c->comm, c->stream));
      NSB_NCCL(ncclRecv(m.recvbuf + 3 * m.nbr_off[i], (size_t)NF * cnt, ncclDouble, m.nbr_rank[i], c->comm, c->stream));
    }
    NSB_NCCL(ncclGroupEnd());
    if (m.nseg > 0) {
      k_gs_sum<NF, true><<<(m.nseg + T - 1) / T, T, 0, c->stream>>>(m.nseg, m.seg_off, m.seg_idx, m.rseg_off, m.rseg_pos,
                                                                   d_rseg_cnt, m.rseg_nbefore, m.recvbuf, u, stride, skip);
      nsb_count_launch();
    }
  } else if (m.nseg > 0) {
    k_gs_sum<NF, false><<<(m.nseg + T - 1) / T, T, 0, c->stream>>>(m.nseg, m.seg_off, m.seg_idx, nullptr, nullptr, nullptr,
                                                                  nullptr, nullptr, u, stride, skip);
    nsb_count_launch();
  }
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int gs_dssum(Ctx* c, double* u, int nfields, long long stride, const CGState* skip) {
  switch (nfields) {
    case 1: return dssum_nf<1>(c, u, stride, skip);
    case 2: return dssum_nf<2>(c, u, stride, skip);
    case 3: return dssum_nf<3>(c, u, stride, skip);
  }
  nsb_set_error("gs_dssum: nfields must be 1..3");
  return 1;
}
