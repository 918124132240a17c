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

#include "elem_common.cuh"


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
  int seg = blockIdx.x * blockDim.x + threadIdx.x;      // nseg may be a leading sub-range (the interior segments)
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

// Note (r1c, measured): splitting the two-copy (face) segments into a separate offset-free pair list, to shorten the
// dependent load chain, made the 3-field dssum SLOWER (0.173 vs 0.117 ms on cfg 5): face, edge and vertex nodes share
// 32-byte sectors, and processing them in one dof-ordered sweep is what keeps every sector to one read and one write.
template <class T>
static int upload(T** dptr, const std::vector<T>& h) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  NSB_CUDA(cudaMalloc((void**)dptr, bytes));
  if (!h.empty()) NSB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int gs_free_map(Ctx* c, GSMap& m, P2P& p) {
  p2p_free(c, p);
  cudaFree(m.seg_off); cudaFree(m.seg_idx); cudaFree(m.send_seg); cudaFree(m.rseg_off); cudaFree(m.rseg_pos);
  cudaFree(m.rseg_nbefore); cudaFree(m.sendbuf); cudaFree(m.recvbuf);
  cudaFree(m.send_base); cudaFree(m.send_cnt); cudaFree(m.rseg_cnt);
  m = GSMap();
  return 0;
}
int gs_free(Ctx* c) {
  gs_free_map(c, c->gsv, c->p2pv);
  gs_free_map(c, c->gsp, c->p2pp);
  return gs_free_map(c, c->gs, c->p2p);
}

// ------------------------------------------------------------------------------------------------ host-side plan
// Everything below up to gs_setup is pure host code (no CUDA, no NCCL) so that the multi-rank map construction can be
// exercised on a CPU-only box (tests/test_multirank_gloo.py drives it through nsb_gs_host_* with gloo as the transport).
struct HostPlan {
  std::vector<int> order, ustart;            // dofs sorted by (global id, dof); run starts per unique id
  std::vector<long long> uid;
  std::vector<int> seg_off, seg_idx;         // segments (sorted by smallest local dof)
  std::vector<int> nbr_rank, nbr_off;        // neighbours (ascending rank), offsets into the send/recv buffers
  std::vector<int> send_seg, send_base, send_cnt;
  std::vector<int> rseg_off, rseg_pos, rseg_cnt, nbefore;
  int nshared = 0;
  int nseg_int = 0;                          // segments [0, nseg_int) have no copy on another rank
};

static bool is_surface(int p, int N, int D) {
  int i = p % N, j = (p / N) % N, kk = (D == 3) ? p / (N * N) : 1;
  return i == 0 || i == N - 1 || j == 0 || j == N - 1 || (D == 3 && (kk == 0 || kk == N - 1));
}

static void plan_sort(HostPlan& P, long long n, const long long* glo) {
  P.order.resize(n);
  std::iota(P.order.begin(), P.order.end(), 0);
  std::sort(P.order.begin(), P.order.end(), [&](int a, int b) { return glo[a] != glo[b] ? glo[a] < glo[b] : a < b; });
  P.uid.clear(); P.ustart.clear();
  for (long long i = 0; i < n; ++i)
    if (i == 0 || glo[P.order[i]] != glo[P.order[i - 1]]) { P.uid.push_back(glo[P.order[i]]); P.ustart.push_back((int)i); }
  P.ustart.push_back((int)n);
}

// ids of this rank's element-surface nodes (ascending): the only nodes another rank can share
static void plan_candidates(const HostPlan& P, int N, int D, int np, std::vector<long long>& cand, int surf_first = 0) {
  cand.clear();
  const int nu = (int)P.uid.size();
  const int ns = np - ((D == 3) ? (N - 2) * (N - 2) * (N - 2) : (N - 2) * (N - 2));
  for (int k = 0; k < nu; ++k) {
    const int q = P.order[P.ustart[k]] % np;
    if (surf_first ? (q < ns) : is_surface(q, N, D)) cand.push_back(P.uid[k]);      // surface-first layout: the surface block leads
  }
}

// counts[r], ids (concatenated, each rank's list ascending) = every rank's candidate list
static void plan_build(HostPlan& P, int rank, int nranks, const long long* counts, const long long* ids) {
  const int nu = (int)P.uid.size();
  std::vector<std::vector<int>> shared_u(nranks);
  long long base = 0;
  for (int r = 0; r < nranks; ++r) {
    const long long no = counts ? counts[r] : 0;
    if (r != rank && counts) {
      const long long* other = ids + base;
      long long a = 0;
      int k = 0;
      while (a < no && k < nu) {                       // both ascending: two-pointer intersection
        if (other[a] < P.uid[k]) ++a;
        else if (other[a] > P.uid[k]) ++k;
        else { shared_u[r].push_back(k); ++a; ++k; }
      }
    }
    base += no;
  }
  std::vector<char> is_shared(nu, 0);
  for (int r = 0; r < nranks; ++r)
    for (int k : shared_u[r]) is_shared[k] = 1;
  std::vector<int> seg_of_u(nu, -1);
  P.seg_off.assign(1, 0); P.seg_idx.clear();
  {
    // order the segments by their smallest local dof: neighbouring threads then touch neighbouring memory
    // (r1a ncu: with global-id order the 3-field dssum moved 2x the algorithmic bytes at 20 % of the HBM roofline)
    std::vector<int> segk;
    for (int k = 0; k < nu; ++k)
      if (P.ustart[k + 1] - P.ustart[k] > 1 || is_shared[k]) segk.push_back(k);
    // segments shared with another rank go LAST: the peer-memory path sums the interior range while the halo is in flight
    // and only the short shared range waits for the neighbours' flags (r2: the flag wait cost 27 us per dssum at 2 GPUs)
    std::sort(segk.begin(), segk.end(), [&](int a, int b) {
      if (is_shared[a] != is_shared[b]) return is_shared[a] < is_shared[b];
      return P.order[P.ustart[a]] < P.order[P.ustart[b]];
    });
    P.nseg_int = 0;
    for (int k : segk) P.nseg_int += is_shared[k] ? 0 : 1;
    for (int k : segk) {
      seg_of_u[k] = (int)P.seg_off.size() - 1;
      for (int j = P.ustart[k]; j < P.ustart[k + 1]; ++j) P.seg_idx.push_back(P.order[j]);
      P.seg_off.push_back((int)P.seg_idx.size());
    }
  }
  const int nseg = (int)P.seg_off.size() - 1;
  std::vector<std::vector<std::pair<int, int>>> rlist(nseg);   // per segment: (pos, cnt) in ascending rank
  P.nbefore.assign(nseg, 0);
  P.nbr_rank.clear(); P.nbr_off.clear(); P.send_seg.clear(); P.send_base.clear(); P.send_cnt.clear();
  int off = 0;
  for (int r = 0; r < nranks; ++r) {
    if (shared_u[r].empty()) continue;
    const int cnt = (int)shared_u[r].size();
    P.nbr_rank.push_back(r);
    P.nbr_off.push_back(off);
    for (int j = 0; j < cnt; ++j) {
      const int seg = seg_of_u[shared_u[r][j]];
      P.send_seg.push_back(seg);
      P.send_base.push_back(3 * off + j);
      P.send_cnt.push_back(cnt);
      rlist[seg].push_back({3 * off + j, cnt});
      if (r < rank) P.nbefore[seg]++;
    }
    off += cnt;
  }
  P.nbr_off.push_back(off);
  P.nshared = off;
  P.rseg_off.assign(1, 0); P.rseg_pos.clear(); P.rseg_cnt.clear();
  for (int sidx = 0; sidx < nseg; ++sidx) {
    for (auto& pr : rlist[sidx]) { P.rseg_pos.push_back(pr.first); P.rseg_cnt.push_back(pr.second); }
    P.rseg_off.push_back((int)P.rseg_pos.size());
  }
}

// ---- host-only debug/test entry points (declared in include/nekstab_b200.h)
static HostPlan g_host_plan;
extern "C" int nsb_gs_host_candidates(int ldim, int lx1, int nelv, const long long* glo_num, long long* ids_out,
                                      long long* count) {
  const int np = (ldim == 3) ? lx1 * lx1 * lx1 : lx1 * lx1;
  plan_sort(g_host_plan, (long long)nelv * np, glo_num);
  std::vector<long long> cand;
  plan_candidates(g_host_plan, lx1, ldim, np, cand);
  if (count) *count = (long long)cand.size();
  if (ids_out) memcpy(ids_out, cand.data(), cand.size() * sizeof(long long));
  return 0;
}
extern "C" int nsb_gs_host_plan(int rank, int nranks, const long long* counts, const long long* ids, int sizes_out[8]) {
  if (g_host_plan.order.empty()) { nsb_set_error("nsb_gs_host_plan: call nsb_gs_host_candidates first"); return 1; }
  plan_build(g_host_plan, rank, nranks, counts, ids);
  const HostPlan& P = g_host_plan;
  const int v[8] = {(int)P.seg_off.size() - 1, (int)P.seg_idx.size(), (int)P.nbr_rank.size(), P.nshared,
                    (int)P.rseg_pos.size(), P.nseg_int, 0, 0};
  for (int i = 0; i < 8; ++i) sizes_out[i] = v[i];
  return 0;
}
// which: 0 seg_off, 1 seg_idx, 2 nbr_rank, 3 nbr_off, 4 send_seg, 5 send_base, 6 send_cnt, 7 rseg_off, 8 rseg_pos, 9 rseg_cnt, 10 nbefore
extern "C" int nsb_gs_host_get(int which, int* out) {
  const HostPlan& P = g_host_plan;
  const std::vector<int>* v[11] = {&P.seg_off, &P.seg_idx, &P.nbr_rank, &P.nbr_off, &P.send_seg, &P.send_base, &P.send_cnt,
                                   &P.rseg_off, &P.rseg_pos, &P.rseg_cnt, &P.nbefore};
  if (which < 0 || which > 10) { nsb_set_error("nsb_gs_host_get: bad selector"); return 1; }
  memcpy(out, v[which]->data(), v[which]->size() * sizeof(int));
  return 0;
}

int gs_setup(Ctx* c, const long long* glo) {
  NSB_TRY(gs_build(c, c->gs, c->p2p, c->n, c->lx1, c->np1, glo));
  const char* env = getenv("NSB_PERM");
  if (c->ldim == 3 && c->lx1 == 8 && !(env && env[0] == '0')) {      // used by the Helmholtz loop of the lx1 = 8 path (k_axhelm3p)
    // second map for loop vectors kept in the surface-first element layout: same nodes, permuted local positions
    std::vector<long long> gp((size_t)c->n);
    std::vector<int> perm(c->np1);
    for (int q = 0; q < c->np1; ++q) perm[q] = surf_first_pos(c->lx1, q);
    for (int e = 0; e < c->nel; ++e)
      for (int q = 0; q < c->np1; ++q) gp[(size_t)e * c->np1 + perm[q]] = glo[(size_t)e * c->np1 + q];
    NSB_TRY(gs_build(c, c->gsp, c->p2pp, c->n, c->lx1, c->np1, gp.data(), 1));
    c->gsp_ready = true;
  }
  return 0;
}
int gs_dssum_w(Ctx* c, double* w, int nfields, bool permuted, const CGState* skip) {
  if (permuted) return gs_dssum_map(c, c->gsp, c->p2pp, w, nfields, c->n, skip);
  return gs_dssum_map(c, c->gs, c->p2p, w, nfields, c->n, skip);
}

// Build a gather-scatter map over `n` local dofs with global ids `glo` (np dofs per element on an N^ldim grid): the
// velocity mesh (N = lx1) or the element-vertex mesh of the pressure preconditioner (N = 2).
int gs_build(Ctx* c, GSMap& m, P2P& p2p, long long n, int N1, int np_e, const long long* glo, int surf_first) {
  if (n >= (1LL << 31)) { nsb_set_error("gs_setup: more than 2^31 local dofs"); return 1; }
  HostPlan P;
  plan_sort(P, n, glo);
  std::vector<long long> cnts(c->nranks, 0), all;
  if (c->nranks > 1) {
    std::vector<long long> cand;
    plan_candidates(P, N1, c->ldim, np_e, cand, surf_first);
    // allgather counts then ids (padded) through NCCL
    long long mycnt = (long long)cand.size();
    long long* d_cnt = nullptr;
    NSB_CUDA(cudaMalloc(&d_cnt, sizeof(long long) * (c->nranks + 1)));
    NSB_CUDA(cudaMemcpy(d_cnt + c->nranks, &mycnt, sizeof(long long), cudaMemcpyHostToDevice));
    NSB_NCCL(ncclAllGather(d_cnt + c->nranks, d_cnt, 1, ncclInt64, c->comm, c->stream));
    NSB_CUDA(cudaStreamSynchronize(c->stream));
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
    std::vector<long long> padded(mx * c->nranks);
    NSB_CUDA(cudaMemcpy(padded.data(), d_all, sizeof(long long) * padded.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_my); cudaFree(d_all);
    for (int r = 0; r < c->nranks; ++r) all.insert(all.end(), padded.begin() + (size_t)r * mx, padded.begin() + (size_t)r * mx + cnts[r]);
  }
  plan_build(P, c->rank, c->nranks, c->nranks > 1 ? cnts.data() : nullptr, all.data());
  m.nseg = (int)P.seg_off.size() - 1;
  m.nseg_int = P.nseg_int;
  m.nbr_rank = P.nbr_rank; m.nbr_off = P.nbr_off;
  m.nnbr = (int)m.nbr_rank.size();
  m.nshared = P.nshared;
  NSB_TRY(upload(&m.seg_off, P.seg_off));
  NSB_TRY(upload(&m.seg_idx, P.seg_idx));
  NSB_TRY(upload(&m.send_seg, P.send_seg));
  NSB_TRY(upload(&m.send_base, P.send_base));
  NSB_TRY(upload(&m.send_cnt, P.send_cnt));
  NSB_TRY(upload(&m.rseg_off, P.rseg_off));
  NSB_TRY(upload(&m.rseg_pos, P.rseg_pos));
  NSB_TRY(upload(&m.rseg_cnt, P.rseg_cnt));
  NSB_TRY(upload(&m.rseg_nbefore, P.nbefore));
  size_t hb = std::max<size_t>((size_t)3 * m.nshared, 1) * sizeof(double);
  NSB_CUDA(cudaMalloc(&m.sendbuf, hb));
  NSB_CUDA(cudaMalloc(&m.recvbuf, hb));
  NSB_CUDA(cudaMemset(m.sendbuf, 0, hb));
  NSB_CUDA(cudaMemset(m.recvbuf, 0, hb));
  if (c->nranks > 1) {
    std::vector<int> send_nbr, send_j;
    for (int i = 0; i < m.nnbr; ++i)
      for (int j = 0; j < m.nbr_off[i + 1] - m.nbr_off[i]; ++j) { send_nbr.push_back(i); send_j.push_back(j); }
    NSB_TRY(p2p_setup(c, p2p, m, send_nbr, send_j));
  }
  return 0;
}

template <int NF>
static int dssum_nf(Ctx* c, GSMap& m, P2P& p2p, double* u, long long stride, const CGState* skip) {
  const int T = 128;
  if (p2p.on) return p2p_dssum(c, p2p, m, u, NF, stride, skip);
  if (m.nshared > 0) {
    k_gs_pack<NF><<<(m.nshared + T - 1) / T, T, 0, c->stream>>>(m.nshared, m.send_seg, m.send_base, m.send_cnt, m.seg_off,
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
                                                                   m.rseg_cnt, m.rseg_nbefore, m.recvbuf, u, stride, skip);
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

int gs_dssum_map(Ctx* c, GSMap& m, P2P& p2p, double* u, int nfields, long long stride, const CGState* skip) {
  switch (nfields) {
    case 1: return dssum_nf<1>(c, m, p2p, u, stride, skip);
    case 2: return dssum_nf<2>(c, m, p2p, u, stride, skip);
    case 3: return dssum_nf<3>(c, m, p2p, u, stride, skip);
  }
  nsb_set_error("gs_dssum: nfields must be 1..3");
  return 1;
}
int gs_dssum(Ctx* c, double* u, int nfields, long long stride, const CGState* skip) {
  return gs_dssum_map(c, c->gs, c->p2p, u, nfields, stride, skip);
}
