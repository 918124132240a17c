// p2p.cu -- NVLink peer-memory collectives of the multi-GPU hot path (no NCCL call inside the time step).
//
// Each rank owns one "arena" in device memory, mapped into every other rank's address space with CUDA IPC at setup
// (handles are exchanged once over the NCCL communicator).  Two collectives of the hot path run on it:
//   * dssum halo: the pack kernel adds the local copies of every interface node and STORES the partial sums straight
//     into the neighbour's arena over NVLink, then the last CTA raises an epoch flag there; the neighbour's segmented-sum
//     kernel waits on the flag and folds the partials in (ascending rank order => bit-identical sums on all ranks).
//   * all-reduce of CG / inner-product scalars: every rank stores its deterministic local sums into slot [my rank] of
//     every arena, raises a flag, waits for all flags and adds the slots in rank order; for the CG loops the scalar
//     update (alpha, beta, convergence flag) happens in the same one-CTA kernel.
// Replaces grouped ncclSend/ncclRecv + ncclAllReduce (two host-enqueued collectives and two extra kernels per pressure
// iteration) by remote stores issued from the producing kernels.  The exchange counters (epochs) live in device memory, so no kernel
// argument changes between exchanges and whole CG batches replay as CUDA graphs on every rank.  Flow control: two parity copies of every buffer; a rank
// can never be more than one exchange ahead of a rank it exchanges with, because each exchange needs the partner's
// data of the same epoch.  Every spin-wait has a time-out that raises an error flag instead of hanging the GPU.
#include <algorithm>

#include "nsb_internal.h"

static constexpr int RSLOT = 4224;                // doubles per rank slot of the all-reduce buffer (>= 4096 aggregate sums + 3 CG scalars)
static constexpr long long SPIN_TIMEOUT_NS = 30000000000LL;   // 30 s: host-side skew between ranks (setup, numpy work) is seconds at most

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ bool p2p_wait(const unsigned long long* flag, unsigned long long epoch, int* err) {
  const long long t0 = gtime();
  while (ld_flag(flag) < epoch) {
    if (gtime() - t0 > SPIN_TIMEOUT_NS) { *err = 1; return false; }
  }
  return true;
}

// arena layout (units: 8 bytes)
static inline long long off_red(int parity, int rank, int nranks) { return ((long long)parity * nranks + rank) * RSLOT; }
static inline long long off_redflag(int parity, int rank, int nranks) { return 2LL * nranks * RSLOT + (long long)parity * nranks + rank; }
static inline long long off_haloflag(int parity, int rank, int nranks) { return 2LL * nranks * RSLOT + 2LL * nranks + (long long)parity * nranks + rank; }
static inline long long off_halo(int nranks) { return 2LL * nranks * RSLOT + 4LL * nranks; }

int p2p_free(Ctx* c, P2P& p) {
  if (!p.on) return 0;
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && p.peer[r]) cudaIpcCloseMemHandle(p.peer[r]);
  cudaFree(p.arena); cudaFree(p.d_peer_dst); cudaFree(p.d_peer_flag); cudaFree(p.d_nbr_rank); cudaFree(p.d_err);
  cudaFree(p.d_send_nbr); cudaFree(p.d_send_j); cudaFree(p.d_cnt); cudaFree(p.d_peer_base); cudaFree(p.d_epoch);
  p = P2P();
  return 0;
}

// called at the end of gs_setup (multi-rank); send_nbr/send_j: neighbour index and position of every send entry
int p2p_setup(Ctx* c, P2P& p, const GSMap& m, const std::vector<int>& send_nbr, const std::vector<int>& send_j) {
  const char* env = getenv("NSB_P2P");
  if (c->nranks <= 1 || c->nranks > 16) return 0;         // the same on every rank
  const int R = c->nranks;
  // The decision "peer-memory path usable" must be identical on all ranks, or some would enter the collectives below while
  // others fall back to NCCL: exchange the real device ordinals (CUDA_VISIBLE_DEVICES may remap them: a peer's ordinal is
  // not its rank), test peer access locally, then take the MINIMUM of the local verdicts over the communicator.
  int* d_dev = nullptr;
  NSB_CUDA(cudaMalloc(&d_dev, sizeof(int) * (R + 1)));
  NSB_CUDA(cudaMemcpy(d_dev + R, &c->device, sizeof(int), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllGather(d_dev + R, d_dev, 1, ncclInt32, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<int> devs(R);
  NSB_CUDA(cudaMemcpy(devs.data(), d_dev, sizeof(int) * R, cudaMemcpyDeviceToHost));
  int ok = (env && env[0] == '0') ? 0 : 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) ok = 0;
  for (int r = 0; r < R && ok; ++r) {
    if (r == c->rank) continue;
    int can = 0;
    // one process per GPU on ONE node: the peer's ordinal must be visible here and distinct from ours
    if (devs[r] < 0 || devs[r] >= ndev || devs[r] == c->device) { ok = 0; break; }
    if (cudaDeviceCanAccessPeer(&can, c->device, devs[r]) != cudaSuccess || !can) ok = 0;
  }
  NSB_CUDA(cudaMemcpy(d_dev, &ok, sizeof(int), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllReduce(d_dev, d_dev, 1, ncclInt32, ncclMin, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  NSB_CUDA(cudaMemcpy(&ok, d_dev, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_dev);
  if (!ok) return 0;                                       // every rank falls back to NCCL together
  // allgather nshared and the table "where does rank q receive data from rank r" (in doubles, -1: not a neighbour)
  std::vector<long long> mine(R + 1, -1), all((size_t)R * (R + 1));
  for (int i = 0; i < m.nnbr; ++i) mine[m.nbr_rank[i]] = 3LL * m.nbr_off[i];
  mine[R] = m.nshared;
  long long *d_my = nullptr, *d_all = nullptr;
  NSB_CUDA(cudaMalloc(&d_my, sizeof(long long) * (R + 1)));
  NSB_CUDA(cudaMalloc(&d_all, sizeof(long long) * (R + 1) * R));
  NSB_CUDA(cudaMemcpy(d_my, mine.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllGather(d_my, d_all, R + 1, ncclInt64, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  NSB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(long long) * all.size(), cudaMemcpyDeviceToHost));
  cudaFree(d_my); cudaFree(d_all);
  // arena
  p.halo_stride = 3LL * std::max(m.nshared, 1);
  p.words = off_halo(R) + 2 * p.halo_stride;
  NSB_CUDA(cudaMalloc(&p.arena, p.words * sizeof(double)));
  NSB_CUDA(cudaMemset(p.arena, 0, p.words * sizeof(double)));
  cudaIpcMemHandle_t h;
  NSB_CUDA(cudaIpcGetMemHandle(&h, p.arena));
  char *d_h = nullptr, *d_hall = nullptr;
  NSB_CUDA(cudaMalloc(&d_h, sizeof(h)));
  NSB_CUDA(cudaMalloc(&d_hall, sizeof(h) * R));
  NSB_CUDA(cudaMemcpy(d_h, &h, sizeof(h), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllGather(d_h, d_hall, sizeof(h), ncclChar, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<cudaIpcMemHandle_t> hs(R);
  NSB_CUDA(cudaMemcpy(hs.data(), d_hall, sizeof(h) * R, cudaMemcpyDeviceToHost));
  cudaFree(d_h); cudaFree(d_hall);
  for (int r = 0; r < R; ++r) {
    if (r == c->rank) { p.peer[r] = p.arena; continue; }
    void* ptr = nullptr;
    NSB_CUDA(cudaIpcOpenMemHandle(&ptr, hs[r], cudaIpcMemLazyEnablePeerAccess));
    p.peer[r] = (double*)ptr;
  }
  // per neighbour: destination (parity 0) inside the neighbour's arena, its halo stride, its flag slot for me
  std::vector<double*> dst(2 * std::max(m.nnbr, 1));
  std::vector<unsigned long long*> flg(2 * std::max(m.nnbr, 1));
  std::vector<int> cnt(std::max(m.nnbr, 1));
  for (int i = 0; i < m.nnbr; ++i) {
    const int q = m.nbr_rank[i];
    const long long where = all[(size_t)q * (R + 1) + c->rank];        // offset inside q's halo block for data from me
    const long long qstride = 3LL * std::max<long long>(all[(size_t)q * (R + 1) + R], 1);
    if (where < 0) { nsb_set_error("p2p_setup: asymmetric neighbour table (rank %d <-> %d)", c->rank, q); return 1; }
    for (int par = 0; par < 2; ++par) {
      dst[par * m.nnbr + i] = p.peer[q] + off_halo(R) + par * qstride + where;
      flg[par * m.nnbr + i] = (unsigned long long*)(p.peer[q] + off_haloflag(par, c->rank, R));
    }
    cnt[i] = m.nbr_off[i + 1] - m.nbr_off[i];
  }
  NSB_CUDA(cudaMalloc(&p.d_peer_dst, dst.size() * sizeof(double*)));
  NSB_CUDA(cudaMemcpy(p.d_peer_dst, dst.data(), dst.size() * sizeof(double*), cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMalloc(&p.d_peer_flag, flg.size() * sizeof(void*)));
  NSB_CUDA(cudaMemcpy(p.d_peer_flag, flg.data(), flg.size() * sizeof(void*), cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMalloc(&p.d_cnt, cnt.size() * sizeof(int)));
  NSB_CUDA(cudaMemcpy(p.d_cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice));
  std::vector<int> nr(std::max(m.nnbr, 1), 0);
  for (int i = 0; i < m.nnbr; ++i) nr[i] = m.nbr_rank[i];
  NSB_CUDA(cudaMalloc(&p.d_nbr_rank, nr.size() * sizeof(int)));
  NSB_CUDA(cudaMemcpy(p.d_nbr_rank, nr.data(), nr.size() * sizeof(int), cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMalloc(&p.d_send_nbr, std::max<size_t>(send_nbr.size(), 1) * sizeof(int)));
  NSB_CUDA(cudaMalloc(&p.d_send_j, std::max<size_t>(send_j.size(), 1) * sizeof(int)));
  if (!send_nbr.empty()) {
    NSB_CUDA(cudaMemcpy(p.d_send_nbr, send_nbr.data(), send_nbr.size() * sizeof(int), cudaMemcpyHostToDevice));
    NSB_CUDA(cudaMemcpy(p.d_send_j, send_j.data(), send_j.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  NSB_CUDA(cudaMalloc(&p.d_err, 4 * sizeof(int)));
  NSB_CUDA(cudaMemset(p.d_err, 0, 4 * sizeof(int)));
  NSB_CUDA(cudaMalloc(&p.d_epoch, 2 * sizeof(unsigned long long)));       // [0] halo, [1] all-reduce exchange counters
  NSB_CUDA(cudaMemset(p.d_epoch, 0, 2 * sizeof(unsigned long long)));
  // peer table of arena bases for the all-reduce kernel
  NSB_CUDA(cudaMalloc(&p.d_peer_base, 16 * sizeof(double*)));
  NSB_CUDA(cudaMemcpy(p.d_peer_base, p.peer, 16 * sizeof(double*), cudaMemcpyHostToDevice));
  // everybody must have mapped everybody before the first remote store
  double* d_tok = nullptr;
  NSB_CUDA(cudaMalloc(&d_tok, sizeof(double)));
  NSB_CUDA(cudaMemset(d_tok, 0, sizeof(double)));
  NSB_NCCL(ncclAllReduce(d_tok, d_tok, 1, ncclDouble, ncclSum, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tok);
  p.on = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ halo
// The epoch of every exchange lives in DEVICE memory (ep[0]: halo, ep[1]: all-reduce): the first kernel of an exchange uses ep + 1 and
// its last CTA stores it back, the consumer kernel reads the stored value.  No kernel argument changes from one exchange to the next,
// so whole CG batches can be captured in a CUDA graph and replayed on several ranks (r2; r1 passed the epoch as an argument).
template <int NF>
__global__ void k_gs_pack_p2p(int nshared, const int* __restrict__ send_seg, const int* __restrict__ send_nbr,
                              const int* __restrict__ send_j, const int* __restrict__ nbr_cnt, const int* __restrict__ seg_off,
                              const int* __restrict__ seg_idx, const double* __restrict__ u, long long stride,
                              double* const* __restrict__ peer_dst2, unsigned long long* const* __restrict__ peer_flag2, int nnbr,
                              unsigned long long* ep, unsigned* counter, const CGState* skip) {
  if (skip && skip->done) return;
  const unsigned long long epoch = *(volatile unsigned long long*)ep + 1;   // read HERE, before this CTA arrives at the counter below
  const int par = (int)(epoch & 1);
  double* const* peer_dst = peer_dst2 + par * nnbr;
  unsigned long long* const* peer_flag = peer_flag2 + par * nnbr;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nshared) {
    const int seg = send_seg[s];
    const int a = seg_off[seg], b = seg_off[seg + 1];
    double acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.0;
    for (int j = a; j < b; ++j) {
      const int idx = seg_idx[j];
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] += u[(long long)f * stride + idx];
    }
    const int nb = send_nbr[s];
    double* dst = peer_dst[nb] + send_j[s];
    const int cnt = nbr_cnt[nb];
#pragma unroll
    for (int f = 0; f < NF; ++f) dst[(long long)f * cnt] = acc[f];        // remote store over NVLink
  }
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicInc(counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < nnbr) {
      __threadfence_system();
      st_flag(peer_flag[threadIdx.x], epoch);
    }
    if (threadIdx.x == 0) ep[0] = epoch;                 // every CTA has read ep[0] before arriving at the counter
  }
}

// segments without remote copies: plain local segmented sum (same arithmetic as k_gs_sum<NF,false> of gs.cu)
template <int NF>
__global__ void k_gs_sum_local(int nseg, const int* __restrict__ seg_off, const int* __restrict__ seg_idx, double* __restrict__ u,
                               long long stride, const CGState* skip) {
  if (skip && skip->done) return;
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= nseg) return;
  const int a = seg_off[seg], b = seg_off[seg + 1];
  double acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = 0.0;
  for (int j = a; j < b; ++j) {
    const int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] += u[(long long)f * stride + idx];
  }
  for (int j = a; j < b; ++j) {
    const int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) u[(long long)f * stride + idx] = acc[f];
  }
}

template <int NF>
__global__ void k_gs_sum_p2p(int nseg, const int* __restrict__ seg_off, const int* __restrict__ seg_idx,
                             const int* __restrict__ rseg_off, const int* __restrict__ rseg_pos, const int* __restrict__ rseg_cnt,
                             const int* __restrict__ rseg_nbefore, const double* __restrict__ recvbuf, double* __restrict__ u,
                             long long stride, const unsigned long long* __restrict__ flags2, const int* __restrict__ nbr_rank,
                             int nnbr, const unsigned long long* ep, long long halo_stride, int nranks, int* err,
                             const CGState* skip, int seg0) {
  if (skip && skip->done) return;
  const unsigned long long epoch = *(const volatile unsigned long long*)ep;   // stored by the pack kernel of this exchange
  const int par = (int)(epoch & 1);
  recvbuf += (long long)par * halo_stride;              // parity copy of the halo block and of its flags
  const unsigned long long* flags = flags2 + (long long)par * nranks;
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < nnbr)
    if (!p2p_wait(flags + nbr_rank[threadIdx.x], epoch, err)) ok = 0;
  __syncthreads();
  if (!ok) return;
  const int seg = seg0 + blockIdx.x * blockDim.x + threadIdx.x;      // only the segments shared with other ranks: [seg0, nseg)
  if (seg >= nseg) return;
  const int a = seg_off[seg], b = seg_off[seg + 1];
  double acc[NF], loc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) { acc[f] = 0.0; loc[f] = 0.0; }
  for (int j = a; j < b; ++j) {
    const int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) loc[f] += u[(long long)f * stride + idx];
  }
  const int ra = rseg_off[seg], rb = rseg_off[seg + 1], nb = rseg_nbefore[seg];
  for (int j = ra; j < ra + nb; ++j) {
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] += __ldcv(&recvbuf[rseg_pos[j] + f * rseg_cnt[j]]);
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = (nb > 0) ? acc[f] + loc[f] : loc[f];
  for (int j = ra + nb; j < rb; ++j) {
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] += __ldcv(&recvbuf[rseg_pos[j] + f * rseg_cnt[j]]);
  }
  for (int j = a; j < b; ++j) {
    const int idx = seg_idx[j];
#pragma unroll
    for (int f = 0; f < NF; ++f) u[(long long)f * stride + idx] = acc[f];
  }
}

template <int NF>
static int dssum_p2p_nf(Ctx* c, P2P& p, GSMap& m, double* u, long long stride, const CGState* skip) {
  const int* d_send_seg = m.send_seg;
  const int* d_rseg_cnt = m.rseg_cnt;
  const int T = 128, R = c->nranks;
  if (m.nshared > 0) {
    k_gs_pack_p2p<NF><<<(m.nshared + T - 1) / T, T, 0, c->stream>>>(m.nshared, d_send_seg, p.d_send_nbr, p.d_send_j, p.d_cnt, m.seg_off,
                                                                   m.seg_idx, u, stride, p.d_peer_dst, p.d_peer_flag, m.nnbr, p.d_epoch,
                                                                   c->red_count + 2, skip);
    nsb_count_launch();
  }
  // interior segments (no copy on another rank) are summed while the partial sums travel over NVLink ...
  if (m.nseg_int > 0) {
    k_gs_sum_local<NF><<<(m.nseg_int + T - 1) / T, T, 0, c->stream>>>(m.nseg_int, m.seg_off, m.seg_idx, u, stride, skip);
    nsb_count_launch();
  }
  // ... and only the shared segments wait for the neighbours' flags
  if (m.nseg > m.nseg_int) {
    const double* recv = p.arena + off_halo(R);                                                     // parity 0; the kernel adds par * stride
    const unsigned long long* flags = (const unsigned long long*)(p.arena + off_haloflag(0, 0, R));
    const int ns = m.nseg - m.nseg_int;
    k_gs_sum_p2p<NF><<<(ns + T - 1) / T, T, 0, c->stream>>>(m.nseg, m.seg_off, m.seg_idx, m.rseg_off, m.rseg_pos, d_rseg_cnt,
                                                           m.rseg_nbefore, recv, u, stride, flags, p.d_nbr_rank, m.nnbr, p.d_epoch,
                                                           p.halo_stride, R, p.d_err, skip, m.nseg_int);
    nsb_count_launch();
  }
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int p2p_dssum(Ctx* c, P2P& p, GSMap& m, double* u, int nfields, long long stride, const CGState* skip) {
  switch (nfields) {
    case 1: return dssum_p2p_nf<1>(c, p, m, u, stride, skip);
    case 2: return dssum_p2p_nf<2>(c, p, m, u, stride, skip);
    case 3: return dssum_p2p_nf<3>(c, p, m, u, stride, skip);
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------ all-reduce (+ CG scalar update)
__device__ void cg_apply(CGState* s, const double* sums, int ncomp, int kind) {
  for (int f = 0; f < ncomp; ++f) {
    if (kind == 0) {
      s[f].rtz1 = sums[2 * f]; s[f].rtz2 = 1.0; s[f].beta = 0.0; s[f].alpha = 0.0;
      s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
      s[f].iter = 0;
      s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].maxit <= 0);
    } else if (kind == 1) {
      if (s[f].done) continue;
      s[f].rtz2 = s[f].rtz1; s[f].rtz1 = sums[2 * f]; s[f].beta = s[f].rtz1 / s[f].rtz2;
      s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
      s[f].iter += 1;
      s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].iter >= s[f].maxit) || !(s[f].rnorm == s[f].rnorm);
    } else if (kind == 2) {
      if (!s[f].done) { s[f].rho = sums[f]; s[f].alpha = s[f].rtz1 / sums[f]; }
    } else if (kind == 3) {          // init, residual norm only (separate preconditioner delivers z^T r with kind 5)
      s[f].rtz1 = 1.0; s[f].rtz2 = 1.0; s[f].beta = 0.0; s[f].alpha = 0.0;
      s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
      s[f].iter = 0;
      s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].maxit <= 0);
    } else if (kind == 4) {
      if (s[f].done) continue;
      s[f].rtz2 = s[f].rtz1;
      s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
      s[f].iter += 1;
      s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].iter >= s[f].maxit) || !(s[f].rnorm == s[f].rnorm);
    } else if (!s[f].done) {         // kind 5
      s[f].rtz1 = sums[f]; s[f].beta = (s[f].iter == 0) ? 0.0 : sums[f] / s[f].rtz2;
    }
  }
}

// one CTA; op 0: sum, 1: max.  vals (count <= RSLOT) are reduced in place over all ranks, in rank order.
__global__ void k_p2p_allreduce(double* vals, int count, int op, double* const* __restrict__ peer_base, double* arena, int rank,
                                int nranks, unsigned long long* ep, int* err, CGState* cgs, int ncomp, int kind, const CGState* skip) {
  if (skip && skip->done) return;
  __shared__ int ok;
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) { s_epoch = ep[1] + 1; ep[1] = s_epoch; }        // one CTA: the epoch of this exchange
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int parity = (int)(epoch & 1);
  const long long slot_off = ((long long)parity * nranks) * RSLOT;                       // off_red(parity, 0, nranks)
  const long long flag_off = 2LL * nranks * RSLOT + (long long)parity * nranks;          // off_redflag(parity, 0, nranks)
  if (threadIdx.x == 0) ok = 1;
  // 1) publish my values in slot [rank] of every arena
  for (int i = threadIdx.x; i < count * nranks; i += blockDim.x) {
    const int r = i / count, k = i - r * count;
    peer_base[r][slot_off + (long long)rank * RSLOT + k] = vals[k];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < nranks) {
    unsigned long long* f = (unsigned long long*)(peer_base[threadIdx.x] + flag_off + rank);
    st_flag(f, epoch);
  }
  // 2) wait for everybody's contribution
  if (threadIdx.x < nranks)
    if (!p2p_wait((const unsigned long long*)(arena + flag_off + threadIdx.x), epoch, err)) ok = 0;
  __syncthreads();
  if (!ok) return;
  // 3) combine in rank order (identical on every rank)
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    double a = __ldcv(&arena[slot_off + k]);
    for (int r = 1; r < nranks; ++r) {
      const double b = __ldcv(&arena[slot_off + (long long)r * RSLOT + k]);
      a = op ? fmax(a, b) : a + b;
    }
    vals[k] = a;
  }
  if (cgs) {
    __syncthreads();
    if (threadIdx.x == 0) cg_apply(cgs, vals, ncomp, kind);
  }
}

// Wide all-reduce (the aggregate sums of the pressure preconditioner: up to 4096 + 3 values per iteration).  r2 measurement at 8 GPUs:
// the one-CTA kernel above needs 0.11 ms for 4099 values (32 k remote stores issued by 256 threads).  Here CTA b of the first kernel
// pushes this rank's values into peer b's slot and the last CTA to finish raises the flags; the second kernel waits for the peers'
// flags and adds the slots in rank order with one thread per value.
__global__ void __launch_bounds__(512) k_p2p_ar_publish(const double* __restrict__ vals, int count, double* const* __restrict__ peer_base,
                                                        int rank, int nranks, unsigned long long* ep, unsigned* counter) {
  const unsigned long long epoch = *((volatile unsigned long long*)ep + 1) + 1;   // read HERE, before this CTA arrives at the counter
  const int parity = (int)(epoch & 1);
  const long long slot_off = ((long long)parity * nranks) * RSLOT;
  const long long flag_off = 2LL * nranks * RSLOT + (long long)parity * nranks;
  double* dst = peer_base[blockIdx.x] + slot_off + (long long)rank * RSLOT;
  for (int k = threadIdx.x; k < count; k += blockDim.x) dst[k] = vals[k];          // remote stores over NVLink (local for b == rank)
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicInc(counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < nranks) {
      __threadfence_system();
      st_flag((unsigned long long*)(peer_base[threadIdx.x] + flag_off + rank), epoch);
    }
    if (threadIdx.x == 0) ep[1] = epoch;
  }
}
__global__ void __launch_bounds__(256) k_p2p_ar_combine(double* __restrict__ vals, int count, int op, const double* __restrict__ arena,
                                                        int nranks, const unsigned long long* ep, int* err) {
  const unsigned long long epoch = *((const volatile unsigned long long*)ep + 1);   // stored by the publish kernel of this exchange
  const int parity = (int)(epoch & 1);
  const long long slot_off = ((long long)parity * nranks) * RSLOT;
  const long long flag_off = 2LL * nranks * RSLOT + (long long)parity * nranks;
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < nranks)
    if (!p2p_wait((const unsigned long long*)(arena + flag_off + threadIdx.x), epoch, err)) ok = 0;
  __syncthreads();
  if (!ok) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double a = __ldcv(&arena[slot_off + k]);
  for (int r = 1; r < nranks; ++r) {
    const double b = __ldcv(&arena[slot_off + (long long)r * RSLOT + k]);
    a = op ? fmax(a, b) : a + b;
  }
  vals[k] = a;
}

int p2p_allreduce(Ctx* c, double* dev, int count, int op, CGState* cgs, int ncomp, int kind) {
  P2P& p = c->p2p;
  const int R = c->nranks;
  if (!cgs && count > 64) {
    for (int done = 0; done < count; done += RSLOT) {
      const int cnt = std::min(RSLOT, count - done);
      k_p2p_ar_publish<<<R, 512, 0, c->stream>>>(dev + done, cnt, p.d_peer_base, c->rank, R, p.d_epoch, c->red_count + 3);
      k_p2p_ar_combine<<<(cnt + 255) / 256, 256, 0, c->stream>>>(dev + done, cnt, op, p.arena, R, p.d_epoch, p.d_err);
      nsb_count_launch(2);
    }
    NSB_CUDA(cudaGetLastError());
    return 0;
  }
  for (int done = 0; done < count; done += RSLOT) {
    const int cnt = std::min(RSLOT, count - done);
    k_p2p_allreduce<<<1, 256, 0, c->stream>>>(dev + done, cnt, op, p.d_peer_base, p.arena, c->rank, R, p.d_epoch, p.d_err,
                                              (done + RSLOT >= count) ? cgs : nullptr, ncomp, kind, nullptr);
    nsb_count_launch();
  }
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int p2p_check_error(Ctx* c) {
  for (P2P* p : {&c->p2p, &c->p2pv, &c->p2pp}) {
    if (!p->on) continue;
    int e[4] = {0, 0, 0, 0};
    NSB_CUDA(cudaMemcpy(e, p->d_err, sizeof(e), cudaMemcpyDeviceToHost));
    if (e[0]) { nsb_set_error("peer-memory collective timed out (a rank stopped participating)"); return 3; }
  }
  return 0;
}
