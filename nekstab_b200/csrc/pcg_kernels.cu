// pcg_kernels.cu -- second-generation 3-D kernels of the pressure operator E = D (mask B~^-1 QQ^T) D^T, the
// operator applied thousands of times per time step by the Jacobi-PCG pressure solve (>95 % of a step).
//
//   k_gradt3 : w_c = D_c^T p      mesh 2 (GL, lx2^3) -> mesh 1 (GLL, lx1^3)   [UPSTREAM navier1.f opgradt/cdtp]
//   k_div3   : q = sum_c D_c u_c  mesh 1 -> mesh 2                              [UPSTREAM navier1.f opdiv/multd]
// optionally fused with the CG vector work (direction update before gradt; p.Ep partial after div).  (r1b: a variant of k_div3
// that summed the copies of every element-surface node itself from a per-element gather table -- no dssum launch -- was measured
// slower than the separate segment-sorted dssum, 0.465 vs 0.115 + 0.238 ms, and was removed in r2.)
//
// r1a ncu (profiles/r1a_*): the first-generation kernels were bound by the shared-memory pipe (LSU wavefronts 89 %,
// barrier stalls), not by DRAM.  Changes here: (1) all three components go through each tensor stage together
// (4 barriers per launch instead of 12); (2) the two matrices that feed one output are applied in one pass with
// register accumulation (no shared-memory read-modify-write); (3) the metric multiply is folded into the first
// (gradt) / last (div) stage with the metrics read straight from global memory in column order; (4) gradt's last
// stage stores to global memory directly.  Shared-memory accesses per element drop by ~half.
#include <cuda_pipeline.h>

#include <algorithm>
#include <cstdint>

#include "elem_common.cuh"

namespace {

// 384 threads: every tensor stage is at most ONE column task per thread.  (A task *loop* lets the compiler hoist the
// loop-invariant constant-bank matrix operands into registers: 150+ registers or heavy spills, see DESIGN.md.)
constexpr int PK_TPB = 384;

// one column of a pitched (NK,NJ,NI) array along axis AX
template <int AX, int NK, int NJ, int NI>
struct ColIn {
  static constexpr int PI = OddPitch<NI>::v;
  static constexpr int ncol = (AX == 0) ? NK * NJ : ((AX == 1) ? NK * NI : NJ * NI);
  static constexpr int stride = (AX == 0) ? 1 : ((AX == 1) ? PI : NJ * PI);
  __device__ __forceinline__ static int base(int c) {
    if (AX == 0) return c * PI;
    if (AX == 1) return (c / NI) * NJ * PI + (c % NI);
    return (c / NI) * PI + (c % NI);
  }
};

template <int NO, int NL>
__device__ __forceinline__ void apply(const double* __restrict__ M, const double (&v)[NL], double (&o)[NO]) {
#pragma unroll
  for (int a = 0; a < NO; ++a) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < NL; ++l) s = fma(M[a * NL + l], v[l], s);
    o[a] = s;
  }
}
template <int NO, int NL>
__device__ __forceinline__ void apply_acc(const double* __restrict__ M, const double (&v)[NL], double (&o)[NO]) {
#pragma unroll
  for (int a = 0; a < NO; ++a) {
    double s = o[a];
#pragma unroll
    for (int l = 0; l < NL; ++l) s = fma(M[a * NL + l], v[l], s);
    o[a] = s;
  }
}

// out[a*so] = sum_l M[a][l] v[l]                      (one output at a time: low register pressure)
template <int NO, int NL>
__device__ __forceinline__ void apply_store(const double* __restrict__ M, const double (&v)[NL], double* __restrict__ po, int so) {
#pragma unroll
  for (int a = 0; a < NO; ++a) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < NL; ++l) s = fma(M[a * NL + l], v[l], s);
    po[a * so] = s;
  }
}
// out[a*so] = sum_l M1[a][l] v1[l] + M2[a][l] v2[l]
template <int NO, int NL>
__device__ __forceinline__ void apply2_store(const double* __restrict__ M1, const double (&v1)[NL], const double* __restrict__ M2,
                                             const double (&v2)[NL], double* __restrict__ po, long long so) {
#pragma unroll
  for (int a = 0; a < NO; ++a) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < NL; ++l) s = fma(M1[a * NL + l], v1[l], s);
#pragma unroll
    for (int l = 0; l < NL; ++l) s = fma(M2[a * NL + l], v2[l], s);
    po[a * so] = s;
  }
}

// ------------------------------------------------------------------------------------------------ gradt
// MODE 0: w = D^T p.  MODE 1: CG direction with a pointwise preconditioner, p = dinvE*r + beta*pdir.  MODE 2: CG direction with the
// three-level preconditioner of pmg.cu in its fused form: `p` holds the element-block part zloc of z = M^-1 r (k_pcg_fused); the Q1
// vertex-mesh part (trilinear interpolation of the element's 8 corner values) and the aggregate value are added here; in this mode the
// argument `xv` is the per-element table xc[nel][9] (k_pm_corner_values) and `x2` the table of hat-function values at the GL points.
template <int N, int MODE>
__global__ void __launch_bounds__(PK_TPB, 3)      // MODE 0/1: 36-40 registers (4 CTAs per SM fit); MODE 2: 54 registers, 3 CTAs per SM
k_gradt3(const double* __restrict__ p, double* __restrict__ w, const double* __restrict__ RW2,
         const double* __restrict__ dinvE, double* __restrict__ pdir, const CGState* __restrict__ cgs, long long n,
         long long n2, const double* __restrict__ xv, const int* __restrict__ vid, const double* __restrict__ x2,
         const int* __restrict__ agg, double* __restrict__ xsol) {
  constexpr int N2 = N - 2, TPB = PK_TPB;
  constexpr int NP1 = N * N * N, NP2 = N2 * N2 * N2;
  using S2 = Shp<N2, N2, N2>;
  using SA = Shp<N2, N2, N>;      // after the r stage
  using SB = Shp<N2, N, N>;       // after the s stage
  using C0 = ColIn<0, N2, N2, N2>;
  using C1 = ColIn<1, N2, N2, N>;
  using C2 = ColIn<2, N2, N, N>;
  __shared__ double sa[9][SA::size];
  __shared__ double sb[6][SB::size];
  // products RW2[i][c]*p, one array per (i,c); consumed by the r stage before the s stage overwrites sb
  double* sq = &sb[0][0];
  static_assert(9 * S2::size <= 6 * SB::size, "sq alias");
  const int tid = threadIdx.x;
  if (MODE >= 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * NP2;
  const long long e1 = (long long)blockIdx.x * NP1;
  static_assert(NP2 <= TPB, "one mesh-2 point per thread");
  if (MODE == 2) {
    // xc = the element's 8 corner values of the vertex level + its aggregate value (k_pm_corner_values), one contiguous record per
    // element.  Otherwise idle threads fetch it (and the hat-function table) into shared memory while the point threads have their
    // 12 streaming loads in flight; one barrier, then the interpolation reads shared memory.  History (r2, cfg 5): gathering
    // xv[vid[..]] here (two dependent, L2-evicted loads before the barrier) 0.183 ms; warp-uniform global loads of xc without a
    // barrier 0.186 ms (24 loads per thread at a 56-register cap: the compiler split them into two dependent rounds); MODE 1: 0.150 ms.
    __shared__ double sxc[11], sh1[8];
    if (tid >= TPB - 9) sxc[tid - (TPB - 9)] = xv[(long long)blockIdx.x * 9 + (tid - (TPB - 9))];
    else if (tid == TPB - 10) sxc[9] = cgs->beta;        // the CG scalars travel the same way: no global load behind the barrier
    else if (tid == TPB - 11) sxc[10] = cgs->alpha;
    else if (tid >= TPB - 32 && tid < TPB - 32 + N2) sh1[tid - (TPB - 32)] = x2[tid - (TPB - 32)];      // x2 = hat-function table
    double zl = 0.0, pd = 0.0, xo = 0.0, m9[9];
    if (tid < NP2) {
      zl = p[e2 + tid];
      pd = pdir[e2 + tid];
      xo = xsol[e2 + tid];
#pragma unroll
      for (int g = 0; g < 9; ++g) m9[g] = RW2[(long long)g * n2 + e2 + tid];   // coalesced
    }
    __syncthreads();
    if (tid < NP2) {
      const int q = tid;
      const int i0 = q % N2, i1 = (q / N2) % N2, i2 = q / (N2 * N2);
      const double a0 = sh1[i0], a1 = sh1[i1], a2 = sh1[i2];
      const double c0 = fma(sxc[1] - sxc[0], a0, sxc[0]), c1 = fma(sxc[3] - sxc[2], a0, sxc[2]);
      const double d0 = fma(sxc[5] - sxc[4], a0, sxc[4]), d1 = fma(sxc[7] - sxc[6], a0, sxc[6]);
      const double q0 = fma(c1 - c0, a1, c0), q1 = fma(d1 - d0, a1, d0);
      const double z = (zl + fma(q1 - q0, a2, q0)) + sxc[8];
      const double v = fma(sxc[9], pd, z);
      pdir[e2 + q] = v;
      // the solution update of the PREVIOUS iteration rides along here (pd = p_k is in a register anyway): x_{k+1} = x_k + alpha_k p_k;
      // the update of the last iteration is applied by k_pcg_xfix after the loop (this kernel is skipped once converged)
      xsol[e2 + q] = fma(sxc[10], pd, xo);
      const int o = S2::lin(q);
#pragma unroll
      for (int g = 0; g < 9; ++g) sq[g * S2::size + o] = m9[g] * v;
    }
  } else if (tid < NP2) {
    const int q = tid;
    double v;
    if (MODE == 1) {
      v = dinvE[e2 + q] * p[e2 + q] + cgs->beta * pdir[e2 + q];
      pdir[e2 + q] = v;
    } else {
      v = p[e2 + q];
    }
    const int o = S2::lin(q);
#pragma unroll
    for (int g = 0; g < 9; ++g) sq[g * S2::size + o] = RW2[(long long)g * n2 + e2 + q] * v;   // coalesced
  }
  __syncthreads();
  // ---- r stage: sa[i*3+c] = (i==0 ? D12^T : J12^T) applied along r to (RW2[i][c] * p)
  static_assert(9 * C0::ncol <= TPB, "one task per thread");
  if (tid < 9 * C0::ncol) {
    const int t = tid;
    const int combo = t / C0::ncol, col = t - combo * C0::ncol;
    const double* pin = sq + combo * S2::size + C0::base(col);
    double v[N2];
#pragma unroll
    for (int l = 0; l < N2; ++l) v[l] = pin[l];
    double* po = sa[combo] + col * SA::PI;
    if (combo < 3) apply_store<N, N2>(cm.D12t, v, po, 1);
    else apply_store<N, N2>(cm.J12t, v, po, 1);
  }
  __syncthreads();
  // ---- s stage: sb[c*2+0] = J12^T sa[0*3+c] + D12^T sa[1*3+c] ; sb[c*2+1] = J12^T sa[2*3+c]
  static_assert(6 * C1::ncol <= TPB, "one task per thread");
  if (tid < 6 * C1::ncol) {
    const int t = tid;
    const int grp = t / C1::ncol, col = t - grp * C1::ncol;
    const int which = grp / 3, c = grp - which * 3;
    const int bi = C1::base(col);
    const int bo = (col / N) * N * SB::PI + (col % N);
    double v[N2], v2[N2];
    double* po = sb[c * 2 + which] + bo;
    if (which == 0) {
#pragma unroll
      for (int l = 0; l < N2; ++l) { v[l] = sa[c][bi + l * C1::stride]; v2[l] = sa[3 + c][bi + l * C1::stride]; }
      apply2_store<N, N2>(cm.J12t, v, cm.D12t, v2, po, SB::PI);
    } else {
#pragma unroll
      for (int l = 0; l < N2; ++l) v[l] = sa[6 + c][bi + l * C1::stride];
      apply_store<N, N2>(cm.J12t, v, po, SB::PI);
    }
  }
  __syncthreads();
  // ---- t stage: w_c = J12^T sb[c*2] + D12^T sb[c*2+1], stored straight to global memory (coalesced in (j,i))
  //      (r2, measured: sharing every column between two threads -- 384 equal tasks instead of 192 -- balanced the stage but doubled
  //      its shared-memory loads; the kernel is LSU-bound and got 9 % slower: 0.173 -> 0.189 ms.  Not kept.)
  static_assert(3 * C2::ncol <= TPB, "one task per thread");
  if (tid < 3 * C2::ncol) {
    const int t = tid;
    const int c = t / C2::ncol, col = t - c * C2::ncol;
    const int bi = C2::base(col);
    double v[N2], v2[N2];
#pragma unroll
    for (int l = 0; l < N2; ++l) { v[l] = sb[c * 2][bi + l * C2::stride]; v2[l] = sb[c * 2 + 1][bi + l * C2::stride]; }
    apply2_store<N, N2>(cm.J12t, v, cm.D12t, v2, w + (long long)c * n + e1 + col, N * N);
  }
}

// ------------------------------------------------------------------------------------------------ div
// MODE 0: q = sign * sum_c D_c (scale_c u_c)      MODE 1: CG (Ep, rho = sum pdir*Ep)
template <int N, int MODE>
__global__ void __launch_bounds__(PK_TPB, 3)
k_div3(const double* __restrict__ u, const double* __restrict__ scale0, const double* __restrict__ scale1,
       const double* __restrict__ scale2, double* __restrict__ qout, const double* __restrict__ RW2,
       const double* __restrict__ pdir, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter,
       double* __restrict__ red_out, int finalize, long long n, long long n2, double sign) {
  constexpr int N2 = N - 2, TPB = PK_TPB;
  constexpr int NP1 = N * N * N, NP2 = N2 * N2 * N2;
  using S1 = Shp<N, N, N>;
  using SA = Shp<N2, N, N>;      // after the t stage
  using SB = Shp<N2, N2, N>;     // after the s stage
  using C2 = ColIn<2, N, N, N>;
  using C1 = ColIn<1, N2, N, N>;
  using C0 = ColIn<0, N2, N2, N>;
  // su (3 x S1) is dead after the t stage and sbuf (9 x SB) is first written in the s stage: they share storage.
  // Dynamic shared memory: sus | sa (reused for the 9 partial-product arrays) | srw (9 metric arrays + pdir, filled
  // asynchronously with cp.async at kernel entry so that their DRAM latency hides behind the three tensor stages).
  constexpr int SU = 3 * S1::size, SBB = 9 * SB::size;
  constexpr int SUS = (SU > SBB) ? SU : SBB;
  extern __shared__ double dsm[];
  double* sus = dsm;
  double (*sa)[SA::size] = reinterpret_cast<double (*)[SA::size]>(dsm + SUS);
  double* srw = dsm + SUS + 6 * SA::size;          // [10][NP2]
  __shared__ double sred[32];
  double (*spart)[NP2] = reinterpret_cast<double (*)[NP2]>(&sa[0][0]);
  static_assert(9 * NP2 <= 6 * SA::size, "spart alias");
  double* su = sus;
  double* sbuf = sus;
  const int tid = threadIdx.x;
  if (MODE == 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * NP2;
  const long long e1 = (long long)blockIdx.x * NP1;
  static_assert(NP2 <= TPB, "one mesh-2 point per thread");
  if (tid < NP2) {                                  // LDGSTS: global -> shared without staging through registers
#pragma unroll
    for (int g = 0; g < 9; ++g) __pipeline_memcpy_async(&srw[g * NP2 + tid], &RW2[(long long)g * n2 + e2 + tid], sizeof(double));
    if (MODE == 1) __pipeline_memcpy_async(&srw[9 * NP2 + tid], &pdir[e2 + tid], sizeof(double));
  }
  __pipeline_commit();
  // ---- t stage: aJ = J12 u, aD = D12 u along t (same input column, two outputs).  The column is read straight from global
  //      memory: lanes run over (j,i), so each of the N loads is coalesced.
  static_assert(3 * C2::ncol <= TPB, "one task per thread");
  if (tid < 3 * C2::ncol) {
    const int t = tid;
    const int c = t / C2::ncol, col = t - c * C2::ncol;
    double v[N];
    {
      const double* sc = (c == 0 || !scale1) ? scale0 : (c == 1 ? scale1 : scale2);
      const double* pu = u + (long long)c * n + e1 + col;
      if (sc) {
        const double* ps = sc + e1 + col;
#pragma unroll
        for (int l = 0; l < N; ++l) v[l] = pu[l * N * N] * ps[l * N * N];
      } else {
#pragma unroll
        for (int l = 0; l < N; ++l) v[l] = pu[l * N * N];
      }
    }
    // output (N2,N,N): same (j,i) base, plane stride N*PI
    apply_store<N2, N>(cm.J12, v, sa[c * 2] + C2::base(col), C2::stride);
    apply_store<N2, N>(cm.D12, v, sa[c * 2 + 1] + C2::base(col), C2::stride);
  }
  __syncthreads();
  // ---- s stage: from aJ: bJJ = J12 aJ, bDJ = D12 aJ ; from aD: bJD = J12 aD
  static_assert(6 * C1::ncol <= TPB, "one task per thread");
  if (tid < 6 * C1::ncol) {
    const int t = tid;
    const int grp = t / C1::ncol, col = t - grp * C1::ncol;
    const int which = grp / 3, c = grp - which * 3;
    const int bi = C1::base(col);
    const int bo = (col / N) * N2 * SB::PI + (col % N);
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = sa[c * 2 + which][bi + l * C1::stride];
    apply_store<N2, N>(cm.J12, v, sbuf + (c * 3 + (which == 0 ? 0 : 2)) * SB::size + bo, SB::PI);
    if (which == 0) apply_store<N2, N>(cm.D12, v, sbuf + (c * 3 + 1) * SB::size + bo, SB::PI);
  }
  __syncthreads();
  // ---- r stage, one (direction, component) pair per task:
  //      part[dir*3+c] = (dir==0 ? D12 : J12) applied along r to {bJJ, bDJ, bJD}[dir]; the metric multiply and the sum
  //      over the 9 pairs happen in the final, coalesced pass
  static_assert(9 * C0::ncol <= TPB, "one task per thread");
  if (tid < 9 * C0::ncol) {
    const int grp = tid / C0::ncol, col = tid - grp * C0::ncol;
    const int dir = grp / 3, c = grp - dir * 3;
    const double* b0 = sbuf + (c * 3 + dir) * SB::size + C0::base(col);
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = b0[l];
    double* po = &spart[grp][col * N2];
    if (dir == 0) apply_store<N2, N>(cm.D12, v, po, 1);
    else apply_store<N2, N>(cm.J12, v, po, 1);
  }
  __pipeline_wait_prior(0);
  __syncthreads();
  double rho[1] = {0.0};
  for (int q = tid; q < NP2; q += TPB) {
    double acc = 0.0;
#pragma unroll
    for (int g = 0; g < 9; ++g) acc = fma(srw[g * NP2 + q], spart[g][q], acc);
    const double v = sign * acc;
    qout[e2 + q] = v;
    if (MODE == 1) rho[0] += srw[9 * NP2 + q] * v;
  }
  if (MODE == 1) {
    if (grid_sum_finish<1>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
      cgs->rho = red_out[0];
      cgs->alpha = cgs->rtz1 / red_out[0];
    }
  }
}

// ------------------------------------------------------------------------------------------------ TMA (bulk async copy) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk tensor-memory-accelerator copy global -> shared; completion (bytes) is signalled on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------------ persistent, TMA-pipelined pressure-CG kernels
// One CTA per SM slot (2 per SM) walks over elements e = blockIdx.x, += gridDim.x.  All global -> shared traffic of an
// element (metrics, CG vectors, the three w components) is moved by bulk TMA copies into a double-buffered stage while
// the previous element is being contracted, so DRAM requests are in flight for the whole life of the CTA instead of
// only during each CTA's load phase (r1b ncu: k_div3 27 % DRAM, barrier + long-scoreboard stalls).
// ------------------------------------------------------------------------------------------------ k_div3q: single staging buffer, early refill
// r1's k_div3p (double-buffered stage, 111 KB per CTA, 2 CTAs/SM): 59 % of the HBM roofline, 15 % of the stall samples in the mbarrier
// wait -- not enough bytes in flight; replaced by this kernel in r2 (0.186 -> 0.176 ms).  Here the element's inputs live in ONE staging buffer whose two halves are refilled as
// soon as they have been consumed (the scheme of k_axhelm3p):
//   half A (w x3, mask*binv)   consumed by the t stage      -> element e+1 requested right after the t-stage barrier
//   half B (9 metric arrays)   consumed by the final stage  -> element e+1 requested right after the final stage
// pdir is read straight from global memory into a register at the start of the element.  74.5 KB per CTA => THREE CTAs per SM.
template <int N>
struct DivQ {
  static constexpr int N2 = N - 2, NP1 = N * N * N, NP2 = N2 * N2 * N2;
  using SA = Shp<N2, N, N>;
  using SB = Shp<N2, N2, N>;
  static constexpr int halfA = 4 * NP1, halfB = 9 * NP2;
  static constexpr size_t smem = sizeof(double) * (halfA + halfB + 6 * SA::size + 9 * SB::size) + 2 * sizeof(uint64_t);
};
struct DivQArgs {
  const double* u; const double* mbinv; const double* RW2; const double* pdir; double* qout;
  long long n, n2;
};
template <int N>
__device__ __forceinline__ void div3q_issue_a(const DivQArgs* A, double* stg, uint64_t* bar, int e) {
  constexpr int NP1 = N * N * N;
  const long long e1 = (long long)e * NP1;
  mbar_expect_tx(bar, 4 * NP1 * (uint32_t)sizeof(double));
#pragma unroll
  for (int c = 0; c < 3; ++c) tma_bulk_g2s(stg + c * NP1, A->u + (long long)c * A->n + e1, NP1 * sizeof(double), bar);
  tma_bulk_g2s(stg + 3 * NP1, A->mbinv + e1, NP1 * sizeof(double), bar);
}
template <int N>
__device__ __forceinline__ void div3q_issue_b(const DivQArgs* A, double* stg, uint64_t* bar, int e) {
  constexpr int NP1 = N * N * N, NP2 = (N - 2) * (N - 2) * (N - 2);
  const long long e2 = (long long)e * NP2;
  mbar_expect_tx(bar, 9 * NP2 * (uint32_t)sizeof(double));
#pragma unroll
  for (int g = 0; g < 9; ++g) tma_bulk_g2s(stg + 4 * NP1 + g * NP2, A->RW2 + (long long)g * A->n2 + e2, NP2 * sizeof(double), bar);
}

template <int N>
__device__ __noinline__ double div3q_element(const DivQArgs* __restrict__ A, double* __restrict__ stg, double* __restrict__ sa,
                                             double* __restrict__ sbuf, uint64_t* bar, int e, int e_next, uint32_t parity, int tid) {
  using P = DivQ<N>;
  constexpr int N2 = P::N2, NP1 = P::NP1, NP2 = P::NP2;
  using SA = typename P::SA;
  using SB = typename P::SB;
  using C2 = ColIn<2, N, N, N>;
  using C1 = ColIn<1, N2, N, N>;
  using C0 = ColIn<0, N2, N2, N>;
  double* spart = sa;
  const double* inmb = stg + 3 * NP1;
  const double* inrw = stg + 4 * NP1;
  const double pd = (tid < NP2) ? A->pdir[(long long)e * NP2 + tid] : 0.0;      // needed in the final stage only: latency hidden
  mbar_wait(&bar[0], parity);
  if (tid < 3 * C2::ncol) {
    const int c = tid / C2::ncol, col = tid - c * C2::ncol;
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = stg[c * NP1 + l * N * N + col] * inmb[l * N * N + col];
    apply_store<N2, N>(cm.J12, v, sa + (c * 2) * SA::size + C2::base(col), C2::stride);
    apply_store<N2, N>(cm.D12, v, sa + (c * 2 + 1) * SA::size + C2::base(col), C2::stride);
  }
  __syncthreads();                                   // half A consumed
  if (tid == 0 && e_next >= 0) div3q_issue_a<N>(A, stg, &bar[0], e_next);
  if (tid < 6 * C1::ncol) {
    const int grp = tid / C1::ncol, col = tid - grp * C1::ncol;
    const int which = grp / 3, c = grp - which * 3;
    const int bi = C1::base(col);
    const int bo = (col / N) * N2 * SB::PI + (col % N);
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = sa[(c * 2 + which) * SA::size + bi + l * C1::stride];
    apply_store<N2, N>(cm.J12, v, sbuf + (c * 3 + (which == 0 ? 0 : 2)) * SB::size + bo, SB::PI);
    if (which == 0) apply_store<N2, N>(cm.D12, v, sbuf + (c * 3 + 1) * SB::size + bo, SB::PI);
  }
  __syncthreads();
  if (tid < 9 * C0::ncol) {
    const int grp = tid / C0::ncol, col = tid - grp * C0::ncol;
    const int dir = grp / 3, c = grp - dir * 3;
    const double* b0 = sbuf + (c * 3 + dir) * SB::size + C0::base(col);
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = b0[l];
    double* po = spart + grp * NP2 + col * N2;
    if (dir == 0) apply_store<N2, N>(cm.D12, v, po, 1);
    else apply_store<N2, N>(cm.J12, v, po, 1);
  }
  mbar_wait(&bar[1], parity);
  __syncthreads();
  double rho = 0.0;
  if (tid < NP2) {
    const int q = tid;
    double acc = 0.0;
#pragma unroll
    for (int g = 0; g < 9; ++g) acc = fma(inrw[g * NP2 + q], spart[g * NP2 + q], acc);
    A->qout[(long long)e * NP2 + q] = acc;
    rho = pd * acc;
  }
  __syncthreads();                                   // half B consumed (and spart free for the next element's t stage)
  if (tid == 0 && e_next >= 0) div3q_issue_b<N>(A, stg, &bar[1], e_next);
  return rho;
}

template <int N>
__global__ void __launch_bounds__(PK_TPB, 3)
k_div3q(DivQArgs args, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter, double* __restrict__ red_out,
        int finalize, int nel) {
  using P = DivQ<N>;
  using SA = typename P::SA;
  using SB = typename P::SB;
  extern __shared__ __align__(128) double dsm[];
  double* stg = dsm;                                   // [halfA | halfB]
  double* sa = dsm + P::halfA + P::halfB;              // [6][SA::size], reused for the 9 partial arrays
  double* sbuf = sa + 6 * SA::size;                    // [9][SB::size]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbuf + 9 * SB::size);
  __shared__ double sred[32];
  __shared__ DivQArgs sargs;
  static_assert(9 * P::NP2 <= 6 * SA::size, "spart alias");
  const int tid = threadIdx.x;
  if (cgs->done) return;
  if (tid == 0) {
    sargs = args;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < nel) {
    div3q_issue_a<N>(&sargs, stg, &bar[0], blockIdx.x);
    div3q_issue_b<N>(&sargs, stg, &bar[1], blockIdx.x);
  }
  double rho[1] = {0.0};
  int it = 0;
  for (int e = blockIdx.x; e < nel; e += gridDim.x, ++it) {
    const int en = (e + (int)gridDim.x < nel) ? e + (int)gridDim.x : -1;
    rho[0] += div3q_element<N>(&sargs, stg, sa, sbuf, bar, e, en, (uint32_t)(it & 1), tid);
  }
  if (grid_sum_finish<1>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
    cgs->rho = red_out[0];
    cgs->alpha = cgs->rtz1 / red_out[0];
  }
}

// ------------------------------------------------------------------------------------------------ axhelm
// w_f = (h1 A + h2 B) u_f for up to 3 fields at once [UPSTREAM hmholtz.f axhelm = local_grad3 -> G -> local_grad3_t]:
// every (field, direction) pair is one column task per thread in both tensor stages; the stiffness factors are read
// once per point for all fields.  MODE 0: w = H u.  MODE 1: r = b + r - H u (cresvipp).  MODE 2: Helmholtz-CG
// iteration head: p = dinv*r + beta*p (stored), w = H p, rho_f = sum p*w (per-component CG states).
template <int N>
struct AxCfg {
  static constexpr int NP1 = N * N * N;
  static constexpr int NT = 9 * N * N;
  static constexpr int TPB = (((NT > NP1) ? NT : NP1) + 31) / 32 * 32;
};

// PF: the six stiffness factors and bm1 of the element are fetched with cp.async into shared memory at kernel entry and
// consumed two barriers later (r1d ncu: 8.6 long-scoreboard stall warps per issue, most of them at the G loads that sat
// right behind a barrier).
template <int N, int MODE, bool PF>
__global__ void __launch_bounds__(AxCfg<N>::TPB, 2)
k_axhelm3(const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ b, const double* __restrict__ G,
          const double* __restrict__ bm1, const double* __restrict__ dinv, double* __restrict__ pdir, CGState* __restrict__ cgs,
          double* __restrict__ part, unsigned* counter, double* __restrict__ red_out, int finalize, int nfields, long long n,
          double h1, double h2) {
  constexpr int NP1 = N * N * N, NC = N * N;
  using S1 = Shp<N, N, N>;
  constexpr int PI = S1::PI;
  extern __shared__ double dsm[];
  double* su = dsm;                     // [3][S1::size]
  double* sr = dsm + 3 * S1::size;      // [9][S1::size]  (field*3 + direction)
  __shared__ double sred[3 * 32];
  const int tid = threadIdx.x;
  const long long e0 = (long long)blockIdx.x * NP1;
  bool act[3];
  double beta[3];
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    act[f] = f < nfields && !(MODE == 2 && cgs[f].done);
    beta[f] = (MODE == 2 && f < nfields) ? cgs[f].beta : 0.0;
  }
  if (MODE == 2 && !(act[0] || act[1] || act[2])) return;
  double* sG = dsm + 12 * S1::size;     // [7][NP1]: g1..g6, bm1 (PF only)
  if (PF) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sG);
    for (int t = tid; t < 7 * NP1 / 2; t += AxCfg<N>::TPB) {
      const int a = t / (NP1 / 2), w2 = t - a * (NP1 / 2);
      const double* src = (a < 6 ? G + (long long)a * n : bm1) + e0 + 2 * w2;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (unsigned)((a * NP1 + 2 * w2) * 8)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  double uo[3] = {0.0, 0.0, 0.0};
  if (tid < NP1) {
    const int o = S1::lin(tid);
    const double di = (MODE == 2) ? dinv[e0 + tid] : 0.0;
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (act[f]) {
        const long long gi = (long long)f * n + e0 + tid;
        double v;
        if (MODE == 2) {
          v = di * u[gi] + beta[f] * pdir[gi];          // u = r
          pdir[gi] = v;
        } else {
          v = u[gi];
        }
        uo[f] = v;
        su[f * S1::size + o] = v;
      }
  }
  __syncthreads();
  // ---- derivatives: sr[f*3+dir] = D applied along dir to su[f]
  const int tf = tid / (3 * NC), tdir = (tid / NC) % 3, tcol = tid % NC;
  const bool task = tid < 9 * NC && tf < nfields && act[tf < 3 ? tf : 0];
  const int cbase = (tdir == 0) ? tcol * PI : ((tdir == 1) ? (tcol / N) * N * PI + (tcol % N) : (tcol / N) * PI + (tcol % N));
  const int cstr = (tdir == 0) ? 1 : ((tdir == 1) ? PI : N * PI);
  if (task) {
    const double* pin = su + tf * S1::size + cbase;
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = pin[l * cstr];
    apply_store<N, N>(cm.D, v, sr + (tf * 3 + tdir) * S1::size + cbase, cstr);
  }
  if (PF) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- geometric factors, all fields of a point at once
  if (tid < NP1) {
    const int o = S1::lin(tid);
    double g[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) g[q] = PF ? sG[q * NP1 + tid] : G[(long long)q * n + e0 + tid];
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (act[f]) {
        double* p0 = sr + (f * 3) * S1::size + o;
        const double d0 = p0[0], d1 = p0[S1::size], d2 = p0[2 * S1::size];
        p0[0] = g[0] * d0 + g[3] * d1 + g[4] * d2;
        p0[S1::size] = g[3] * d0 + g[1] * d1 + g[5] * d2;
        p0[2 * S1::size] = g[4] * d0 + g[5] * d1 + g[2] * d2;
      }
  }
  __syncthreads();
  // ---- transposed derivatives, in place (each thread owns its column)
  if (task) {
    double* pc = sr + (tf * 3 + tdir) * S1::size + cbase;
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = pc[l * cstr];
    apply_store<N, N>(cm.Dt, v, pc, cstr);
  }
  __syncthreads();
  double rho[3] = {0.0, 0.0, 0.0};
  if (tid < NP1) {
    const int o = S1::lin(tid);
    const double bm = PF ? sG[6 * NP1 + tid] : bm1[e0 + tid];
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (act[f]) {
        const double* p0 = sr + (f * 3) * S1::size + o;
        const double hv = h1 * ((p0[0] + p0[S1::size]) + p0[2 * S1::size]) + h2 * bm * uo[f];
        const long long gi = (long long)f * n + e0 + tid;
        if (MODE == 1) {
          w[gi] = b[gi] + w[gi] - hv;
        } else {
          w[gi] = hv;
          if (MODE == 2) rho[f] = uo[f] * hv;
        }
      }
  }
  if (MODE == 2) {
    if (grid_sum_finish<3>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
      for (int f = 0; f < nfields; ++f)
        if (!cgs[f].done) {
          cgs[f].rho = red_out[f];
          cgs[f].alpha = cgs[f].rtz1 / red_out[f];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ persistent, TMA-pipelined axhelm (Helmholtz-CG head)
// r1d ncu on k_axhelm3<8,2>: 45 % of the HBM roofline, stalls = load latency at the kernel head (one element per CTA, 2 CTAs/SM, nothing
// in flight while the four tensor phases run).  Here one CTA per SM slot (2 per SM) walks over elements; the 14 input arrays of an
// element live in ONE staging buffer whose two halves are refilled by bulk TMA copies as soon as they have been consumed:
//   head half (r x3, pdir x3, dinv) is consumed by the first phase  -> element e+1's head is requested right after the first barrier
//   G half (g1..g6, bm1) is consumed by the geometric-factor phase  -> element e+1's factors are requested right after that barrier
// so the loads of element e+1 are in flight during three of the four phases of element e.  Shared memory: 14 + 12*1.125 element
// arrays = 112 KB per CTA, two CTAs per SM.
struct AxPArgs {
  const double* r; const double* dinv; const double* G; const double* bm1;
  double* pdir; double* w;
  long long n;
  double h1, h2;
  int perm;                 // 1: w is written in the surface-first element layout (read back by k_hcg_update through the same map)
};
template <int N>
struct AxP {
  static constexpr int NP1 = N * N * N;
  using S1 = Shp<N, N, N>;
  static constexpr int TPB = AxCfg<N>::TPB;
  static constexpr int stage = 14 * NP1;
  static constexpr size_t smem = sizeof(double) * (stage + 12 * S1::size) + 2 * sizeof(uint64_t) + sizeof(AxPArgs);
};
struct Rho3 { double a, b, c; };

template <int N>
__device__ __forceinline__ void ax3p_issue_head(const AxPArgs* A, double* stg, uint64_t* bar, int e) {
  constexpr int NP1 = N * N * N;
  const long long e0 = (long long)e * NP1;
  mbar_expect_tx(bar, 7 * NP1 * (uint32_t)sizeof(double));
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    tma_bulk_g2s(stg + f * NP1, A->r + (long long)f * A->n + e0, NP1 * sizeof(double), bar);
    tma_bulk_g2s(stg + (3 + f) * NP1, A->pdir + (long long)f * A->n + e0, NP1 * sizeof(double), bar);
  }
  tma_bulk_g2s(stg + 6 * NP1, A->dinv + e0, NP1 * sizeof(double), bar);
}
template <int N>
__device__ __forceinline__ void ax3p_issue_g(const AxPArgs* A, double* stg, uint64_t* bar, int e) {
  constexpr int NP1 = N * N * N;
  const long long e0 = (long long)e * NP1;
  mbar_expect_tx(bar, 7 * NP1 * (uint32_t)sizeof(double));
#pragma unroll
  for (int q = 0; q < 6; ++q) tma_bulk_g2s(stg + (7 + q) * NP1, A->G + (long long)q * A->n + e0, NP1 * sizeof(double), bar);
  tma_bulk_g2s(stg + 13 * NP1, A->bm1 + e0, NP1 * sizeof(double), bar);
}

// One element (not inlined: inside the element loop nvcc would hoist the constant-bank operator matrices into registers, see
// div3q_element).  actmask: bit f = component f still iterating.
template <int N>
__device__ __noinline__ Rho3 ax3p_element(const AxPArgs* __restrict__ A, double* __restrict__ stg, double* __restrict__ su,
                                          double* __restrict__ sr, uint64_t* bar, int e, int e_next, uint32_t parity, int actmask,
                                          double b0, double b1, double b2, int tid) {
  constexpr int NP1 = N * N * N, NC = N * N;
  using S1 = Shp<N, N, N>;
  constexpr int PI = S1::PI;
  const long long e0 = (long long)e * NP1;
  const double beta[3] = {b0, b1, b2};
  double uo[3] = {0.0, 0.0, 0.0};
  mbar_wait(&bar[0], parity);
  if (tid < NP1) {
    const int o = S1::lin(tid);
    const double di = stg[6 * NP1 + tid];
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (actmask & (1 << f)) {
        const double v = di * stg[f * NP1 + tid] + beta[f] * stg[(3 + f) * NP1 + tid];
        A->pdir[(long long)f * A->n + e0 + tid] = v;
        uo[f] = v;
        su[f * S1::size + o] = v;
      }
  }
  __syncthreads();                                   // head half consumed
  if (tid == 0 && e_next >= 0) ax3p_issue_head<N>(A, stg, &bar[0], e_next);
  const int tf = tid / (3 * NC), tdir = (tid / NC) % 3, tcol = tid % NC;
  const bool task = tid < 9 * NC && ((actmask >> (tf < 3 ? tf : 0)) & 1);
  const int cbase = (tdir == 0) ? tcol * PI : ((tdir == 1) ? (tcol / N) * N * PI + (tcol % N) : (tcol / N) * PI + (tcol % N));
  const int cstr = (tdir == 0) ? 1 : ((tdir == 1) ? PI : N * PI);
  if (task) {
    const double* pin = su + tf * S1::size + cbase;
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = pin[l * cstr];
    apply_store<N, N>(cm.D, v, sr + (tf * 3 + tdir) * S1::size + cbase, cstr);
  }
  mbar_wait(&bar[1], parity);
  __syncthreads();
  double bm = 0.0;
  if (tid < NP1) {
    const int o = S1::lin(tid);
    double g[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) g[q] = stg[(7 + q) * NP1 + tid];
    bm = stg[13 * NP1 + tid];
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (actmask & (1 << f)) {
        double* p0 = sr + (f * 3) * S1::size + o;
        const double d0 = p0[0], d1 = p0[S1::size], d2 = p0[2 * S1::size];
        p0[0] = g[0] * d0 + g[3] * d1 + g[4] * d2;
        p0[S1::size] = g[3] * d0 + g[1] * d1 + g[5] * d2;
        p0[2 * S1::size] = g[4] * d0 + g[5] * d1 + g[2] * d2;
      }
  }
  __syncthreads();                                   // factor half consumed
  if (tid == 0 && e_next >= 0) ax3p_issue_g<N>(A, stg, &bar[1], e_next);
  if (task) {
    double* pc = sr + (tf * 3 + tdir) * S1::size + cbase;
    double v[N];
#pragma unroll
    for (int l = 0; l < N; ++l) v[l] = pc[l * cstr];
    apply_store<N, N>(cm.Dt, v, pc, cstr);
  }
  __syncthreads();
  Rho3 rho = {0.0, 0.0, 0.0};
  if (tid < NP1) {
    const int o = S1::lin(tid);
    const int wpos = A->perm ? SurfFirst<N>::pos_lin(tid) : tid;
    double rr[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int f = 0; f < 3; ++f)
      if (actmask & (1 << f)) {
        const double* p0 = sr + (f * 3) * S1::size + o;
        const double hv = A->h1 * ((p0[0] + p0[S1::size]) + p0[2 * S1::size]) + A->h2 * bm * uo[f];
        A->w[(long long)f * A->n + e0 + wpos] = hv;
        rr[f] = uo[f] * hv;
      }
    rho.a = rr[0]; rho.b = rr[1]; rho.c = rr[2];
  }
  return rho;
}

template <int N>
__global__ void __launch_bounds__(AxCfg<N>::TPB, 2)
k_axhelm3p(AxPArgs args, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter, double* __restrict__ red_out,
           int finalize, int nel) {
  using P = AxP<N>;
  using S1 = typename P::S1;
  extern __shared__ __align__(128) double dsm[];
  double* stg = dsm;                                 // [14][NP1]
  double* su = dsm + P::stage;                       // [3][S1::size]
  double* sr = su + 3 * S1::size;                    // [9][S1::size]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sr + 9 * S1::size);
  AxPArgs* sargs = reinterpret_cast<AxPArgs*>(bar + 2);
  __shared__ double sred[3 * 32];
  const int tid = threadIdx.x;
  int actmask = 0;
#pragma unroll
  for (int f = 0; f < 3; ++f) actmask |= cgs[f].done ? 0 : (1 << f);
  if (!actmask) return;
  const double b0 = cgs[0].beta, b1 = cgs[1].beta, b2 = cgs[2].beta;
  if (tid == 0) {
    *sargs = args;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0 && (int)blockIdx.x < nel) {
    ax3p_issue_head<N>(sargs, stg, &bar[0], blockIdx.x);
    ax3p_issue_g<N>(sargs, stg, &bar[1], blockIdx.x);
  }
  double rho[3] = {0.0, 0.0, 0.0};
  int it = 0;
  for (int e = blockIdx.x; e < nel; e += gridDim.x, ++it) {
    const int en = (e + (int)gridDim.x < nel) ? e + (int)gridDim.x : -1;
    const Rho3 q = ax3p_element<N>(sargs, stg, su, sr, bar, e, en, (uint32_t)(it & 1), actmask, b0, b1, b2, tid);
    rho[0] += q.a; rho[1] += q.b; rho[2] += q.c;
  }
  if (grid_sum_finish<3>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
    for (int f = 0; f < 3; ++f)
      if (!cgs[f].done) {
        cgs[f].rho = red_out[f];
        cgs[f].alpha = cgs[f].rtz1 / red_out[f];
      }
  }
}

}  // namespace

int pk_upload_constants(const ConstMats& h) {
  NSB_CUDA(cudaMemcpyToSymbol(cm, &h, sizeof(ConstMats)));
  return 0;
}

#define DISPATCH_N(c, ...)                                   \
  do {                                                       \
    switch ((c)->lx1) {                                      \
      case 4: { constexpr int N = 4; __VA_ARGS__; } break;   \
      case 6: { constexpr int N = 6; __VA_ARGS__; } break;   \
      case 8: { constexpr int N = 8; __VA_ARGS__; } break;   \
      default: nsb_set_error("pcg kernels: unsupported lx1=%d", (c)->lx1); return 1; \
    }                                                        \
  } while (0)

int pk_gradt(Ctx* c, const double* p, double* w) {
  DISPATCH_N(c, k_gradt3<N, 0><<<c->nel, PK_TPB, 0, c->stream>>>(p, w, c->RW2, nullptr, nullptr, nullptr, c->n, c->n2, nullptr, nullptr, nullptr, nullptr, nullptr));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
template <class K>
static int set_smem(K kernel, size_t bytes);
static int persistent_grid(Ctx* c) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  return std::min(c->nel, 2 * sms);
}
int pk_pcg_dir_gradt(Ctx* c, int adj) {
  // measured (r1c, cfg 5): a persistent, TMA-pipelined gradt was slower (0.198 ms at 2 CTAs/SM vs 0.157 ms one CTA per element at
  // 3-4 CTAs/SM) and was removed in r2; the div kernels are persistent.
  // With a separately applied preconditioner (pc_kind != 0) the kernels read z = M^-1 r in place of r and ones in place of 1/diag(E).
  const double* zsrc = c->pc_kind ? c->pz : c->pk[0];
  const double* zscale = c->pc_kind ? c->ones2 : c->dinvE[adj];
  if (c->pc_kind == 1 && c->pcg_fused) {          // fused preconditioner: pz holds the element-block part, the coarse parts are added here
    const PMG& m = c->pmg[(adj && c->has_adj_masks) ? 1 : 0];
    // (r2 final batch, measured: the same kernel compiled for 4 CTAs per SM / 40 registers: 0.1764 vs 0.1758 ms -- no gain, not kept)
    DISPATCH_N(c, k_gradt3<N, 2><<<c->nel, PK_TPB, 0, c->stream>>>(c->pz, c->wk[2], c->RW2, nullptr, c->pk[2], c->cgs + 3, c->n, c->n2,
                                                                m.xc, nullptr, m.hat, nullptr, c->pk[1]));
    nsb_count_launch();
    NSB_CUDA(cudaGetLastError());
    return 0;
  }
  DISPATCH_N(c, k_gradt3<N, 1><<<c->nel, PK_TPB, 0, c->stream>>>(zsrc, c->wk[2], c->RW2, zscale, c->pk[2], c->cgs + 3,
                                                              c->n, c->n2, nullptr, nullptr, nullptr, nullptr, nullptr));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
template <int N>
static constexpr size_t div3_smem() {
  using S1 = Shp<N, N, N>;
  using SA = Shp<N - 2, N, N>;
  using SB = Shp<N - 2, N - 2, N>;
  constexpr int SU = 3 * S1::size, SBB = 9 * SB::size;
  return sizeof(double) * ((SU > SBB ? SU : SBB) + 6 * SA::size + 10 * (N - 2) * (N - 2) * (N - 2));
}
template <class K>
static int set_smem(K kernel, size_t bytes) {
  NSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // the persistent kernels size their staging buffers for 2-3 CTAs per SM: ask for the largest shared-memory carve-out
  if (bytes > 64 * 1024) NSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  return 0;
}

int pk_div(Ctx* c, const double* u, const double* s0, const double* s1, const double* s2, double* q, double sign) {
  DISPATCH_N(c, NSB_TRY(set_smem(k_div3<N, 0>, div3_smem<N>())); k_div3<N, 0><<<c->nel, PK_TPB, div3_smem<N>(), c->stream>>>(u, s0, s1, s2, q, c->RW2, nullptr, nullptr, nullptr, nullptr,
                                                               nullptr, 0, c->n, c->n2, sign));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int pk_pcg_div(Ctx* c, int adj, int fused) {
  const double* s0 = c->mbinv[adj][0];
  const double* s1 = c->mask_same[adj] ? nullptr : c->mbinv[adj][1];
  const double* s2 = c->mask_same[adj] ? nullptr : c->mbinv[adj][2];
  if (c->persistent_pcg && !fused && c->mask_same[adj]) {      // single staging buffer with early refill, 3 CTAs per SM (NSB_PERSISTENT=0: k_div3)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    DISPATCH_N(c, DivQArgs a{c->wk[2], s0, c->RW2, c->pk[2], c->pk[3], c->n, c->n2};
               NSB_TRY(set_smem(k_div3q<N>, DivQ<N>::smem));
               k_div3q<N><<<std::min(c->nel, 3 * sms), PK_TPB, DivQ<N>::smem, c->stream>>>(a, c->cgs + 3, c->red_part, c->red_count, c->red_out,
                                                                                         c->nranks == 1, c->nel));
    nsb_count_launch();
    NSB_CUDA(cudaGetLastError());
    return 0;
  }
  DISPATCH_N(c, NSB_TRY(set_smem(k_div3<N, 1>, div3_smem<N>())); k_div3<N, 1><<<c->nel, PK_TPB, div3_smem<N>(), c->stream>>>(c->wk[2], s0, s1, s2, c->pk[3], c->RW2, c->pk[2], c->cgs + 3,
                                                               c->red_part, c->red_count, c->red_out, c->nranks == 1, c->n,
                                                               c->n2, 1.0));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

template <int N, bool PF>
static constexpr size_t ax3_smem() { return sizeof(double) * (12 * Shp<N, N, N>::size + (PF ? 7 * N * N * N : 0)); }

template <int N, bool PF>
static int launch_axhelm3(Ctx* c, int mode, const double* u, double* w, const double* b, int nfields, double h1, double h2) {
  constexpr size_t sm = ax3_smem<N, PF>();
  if (mode == 0) {
    NSB_TRY(set_smem(k_axhelm3<N, 0, PF>, sm));
    k_axhelm3<N, 0, PF><<<c->nel, AxCfg<N>::TPB, sm, c->stream>>>(u, w, nullptr, c->G, c->bm1, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                   nullptr, 0, nfields, c->n, h1, h2);
  } else if (mode == 1) {
    NSB_TRY(set_smem(k_axhelm3<N, 1, PF>, sm));
    k_axhelm3<N, 1, PF><<<c->nel, AxCfg<N>::TPB, sm, c->stream>>>(u, w, b, c->G, c->bm1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                   0, nfields, c->n, h1, h2);
  } else {
    NSB_TRY(set_smem(k_axhelm3<N, 2, PF>, sm));
    k_axhelm3<N, 2, PF><<<c->nel, AxCfg<N>::TPB, sm, c->stream>>>(c->rk, c->wk[2], nullptr, c->G, c->bm1, c->dinvH, c->wk[1], c->cgs,
                                                                   c->red_part, c->red_count, c->red_out, c->nranks == 1, nfields, c->n,
                                                                   h1, h2);
  }
  return 0;
}

int pk_axhelm(Ctx* c, int mode, const double* u, double* w, const double* b, int nfields, double h1, double h2) {
  const bool pf = true;      // G and bm1 prefetched with cp.async (r1d: 0.580 -> 0.563 ms); the non-prefetching instantiation is gone
  if (mode == 2 && nfields == 3 && c->ax_persistent && c->lx1 == 8) {      // Helmholtz-CG head: persistent, TMA-pipelined (lx1 = 8: 112 KB per CTA)
    constexpr int N = 8;
    AxPArgs a{c->rk, c->dinvH, c->G, c->bm1, c->wk[1], c->wk[2], c->n, h1, h2, perm_h_active(c) ? 1 : 0};
    NSB_TRY(set_smem(k_axhelm3p<N>, AxP<N>::smem));
    k_axhelm3p<N><<<persistent_grid(c), AxCfg<N>::TPB, AxP<N>::smem, c->stream>>>(a, c->cgs, c->red_part, c->red_count, c->red_out,
                                                                                c->nranks == 1, c->nel);
    nsb_count_launch();
    NSB_CUDA(cudaGetLastError());
    return 0;
  }
  (void)pf;
  DISPATCH_N(c, NSB_TRY((launch_axhelm3<N, true>(c, mode, u, w, b, nfields, h1, h2))));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
