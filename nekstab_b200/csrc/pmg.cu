// pmg.cu -- three-level additive pressure preconditioner for the CG on E = D (mask B~^-1 QQ^T) D^T.
//
// The reference solves this system with GMRES preconditioned by Nek5000's hybrid Schwarz multigrid (`preconditioner =
// semg_xxt`, examples/cylinder/stability/direct/1cyl.par:28; [UPSTREAM] hsmg.f hsmg_solve: element-local fast-
// diagonalisation solves + a coarse problem on the element-vertex mesh).  The north-star's Jacobi-PCG needs 2-7e3
// iterations per step on the shipped meshes; a preconditioner of the reference's class changes the iteration count,
// not the converged pressure.  Here (symmetric positive definite, so CG stays CG):
//
//   M^-1 r = sum_e R_e^T Et_e^-1 R_e r     element blocks, Et_e = separable (box) approximation of the element's diagonal
//                                          block of E, inverted by fast diagonalisation:  (S x S x S) L^-1 (S x S x S)^T
//          + P  diag(P^T E P)^-1 P^T r     one Jacobi sweep on the Q1 space of the element-vertex mesh
//          + Pa (Pa^T E Pa)^-1  Pa^T r     piecewise constants on <= 512 element aggregates (recursive coordinate
//                                          bisection), dense inverse replicated on every GPU and L2-resident
//
// Both Galerkin pieces are measured from E itself at setup (probing with a distance-2 colouring of the vertex graph;
// one E application per aggregate), so boundary conditions, deformation and the adjoint mask set need no special cases.
// Per CG iteration the preconditioner streams r and z once (2 n2 words + 126 words per element of FDM factors):
// ~2 % of the bytes of one E application; the coarse levels are three latency-bound launches.
// CPU restatement: oracle/pmg.py (tests/test_gpu_pmg.py compares M^-1 r, solutions and iteration counts).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>

#include "elem_common.cuh"

static __constant__ double pm_l[2][8];   // Q1 hat functions (1-z)/2, (1+z)/2 at the GL(lx2) points

// ---------------------------------------------------------------------------------------------- device kernels
template <int D, int L>
struct PmShape {
  static constexpr int NP = (D == 3) ? L * L * L : L * L;
  static constexpr int NK = (D == 3) ? 8 : 4;
  static constexpr int NPL = (NP + 31) / 32;   // points per lane
};

template <int D, int L>
__device__ __forceinline__ double pm_phi(int k, int p) {
  const int i0 = p % L, i1 = (p / L) % L;
  double f = pm_l[k & 1][i0] * pm_l[(k >> 1) & 1][i1];
  if (D == 3) f *= pm_l[(k >> 2) & 1][p / (L * L)];
  return f;
}

// rc[e][k] = sum_p phi_k(p) r_e(p); one warp per element (grid-stride).  The hat-function factors of a lane's points sit in
// registers (they do not depend on the element), the 2^D sums are built by successive splitting (k2, k1, k0) and reduced
// with a transposed butterfly: 9 shuffles instead of 40, fixed tree => deterministic.  (First generation: hats from the
// constant bank with lane-varying index, then from a shared table: 0.039 / 0.029 ms for 34 MB.)
template <int D, int L>
__global__ void __launch_bounds__(256) k_pm_restrict(const double* __restrict__ r, double* __restrict__ rc, int nel,
                                                     const CGState* skip) {
  if (skip && skip->done) return;
  using S = PmShape<D, L>;
  constexpr int NK = S::NK, NP = S::NP, NPL = S::NPL;
  __shared__ double sl1[8];
  if (threadIdx.x < L) sl1[threadIdx.x] = pm_l[1][threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  double fa[NPL], fb[NPL], fc[NPL];
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int p = lane + 32 * q;
    const bool in = p < NP;
    fa[q] = in ? sl1[p % L] : 0.0;
    fb[q] = in ? sl1[(p / L) % L] : 0.0;
    fc[q] = (in && D == 3) ? sl1[p / (L * L)] : 0.0;
  }
  for (int e = blockIdx.x * wpb + (threadIdx.x >> 5); e < nel; e += gridDim.x * wpb) {
    double acc[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] = 0.0;
    const double* re = r + (long long)e * NP;
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int p = lane + 32 * q;
      if (p < NP) {
        const double v = re[p];
        if (D == 3) {
          const double t1 = v * fc[q], t0 = v - t1;
          const double t01 = t0 * fb[q], t00 = t0 - t01, t11 = t1 * fb[q], t10 = t1 - t11;
          double h;
          h = t00 * fa[q]; acc[1] += h; acc[0] += t00 - h;
          h = t01 * fa[q]; acc[3] += h; acc[2] += t01 - h;
          h = t10 * fa[q]; acc[5] += h; acc[4] += t10 - h;
          h = t11 * fa[q]; acc[7] += h; acc[6] += t11 - h;
        } else {
          const double t1 = v * fb[q], t0 = v - t1;
          double h;
          h = t0 * fa[q]; acc[1] += h; acc[0] += t0 - h;
          h = t1 * fa[q]; acc[3] += h; acc[2] += t1 - h;
        }
      }
    }
    // transposed butterfly: after the steps with offsets 16, 8 (, 4) every lane holds ONE of the 2^D sums, partially reduced
    double b1;
    if (D == 3) {
      double b4[4], b2[2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool hi = lane & 16;
        const double mine = hi ? acc[4 + j] : acc[j], send = hi ? acc[j] : acc[4 + j];
        b4[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool hi = lane & 8;
        const double mine = hi ? b4[2 + j] : b4[j], send = hi ? b4[j] : b4[2 + j];
        b2[j] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const bool hi = lane & 4;
        const double mine = hi ? b2[1] : b2[0], send = hi ? b2[0] : b2[1];
        b1 = mine + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
      b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
      if ((lane & 3) == 0) rc[(long long)e * NK + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = b1;
    } else {
      double b2[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool hi = lane & 16;
        const double mine = hi ? acc[2 + j] : acc[j], send = hi ? acc[j] : acc[2 + j];
        b2[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      {
        const bool hi = lane & 8;
        const double mine = hi ? b2[1] : b2[0], send = hi ? b2[0] : b2[1];
        b1 = mine + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      b1 += __shfl_xor_sync(0xffffffffu, b1, 4);
      b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
      b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
      if ((lane & 7) == 0) rc[(long long)e * NK + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)] = b1;
    }
  }
}

// threads [0, nv): xv[v] = d1inv[v] * sum of rc over the (element, corner) entries of vertex v (ascending order);
// then one warp per local aggregate: ra[first + a] = sum over its elements of sum_k rc[e][k]  (sum_k phi_k = 1)
// vmode 0: sum the entries (single rank); 1: the entries were already assembled across ranks by the vertex gather-scatter
// (every copy holds the total): take the first; 2: skip the vertex part.  Entries of ra owned by other ranks are zeroed
// (they are filled by the all-reduce that follows).
__global__ void k_pm_coarse(int nv, const int* __restrict__ voff, const int* __restrict__ vent, const double* __restrict__ d1inv,
                            const double* __restrict__ rc, double* __restrict__ xv, int nagg, int nagg_loc, int agg_first,
                            const int* __restrict__ aoff, const int* __restrict__ aent, int nk, double* __restrict__ ra,
                            int vmode, int amode, const CGState* skip) {
  if (skip && skip->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int vthreads = ((max(nv, nagg) + 31) / 32) * 32;
  if (t < vthreads) {
    if (t < nv && vmode != 2) {
      double s = 0.0;
      if (vmode == 0) for (int j = voff[t]; j < voff[t + 1]; ++j) s += rc[vent[j]];
      else s = rc[vent[voff[t]]];
      xv[t] = d1inv ? d1inv[t] * s : s;
    }
    if (amode && t < nagg && (t < agg_first || t >= agg_first + nagg_loc)) ra[t] = 0.0;
    return;
  }
  if (!amode) return;
  const int a = (t - vthreads) >> 5, lane = t & 31;
  if (a >= nagg_loc) return;
  double s = 0.0;
  for (int j = aoff[a] + lane; j < aoff[a + 1]; j += 32) {
    const double* q = rc + (long long)aent[j] * nk;
    double se = 0.0;
    for (int k = 0; k < nk; ++k) se += q[k];
    s += se;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) ra[agg_first + a] = s;
}

// x2 = A2inv ra  (one warp per row; A2inv is <= 2 MB and stays in L2)
__global__ void k_pm_gemv(int nagg, const double* __restrict__ A, const double* __restrict__ ra, double* __restrict__ x2,
                          const CGState* skip) {
  if (skip && skip->done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nagg) return;
  double s = 0.0;
  for (int j = lane; j < nagg; j += 32) s = fma(A[(long long)row * nagg + j], ra[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) x2[row] = s;
}

// out_e = sum_k phi_k xv[vid[e][k]] + x2[agg[e]]   (setup probing only)
template <int D, int L>
__global__ void k_pm_prolong(const double* __restrict__ xv, const int* __restrict__ vid, const double* __restrict__ x2,
                             const int* __restrict__ agg, double* __restrict__ out, int nel) {
  using S = PmShape<D, L>;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nel * S::NP) return;
  const int e = (int)(t / S::NP), p = (int)(t % S::NP);
  double v = x2 ? x2[agg[e]] : 0.0;
  if (xv) {
#pragma unroll
    for (int k = 0; k < S::NK; ++k) v = fma(pm_phi<D, L>(k, p), xv[vid[e * S::NK + k]], v);
  }
  out[t] = v;
}

// z_e = FDM_e(r_e) + sum_k phi_k xv[vid[e][k]] + x2[agg[e]];  rtz = sum z r  (deterministic two-stage reduction).
// mode 0: no CG bookkeeping (operator test); 1: single rank -> the last CTA sets rtz1/beta; 2: multi rank (sum only).
// Element-block kernel: the CTA holds EPB elements in shared memory and every thread owns one
// COLUMN of one element per tensor stage (L loads, L*L DFMAs against rows of the element's S read as 128-bit broadcasts,
// L stores) -- ncu on the first generation (one warp per element, one point per lane and stage) showed the shared-memory
// pipe as the limiter (12 LDS per 6 DFMA) and the Q1 hat functions served from the constant bank with lane-varying indices.
// Rows are padded to an odd length so that column accesses along every axis stay at the 64-bit minimum of 2 wavefronts.
template <int D, int L>
struct Pm2 {
  static constexpr int NCOL = (D == 3) ? L * L : L;
  static constexpr int EPB = (D == 3) ? (L == 6 ? 8 : (L == 4 ? 16 : 32)) : 32;
  static constexpr int NT = EPB * NCOL;
  static constexpr int PI = (L % 2 == 0) ? L + 1 : L;
  static constexpr int ESZ = (D == 3) ? L * L * PI : L * PI;
  static constexpr int NP = (D == 3) ? L * L * L : L * L;
  static constexpr int NK = (D == 3) ? 8 : 4;
};

template <int D, int L>
__global__ void __launch_bounds__(Pm2<D, L>::NT) k_pm_apply2(const double* __restrict__ r, double* __restrict__ z, int nel,
                                                            const double* __restrict__ Sg, const double* __restrict__ lamg,
                                                            const double* __restrict__ xv, const int* __restrict__ vid,
                                                            const double* __restrict__ x2, const int* __restrict__ agg,
                                                            CGState* cgs, int mode, double* part, unsigned* counter, double* out) {
  using P = Pm2<D, L>;
  constexpr int NP = P::NP, NK = P::NK, LL = L * L, PI = P::PI, ESZ = P::ESZ, EPB = P::EPB, NT = P::NT, NCOL = P::NCOL;
  if (mode && cgs->done) return;
  __shared__ __align__(16) double sS[EPB * D * LL];
  __shared__ double sA[EPB * ESZ], sB[EPB * ESZ], sL[EPB * D * L], sX[EPB * (NK + 2)], sl1[8], sred[32];
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * EPB;
  const int ne = min(EPB, nel - e0);
  // ---- load phase (coalesced): residual -> sA (padded rows), FDM factors, corner values, aggregate value
  for (int t = tid; t < ne * NP; t += NT) {
    const int el = t / NP, p = t - el * NP;
    sA[el * ESZ + (p / L) * PI + (p % L)] = r[(long long)e0 * NP + t];
  }
  for (int t = tid; t < ne * D * LL; t += NT) sS[t] = Sg[(long long)e0 * D * LL + t];
  for (int t = tid; t < ne * D * L; t += NT) sL[t] = lamg[(long long)e0 * D * L + t];
  for (int t = tid; t < ne * NK; t += NT) {
    const int el = t / NK, k = t - el * NK;
    sX[el * (NK + 2) + k] = xv ? xv[vid[(long long)(e0 + el) * NK + k]] : 0.0;
  }
  if (tid < ne) sX[tid * (NK + 2) + NK] = x2 ? x2[agg[e0 + tid]] : 0.0;
  if (tid < L) sl1[tid] = pm_l[1][tid];
  __syncthreads();
  if (tid < ne) {                                   // threshold scale of the element: sum_d max_i lam_d[i]
    double mx = 0.0;
    for (int d = 0; d < D; ++d) {
      double m = sL[(tid * D + d) * L];
      for (int i = 1; i < L; ++i) m = fmax(m, sL[(tid * D + d) * L + i]);
      mx += m;
    }
    sX[tid * (NK + 2) + NK + 1] = mx;
  }
  __syncthreads();
  const int el = tid / NCOL, c = tid - el * NCOL;
  const bool act = el < ne;
  double* in = sA + el * ESZ;
  double* ou = sB + el * ESZ;
  const int ca = c % L, cb = c / L;                 // the two (3-D) or one (2-D: cb = 0) indices that label the column
#pragma unroll
  for (int pass = 0; pass < 2 * D; ++pass) {
    const int d = (pass < D) ? pass : (2 * D - 1 - pass);      // forward 0..D-1, backward D-1..0 (the operators commute)
    const bool fwd = pass < D;
    if (act) {
      int base, str;
      if (D == 3) {
        if (d == 0) { base = (cb * L + ca) * PI; str = 1; }
        else if (d == 1) { base = cb * L * PI + ca; str = PI; }
        else { base = cb * PI + ca; str = L * PI; }
      } else {
        if (d == 0) { base = c * PI; str = 1; }
        else { base = c; str = PI; }
      }
      const double* Sd = sS + (el * D + d) * LL;
      double v[L], o[L];
#pragma unroll
      for (int a = 0; a < L; ++a) v[a] = in[base + a * str];
      if (fwd) {
#pragma unroll
        for (int i = 0; i < L; ++i) o[i] = 0.0;
#pragma unroll
        for (int a = 0; a < L; ++a) {
          const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
#pragma unroll
          for (int i2 = 0; i2 < L / 2; ++i2) {
            const double2 s2 = row[i2];
            o[2 * i2] = fma(s2.x, v[a], o[2 * i2]);
            o[2 * i2 + 1] = fma(s2.y, v[a], o[2 * i2 + 1]);
          }
        }
        if (d == D - 1) {                           // all axes are modal now: divide by the eigenvalue sum
          const double* lam = sL + el * D * L;
          const double mx = sX[el * (NK + 2) + NK + 1];
          const double l01 = (D == 3) ? lam[ca] + lam[L + cb] : lam[c];
#pragma unroll
          for (int i = 0; i < L; ++i) {
            const double den = l01 + lam[(D - 1) * L + i];
            o[i] = (den > 1e-12 * mx) ? o[i] / den : 0.0;
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < L; ++a) {
          const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
          double sacc = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < L / 2; ++i2) {
            const double2 s2 = row[i2];
            sacc = fma(s2.x, v[2 * i2], sacc);
            sacc = fma(s2.y, v[2 * i2 + 1], sacc);
          }
          o[a] = sacc;
        }
      }
#pragma unroll
      for (int a = 0; a < L; ++a) ou[base + a * str] = o[a];
    }
    __syncthreads();
    double* tmp = in; in = ou; ou = tmp;
  }
  // ---- z = block solve + trilinear interpolation of the corner values + aggregate value;  partial z.r
  const double* res = ((2 * D) % 2 == 0) ? sA : sB;  // an even number of passes ends in the buffer it started from
  double dot[1] = {0.0};
  for (int t = tid; t < ne * NP; t += NT) {
    const int e1 = t / NP, p = t - e1 * NP;
    const int i0 = p % L, i1 = (p / L) % L;
    const double* X = sX + e1 * (NK + 2);
    const double a0 = sl1[i0], a1 = sl1[i1];
    double c0 = fma(X[1] - X[0], a0, X[0]), c1 = fma(X[3] - X[2], a0, X[2]);
    double q = fma(c1 - c0, a1, c0);
    if (D == 3) {
      const double d0 = fma(X[5] - X[4], a0, X[4]), d1 = fma(X[7] - X[6], a0, X[6]);
      const double q1 = fma(d1 - d0, a1, d0);
      q = fma(q1 - q, sl1[p / LL], q);
    }
    const double val = res[e1 * ESZ + (p / L) * PI + i0] + q + X[NK];
    const long long gi = (long long)e0 * NP + t;
    z[gi] = val;
    dot[0] = fma(val, r[gi], dot[0]);
  }
  if (grid_sum_finish<1>(dot, part, counter, out, sred) && mode == 1 && threadIdx.x == 0) {
    cgs->rtz1 = out[0];
    cgs->beta = (cgs->iter == 0) ? 0.0 : out[0] / cgs->rtz2;
  }
}

// ---------------------------------------------------------------------------------------------- fused pressure-CG tail (3-D)
// One kernel for everything of a CG iteration that touches the mesh-2 vectors between the E application and the next direction:
//   r -= alpha Ep ; |r|^2 (Nek norm: r^2/bm2) ; rc = P^T r (Q1 restriction) ; zloc = FDM_e(r) ; zloc . r
// (the solution update x += alpha p rides along in the next direction kernel, which has p in registers anyway)
// (r1d: k_pcg_update 0.058 + k_pm_restrict 0.019 + k_pm_apply2 0.074 ms, r streamed three times).  The coarse-level parts of
// z = M^-1 r are NOT added here: they need the vertex / aggregate sums of ALL elements, so the direction kernel (k_gradt3 MODE 2,
// pcg_kernels.cu) adds the trilinear interpolation of the vertex values and the aggregate value while it forms p = z + beta p,
// and z.r is assembled from the level contributions: zloc.r (here) + sum_v xv rv + ra^T A2^-1 ra (k_pm_coarse_dot, k_pm_gemv_fin).
// init = 1: start of a solve (x = 0, p = 0, r given).  The two sums go to out[0..1] (deterministic two-stage reduction).
template <int D, int L>
__global__ void __launch_bounds__(Pm2<D, L>::NT) k_pcg_fused(double* __restrict__ r, double* __restrict__ x, double* __restrict__ pdir,
                                                            const double* __restrict__ ep, const double* __restrict__ bm2inv,
                                                            double* __restrict__ zloc, double* __restrict__ rc, double* __restrict__ rc0,
                                                            int nel, const double* __restrict__ Sg, const double* __restrict__ lamg,
                                                            const CGState* __restrict__ cgs, int init, double* part, unsigned* counter,
                                                            double* out) {
  using P = Pm2<D, L>;
  constexpr int NP = P::NP, NK = P::NK, LL = L * L, PI = P::PI, ESZ = P::ESZ, EPB = P::EPB, NT = P::NT, NCOL = P::NCOL;
  constexpr int PER = (EPB * NP + NT - 1) / NT;
  if (!init && cgs->done) return;
  __shared__ __align__(16) double sS[EPB * D * LL];
  __shared__ double sA[EPB * ESZ], sB[EPB * ESZ], sL[EPB * D * L], sMx[EPB], sl0[8], sl1[8], sred[2 * 32];
  __shared__ double sP[EPB * NCOL * 2], sQ[(D == 3) ? EPB * L * 4 : 1];
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * EPB;
  const int ne = min(EPB, nel - e0);
  const double alpha = init ? 0.0 : cgs->alpha;
  // ---- load phase (coalesced): CG update, residual -> sA (padded rows), FDM factors
  double rr[PER];
  double sums[2] = {0.0, 0.0};
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int t = tid + q * NT;
    rr[q] = 0.0;
    if (t < ne * NP) {
      const long long gi = (long long)e0 * NP + t;
      double v = r[gi];
      if (init) {
        x[gi] = 0.0;
        pdir[gi] = 0.0;
      } else {                                    // x += alpha p is done by the next direction kernel (k_gradt3 MODE 2) / k_pcg_xfix
        v = fma(-alpha, ep[gi], v);
        r[gi] = v;
      }
      rr[q] = v;
      sums[0] = fma(v * v, bm2inv[gi], sums[0]);
      const int el = t / NP, p = t - el * NP;
      sA[el * ESZ + (p / L) * PI + (p % L)] = v;
    }
  }
  for (int t = tid; t < ne * D * LL; t += NT) sS[t] = Sg[(long long)e0 * D * LL + t];
  for (int t = tid; t < ne * D * L; t += NT) sL[t] = lamg[(long long)e0 * D * L + t];
  if (tid < L) { sl0[tid] = pm_l[0][tid]; sl1[tid] = pm_l[1][tid]; }
  __syncthreads();
  if (tid < ne) {                                   // threshold scale of the element: sum_d max_i lam_d[i]
    double mx = 0.0;
    for (int d = 0; d < D; ++d) {
      double m = sL[(tid * D + d) * L];
      for (int i = 1; i < L; ++i) m = fmax(m, sL[(tid * D + d) * L + i]);
      mx += m;
    }
    sMx[tid] = mx;
  }
  __syncthreads();
  const int el = tid / NCOL, c = tid - el * NCOL;
  const bool act = el < ne;
  double* in = sA + el * ESZ;
  double* ou = sB + el * ESZ;
  const int ca = c % L, cb = c / L;
#pragma unroll
  for (int pass = 0; pass < 2 * D; ++pass) {
    const int d = (pass < D) ? pass : (2 * D - 1 - pass);
    const bool fwd = pass < D;
    if (act) {
      int base, str;
      if (D == 3) {
        if (d == 0) { base = (cb * L + ca) * PI; str = 1; }
        else if (d == 1) { base = cb * L * PI + ca; str = PI; }
        else { base = cb * PI + ca; str = L * PI; }
      } else {
        if (d == 0) { base = c * PI; str = 1; }
        else { base = c; str = PI; }
      }
      const double* Sd = sS + (el * D + d) * LL;
      double v[L], o[L];
#pragma unroll
      for (int a = 0; a < L; ++a) v[a] = in[base + a * str];
      if (pass == 0) {                              // Q1 restriction, first level: the row's two sums along r
        double s1 = 0.0, s0 = 0.0;
#pragma unroll
        for (int a = 0; a < L; ++a) { s1 = fma(sl1[a], v[a], s1); s0 = fma(sl0[a], v[a], s0); }
        sP[(el * NCOL + c) * 2] = s0;
        sP[(el * NCOL + c) * 2 + 1] = s1;
      }
      if (fwd) {
#pragma unroll
        for (int i = 0; i < L; ++i) o[i] = 0.0;
#pragma unroll
        for (int a = 0; a < L; ++a) {
          const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
#pragma unroll
          for (int i2 = 0; i2 < L / 2; ++i2) {
            const double2 s2 = row[i2];
            o[2 * i2] = fma(s2.x, v[a], o[2 * i2]);
            o[2 * i2 + 1] = fma(s2.y, v[a], o[2 * i2 + 1]);
          }
        }
        if (d == D - 1) {
          const double* lam = sL + el * D * L;
          const double mx = sMx[el];
          const double l01 = (D == 3) ? lam[ca] + lam[L + cb] : lam[c];
#pragma unroll
          for (int i = 0; i < L; ++i) {
            const double den = l01 + lam[(D - 1) * L + i];
            o[i] = (den > 1e-12 * mx) ? o[i] / den : 0.0;
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < L; ++a) {
          const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
          double sacc = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < L / 2; ++i2) {
            const double2 s2 = row[i2];
            sacc = fma(s2.x, v[2 * i2], sacc);
            sacc = fma(s2.y, v[2 * i2 + 1], sacc);
          }
          o[a] = sacc;
        }
      }
#pragma unroll
      for (int a = 0; a < L; ++a) ou[base + a * str] = o[a];
    }
    __syncthreads();
    // ---- Q1 restriction, remaining levels, in the shadow of the tensor passes (fixed summation order => deterministic)
    if (D == 3) {
      if (pass == 0 && tid < ne * L * 4) {          // sum over s (ca) for every (element, t-plane cb, b0, b1)
        const int e1 = tid / (L * 4), rem = tid - e1 * (L * 4);
        const int kb = rem >> 2, b0 = rem & 1, b1 = (rem >> 1) & 1;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < L; ++j) acc = fma(b1 ? sl1[j] : sl0[j], sP[(e1 * NCOL + kb * L + j) * 2 + b0], acc);
        sQ[tid] = acc;
      }
      if (pass == 1 && tid < ne * NK) {             // sum over t (cb) for every (element, corner)
        const int e1 = tid / NK, k = tid - e1 * NK;
        const int b0 = k & 1, b1 = (k >> 1) & 1, b2 = (k >> 2) & 1;
        double acc = 0.0;
#pragma unroll
        for (int kb = 0; kb < L; ++kb) acc = fma(b2 ? sl1[kb] : sl0[kb], sQ[e1 * L * 4 + kb * 4 + b1 * 2 + b0], acc);
        rc[(long long)(e0 + e1) * NK + k] = acc;
        if (rc0) rc0[(long long)(e0 + e1) * NK + k] = acc;
      }
    } else {
      if (pass == 0 && tid < ne * NK) {             // 2-D: sum over s (row index c) for every (element, corner)
        const int e1 = tid / NK, k = tid - e1 * NK;
        const int b0 = k & 1, b1 = (k >> 1) & 1;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < L; ++j) acc = fma(b1 ? sl1[j] : sl0[j], sP[(e1 * NCOL + j) * 2 + b0], acc);
        rc[(long long)(e0 + e1) * NK + k] = acc;
        if (rc0) rc0[(long long)(e0 + e1) * NK + k] = acc;
      }
    }
    double* tmp = in; in = ou; ou = tmp;
  }
  // ---- zloc = block solve (an even number of passes ends in the buffer it started from); partial zloc . r
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int t = tid + q * NT;
    if (t < ne * NP) {
      const int e1 = t / NP, p = t - e1 * NP;
      const double val = sA[e1 * ESZ + (p / L) * PI + (p % L)];
      zloc[(long long)e0 * NP + t] = val;
      sums[1] = fma(val, rr[q], sums[1]);
    }
  }
  grid_sum_finish<2>(sums, part, counter, out, sred);
}

// ---- persistent, TMA-pipelined form of k_pcg_fused (3-D).  r2 measurement of the one-group-per-CTA kernel above: 0.120 ms for 296 MB
// (38 % of the HBM roofline): every CTA first waits for its loads, then runs ten barrier-separated phases with nothing in flight.
// Here one CTA per SM slot walks over groups of EPB elements; the three streamed inputs of a group (r, Ep, 1/bm2) and its FDM factors
// sit in ONE staging buffer that is refilled by bulk TMA copies as soon as it has been consumed (the scheme of k_axhelm3p / k_div3q):
// the streamed half right after the load phase, the factor half right after the last tensor pass.
__device__ __forceinline__ uint32_t pm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pm_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pm_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(pm_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pm_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PM_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PM_WAIT_DONE;\n"
      "bra PM_WAIT_LOOP;\n"
      "PM_WAIT_DONE:\n"
      "}\n" ::"r"(pm_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int L>
struct PfP {
  static constexpr int D = 3, EPB = 8, NP = L * L * L, LL = L * L, NCOL = LL, NT = EPB * NCOL;
  static constexpr int PI = (L % 2 == 0) ? L + 1 : L, ESZ = LL * PI, NK = 8;
  static constexpr int PER = (EPB * NP + NT - 1) / NT;
  static constexpr int stream = 3 * EPB * NP;                   // r | Ep | 1/bm2, EPB elements each
  static constexpr int fac = EPB * (D * LL + D * L);            // S | lam
  static constexpr int work = 2 * EPB * ESZ + EPB * NCOL * 2 + EPB * L * 4 + EPB + 16;
  static constexpr size_t smem = sizeof(double) * (stream + fac + work) + 2 * sizeof(uint64_t);
};
struct PfArgs {
  double* r; const double* ep; const double* bm2inv; const double* Sg; const double* lamg;
  double *x, *pdir, *zloc, *rc, *rc0;
  int nel, init;
  double alpha;
};
template <int L>
__device__ __forceinline__ void pfp_issue_stream(const PfArgs* A, double* stg, uint64_t* bar, int g) {
  using P = PfP<L>;
  const int e0 = g * P::EPB, ne = min(P::EPB, A->nel - e0);
  const uint32_t bytes = (uint32_t)(ne * P::NP * sizeof(double));
  const long long off = (long long)e0 * P::NP;
  pm_mbar_expect_tx(bar, 3 * bytes);
  pm_tma_g2s(stg, A->r + off, bytes, bar);
  pm_tma_g2s(stg + P::EPB * P::NP, A->ep + off, bytes, bar);
  pm_tma_g2s(stg + 2 * P::EPB * P::NP, A->bm2inv + off, bytes, bar);
}
template <int L>
__device__ __forceinline__ void pfp_issue_fac(const PfArgs* A, double* fac, uint64_t* bar, int g) {
  using P = PfP<L>;
  const int e0 = g * P::EPB, ne = min(P::EPB, A->nel - e0);
  const uint32_t bs = (uint32_t)(ne * P::D * P::LL * sizeof(double)), bl = (uint32_t)(ne * P::D * L * sizeof(double));
  pm_mbar_expect_tx(bar, bs + bl);
  pm_tma_g2s(fac, A->Sg + (long long)e0 * P::D * P::LL, bs, bar);
  pm_tma_g2s(fac + P::EPB * P::D * P::LL, A->lamg + (long long)e0 * P::D * L, bl, bar);
}

template <int L>
__global__ void __launch_bounds__(PfP<L>::NT, 2) k_pcg_fused_p(PfArgs args, const CGState* __restrict__ cgs, double* part,
                                                              unsigned* counter, double* out) {
  using P = PfP<L>;
  constexpr int D = 3, NP = P::NP, NK = P::NK, LL = P::LL, PI = P::PI, ESZ = P::ESZ, EPB = P::EPB, NT = P::NT, NCOL = P::NCOL, PER = P::PER;
  if (!args.init && cgs->done) return;
  extern __shared__ __align__(128) double dsm[];
  double* stg = dsm;                               // [3][EPB*NP]
  double* sS = dsm + P::stream;                    // [EPB][D][LL]
  double* sL = sS + EPB * D * LL;                  // [EPB][D][L]
  double* sA = sL + EPB * D * L;
  double* sB = sA + EPB * ESZ;
  double* sP = sB + EPB * ESZ;                     // [EPB][NCOL][2]
  double* sQ = sP + EPB * NCOL * 2;                // [EPB][L][4]
  double* sMx = sQ + EPB * L * 4;                  // [EPB]
  double* sl0 = sMx + EPB;                         // [8]
  double* sl1 = sl0 + 8;                           // [8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sl1 + 8);
  __shared__ double sred[2 * 32];
  __shared__ PfArgs sargs;
  const int tid = threadIdx.x;
  const int ngroups = (args.nel + EPB - 1) / EPB;
  if (tid == 0) {
    sargs = args;
    if (!args.init) sargs.alpha = cgs->alpha;
    pm_mbar_init(&bar[0], 1);
    pm_mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < L) { sl0[tid] = pm_l[0][tid]; sl1[tid] = pm_l[1][tid]; }
  __syncthreads();
  const PfArgs* A = &sargs;
  if (tid == 0 && (int)blockIdx.x < ngroups) {
    pfp_issue_stream<L>(A, stg, &bar[0], blockIdx.x);
    pfp_issue_fac<L>(A, sS, &bar[1], blockIdx.x);
  }
  const double alpha = args.init ? 0.0 : cgs->alpha;
  const int init = args.init;
  double sums[2] = {0.0, 0.0};
  const int el = tid / NCOL, c = tid - el * NCOL;
  const int ca = c % L, cb = c / L;
  int it = 0;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
    const uint32_t parity = (uint32_t)(it & 1);
    const int gn = (g + (int)gridDim.x < ngroups) ? g + (int)gridDim.x : -1;
    const int e0 = g * EPB;
    const int ne = min(EPB, A->nel - e0);
    // ---- load phase from the staged group: CG residual update, norm partial, residual -> sA (padded rows)
    double rr[PER];
    pm_mbar_wait(&bar[0], parity);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int t = tid + q * NT;
      rr[q] = 0.0;
      if (t < ne * NP) {
        const long long gi = (long long)e0 * NP + t;
        double v = stg[t];
        if (init) {
          A->x[gi] = 0.0;
          A->pdir[gi] = 0.0;
        } else {
          v = fma(-alpha, stg[EPB * NP + t], v);
          A->r[gi] = v;
        }
        rr[q] = v;
        sums[0] = fma(v * v, stg[2 * EPB * NP + t], sums[0]);
        const int e1 = t / NP, p = t - e1 * NP;
        sA[e1 * ESZ + (p / L) * PI + (p % L)] = v;
      }
    }
    pm_mbar_wait(&bar[1], parity);
    if (tid < ne) {                                 // threshold scale of the element: sum_d max_i lam_d[i]
      double mx = 0.0;
      for (int d = 0; d < D; ++d) {
        double m = sL[(tid * D + d) * L];
        for (int i = 1; i < L; ++i) m = fmax(m, sL[(tid * D + d) * L + i]);
        mx += m;
      }
      sMx[tid] = mx;
    }
    __syncthreads();                                // streamed half consumed
    if (tid == 0 && gn >= 0) pfp_issue_stream<L>(A, stg, &bar[0], gn);
    const bool act = el < ne;
    double* in = sA + el * ESZ;
    double* ou = sB + el * ESZ;
#pragma unroll
    for (int pass = 0; pass < 2 * D; ++pass) {
      const int d = (pass < D) ? pass : (2 * D - 1 - pass);
      const bool fwd = pass < D;
      if (act) {
        int base, str;
        if (d == 0) { base = (cb * L + ca) * PI; str = 1; }
        else if (d == 1) { base = cb * L * PI + ca; str = PI; }
        else { base = cb * PI + ca; str = L * PI; }
        const double* Sd = sS + (el * D + d) * LL;
        double v[L], o[L];
#pragma unroll
        for (int a = 0; a < L; ++a) v[a] = in[base + a * str];
        if (pass == 0) {
          double s1 = 0.0, s0 = 0.0;
#pragma unroll
          for (int a = 0; a < L; ++a) { s1 = fma(sl1[a], v[a], s1); s0 = fma(sl0[a], v[a], s0); }
          sP[(el * NCOL + c) * 2] = s0;
          sP[(el * NCOL + c) * 2 + 1] = s1;
        }
        if (fwd) {
#pragma unroll
          for (int i = 0; i < L; ++i) o[i] = 0.0;
#pragma unroll
          for (int a = 0; a < L; ++a) {
            const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
#pragma unroll
            for (int i2 = 0; i2 < L / 2; ++i2) {
              const double2 s2 = row[i2];
              o[2 * i2] = fma(s2.x, v[a], o[2 * i2]);
              o[2 * i2 + 1] = fma(s2.y, v[a], o[2 * i2 + 1]);
            }
          }
          if (d == D - 1) {
            const double* lam = sL + el * D * L;
            const double mx = sMx[el];
            const double l01 = lam[ca] + lam[L + cb];
#pragma unroll
            for (int i = 0; i < L; ++i) {
              const double den = l01 + lam[(D - 1) * L + i];
              o[i] = (den > 1e-12 * mx) ? o[i] / den : 0.0;
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < L; ++a) {
            const double2* row = reinterpret_cast<const double2*>(Sd + a * L);
            double sacc = 0.0;
#pragma unroll
            for (int i2 = 0; i2 < L / 2; ++i2) {
              const double2 s2 = row[i2];
              sacc = fma(s2.x, v[2 * i2], sacc);
              sacc = fma(s2.y, v[2 * i2 + 1], sacc);
            }
            o[a] = sacc;
          }
        }
#pragma unroll
        for (int a = 0; a < L; ++a) ou[base + a * str] = o[a];
      }
      __syncthreads();
      if (pass == 0 && tid < ne * L * 4) {
        const int e1 = tid / (L * 4), rem = tid - e1 * (L * 4);
        const int kb = rem >> 2, b0 = rem & 1, b1 = (rem >> 1) & 1;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < L; ++j) acc = fma(b1 ? sl1[j] : sl0[j], sP[(e1 * NCOL + kb * L + j) * 2 + b0], acc);
        sQ[tid] = acc;
      }
      if (pass == 1 && tid < ne * NK) {
        const int e1 = tid / NK, k = tid - e1 * NK;
        const int b0 = k & 1, b1 = (k >> 1) & 1, b2 = (k >> 2) & 1;
        double acc = 0.0;
#pragma unroll
        for (int kb = 0; kb < L; ++kb) acc = fma(b2 ? sl1[kb] : sl0[kb], sQ[e1 * L * 4 + kb * 4 + b1 * 2 + b0], acc);
        A->rc[(long long)(e0 + e1) * NK + k] = acc;
        if (A->rc0) A->rc0[(long long)(e0 + e1) * NK + k] = acc;
      }
      if (pass == 2 * D - 1 && tid == 0 && gn >= 0) pfp_issue_fac<L>(A, sS, &bar[1], gn);      // factor half consumed
      double* tmp = in; in = ou; ou = tmp;
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int t = tid + q * NT;
      if (t < ne * NP) {
        const int e1 = t / NP, p = t - e1 * NP;
        const double val = sA[e1 * ESZ + (p / L) * PI + (p % L)];
        A->zloc[(long long)e0 * NP + t] = val;
        sums[1] = fma(val, rr[q], sums[1]);
      }
    }
    __syncthreads();                                // sA is rewritten by the next group's load phase
  }
  grid_sum_finish<2>(sums, part, counter, out, sred);
}

// Coarse levels with their share of z.r.  Threads [0, vthreads): one vertex each, xv = d1inv * rv with rv = the sum of the corner
// sums over ALL (element, corner) entries of the vertex on ALL ranks (`assembled`: rc already holds that total in every copy after the
// vertex gather-scatter; otherwise it is summed here from the local entries); contribution to z.r: xv * (sum of the LOCAL entries
// rc0), which adds up to sum_v xv rv over the ranks.  Then one warp per local aggregate: ra = sum of its elements' corner sums
// (sum_k phi_k = 1); entries of ra owned by other ranks are zeroed (filled by the all-reduce).  out[0] receives sum xv rv_local.
__global__ void __launch_bounds__(128) k_pm_coarse_dot(int nv, const int* __restrict__ voff, const int* __restrict__ vent,
                                                       const double* __restrict__ d1inv, const double* __restrict__ rc,
                                                       const double* __restrict__ rc0, double* __restrict__ xv, int nagg, int nagg_loc,
                                                       int agg_first, const int* __restrict__ aoff, const int* __restrict__ aent, int nk,
                                                       double* __restrict__ ra, int assembled, const CGState* skip, int init,
                                                       double* part, unsigned* counter, double* out) {
  if (!init && skip && skip->done) return;
  __shared__ double sred[32];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int vthreads = ((max(nv, nagg) + 31) / 32) * 32;
  double contrib[1] = {0.0};
  if (t < vthreads) {
    if (t < nv) {
      double sl = 0.0;
      for (int j = voff[t]; j < voff[t + 1]; ++j) sl += rc0[vent[j]];
      const double st = assembled ? rc[vent[voff[t]]] : sl;
      const double xx = d1inv[t] * st;
      xv[t] = xx;
      contrib[0] = xx * sl;
    }
    if (t < nagg && (t < agg_first || t >= agg_first + nagg_loc)) ra[t] = 0.0;
  } else {
    const int a = (t - vthreads) >> 5, lane = t & 31;
    if (a < nagg_loc) {
      double s = 0.0;
      for (int j = aoff[a] + lane; j < aoff[a + 1]; j += 32) {
        const double* q = rc0 + (long long)aent[j] * nk;
        double se = 0.0;
        for (int k = 0; k < nk; ++k) se += q[k];
        s += se;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) ra[agg_first + a] = s;
    }
  }
  grid_sum_finish<1>(contrib, part, counter, out, sred);
}

// x2 = A2inv ra (one warp per row) + the aggregate level's share ra^T x2 of z.r; the last block then updates the CG scalars from
// sc[0] = |r|^2 (Nek norm), sc[1] = zloc.r, sc[2] = sum xv rv, and its own sum: what k_pcg_update + k_pm_apply2 used to do.
__global__ void __launch_bounds__(128) k_pm_gemv_fin(int nagg, const double* __restrict__ A, const double* __restrict__ ra,
                                                     double* __restrict__ x2, const double* __restrict__ sc, CGState* cgs, int init,
                                                     double* part, unsigned* counter, double* out) {
  if (!init && cgs->done) return;
  __shared__ double sred[32];
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  double contrib[1] = {0.0};
  if (row < nagg) {
    double s = 0.0;
    for (int j = lane; j < nagg; j += 32) s = fma(A[(long long)row * nagg + j], ra[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) { x2[row] = s; contrib[0] = s * ra[row]; }
  }
  if (grid_sum_finish<1>(contrib, part, counter, out, sred) && threadIdx.x == 0) {
    CGState* s = cgs;
    const double rn = sqrt(fmax(sc[0], 0.0) / s->vol);
    if (init) {
      s->rtz1 = 1.0; s->rtz2 = 1.0; s->beta = 0.0; s->alpha = 0.0;
      s->rnorm = rn; s->iter = 0;
      s->done = (rn <= s->tol) || (s->maxit <= 0);
    } else {
      s->rtz2 = s->rtz1;
      s->rnorm = rn;
      s->iter += 1;
      s->done = (rn <= s->tol) || (s->iter >= s->maxit) || !(rn == rn);
    }
    if (!s->done) {
      const double rtz = (sc[1] + sc[2]) + out[0];
      s->rtz1 = rtz;
      s->beta = (s->iter == 0) ? 0.0 : rtz / s->rtz2;
    }
  }
}

// xc[e][k] = xv[vid[e][k]] (k < 2^D), xc[e][2^D] = x2[agg[e]]: the coarse-level values of an element in one contiguous record, so that
// the direction kernel reads them with one hop (the vid -> xv chain does not survive in L2 between iterations: 2 GB stream through it)
__global__ void k_pm_corner_values(int nel, int nk, const double* __restrict__ xv, const int* __restrict__ vid, const double* __restrict__ x2,
                                   const int* __restrict__ agg, double* __restrict__ xc, const CGState* skip) {
  if (skip && skip->done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * (nk + 1)) return;
  const int e = t / (nk + 1), k = t - e * (nk + 1);
  xc[t] = (k < nk) ? xv[vid[e * nk + k]] : x2[agg[e]];
}

// x += alpha p of the LAST iteration of a solve (the direction kernel that would have applied it is skipped once converged)
__global__ void k_pcg_xfix(double* __restrict__ x, const double* __restrict__ p, const CGState* __restrict__ cgs, long long n2) {
  const double a = cgs->alpha;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    x[i] = fma(a, p[i], x[i]);
}
int pm_pcg_xfix(Ctx* c) {
  const long long nb = std::min<long long>((c->n2 + 1023) / 1024, 148 * 8);
  k_pcg_xfix<<<(int)std::max<long long>(nb, 1), 256, 0, c->stream>>>(c->pk[1], c->pk[2], c->cgs + 3, c->n2);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- host helpers
// symmetric eigen-decomposition by cyclic Jacobi rotations (n <= 8): A = V diag(w) V^T, A destroyed
static void jacobi_eig(int n, double* A, double* V, double* w) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? dg : off) += A[i * n + j] * A[i * n + j];
    if (off <= 1e-32 * dg) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = cs * akp - sn * akq;
          A[k * n + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = cs * apk - sn * aqk;
          A[q * n + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = cs * vkp - sn * vkq;
          V[k * n + q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// generalised symmetric problem A s = lam M s (M SPD), S^T M S = I; S row-major [a][i] (a nodal, i mode)
static bool gen_eig(int n, const double* A, const double* M, double* S, double* lam) {
  double Lc[64], Li[64], C[64], T[64], V[64];
  for (int i = 0; i < n * n; ++i) Lc[i] = 0.0;
  for (int j = 0; j < n; ++j) {                         // Cholesky M = Lc Lc^T
    double d = M[j * n + j];
    for (int k = 0; k < j; ++k) d -= Lc[j * n + k] * Lc[j * n + k];
    if (!(d > 0)) return false;
    Lc[j * n + j] = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = M[i * n + j];
      for (int k = 0; k < j; ++k) s -= Lc[i * n + k] * Lc[j * n + k];
      Lc[i * n + j] = s / Lc[j * n + j];
    }
  }
  for (int c = 0; c < n; ++c)                           // Li = Lc^-1 (lower triangular)
    for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= Lc[i * n + k] * Li[k * n + c];
      Li[i * n + c] = s / Lc[i * n + i];
    }
  for (int i = 0; i < n; ++i)                           // T = Li A
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Li[i * n + k] * A[k * n + j];
      T[i * n + j] = s;
    }
  for (int i = 0; i < n; ++i)                           // C = T Li^T
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += T[i * n + k] * Li[j * n + k];
      C[i * n + j] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) C[i * n + j] = C[j * n + i] = 0.5 * (C[i * n + j] + C[j * n + i]);
  jacobi_eig(n, C, V, lam);
  for (int a = 0; a < n; ++a)                           // S = Li^T V
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Li[k * n + a] * V[k * n + i];
      S[a * n + i] = s;
    }
  return true;
}

// dense SPD inverse by Cholesky, in place; false if not positive definite.  O(n^3) scalar code: used up to n = 512, larger
// aggregate operators (multi-rank) go through LAPACK dpotrf/dpotri (host_krylov.cpp nsb_lapack_spd_inverse)
static bool spd_inverse(int n, std::vector<double>& A) {
  std::vector<double> Lc((size_t)n * n, 0.0), Li((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= Lc[(size_t)j * n + k] * Lc[(size_t)j * n + k];
    if (!(d > 0)) return false;
    Lc[(size_t)j * n + j] = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      const double *li = &Lc[(size_t)i * n], *lj = &Lc[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= li[k] * lj[k];
      Lc[(size_t)i * n + j] = s / Lc[(size_t)j * n + j];
    }
  }
  // Li = Lc^-1 stored transposed (LiT[c][i] = Li[i][c]) for unit-stride inner loops
  for (int c = 0; c < n; ++c) {
    double* col = &Li[(size_t)c * n];
    for (int i = c; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      const double* li = &Lc[(size_t)i * n];
      for (int k = c; k < i; ++k) s -= li[k] * col[k];
      col[i] = s / li[i];
    }
  }
  for (int i = 0; i < n; ++i)                           // A^-1 = Li^T Li : (i,j) = sum_k Li[k][i] Li[k][j] = sum_k LiT[i][k] LiT[j][k]
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
      const double *a = &Li[(size_t)i * n], *b = &Li[(size_t)j * n];
      for (int k = i; k < n; ++k) s += a[k] * b[k];
      A[(size_t)i * n + j] = A[(size_t)j * n + i] = s;
    }
  return true;
}

// recursive coordinate bisection of element centroids; ties by element index
static void rcb(const std::vector<double>& cent, int D, std::vector<int>& idx, int lo, int hi, int first, int ng,
                std::vector<int>& out) {
  if (ng == 1 || hi - lo <= 1) {
    for (int i = lo; i < hi; ++i) out[idx[i]] = first;
    return;
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = lo; i < hi; ++i)
    for (int d = 0; d < D; ++d) {
      mn[d] = std::min(mn[d], cent[(size_t)idx[i] * D + d]);
      mx[d] = std::max(mx[d], cent[(size_t)idx[i] * D + d]);
    }
  int dm = 0;
  for (int d = 1; d < D; ++d)
    if (mx[d] - mn[d] > mx[dm] - mn[dm]) dm = d;
  std::sort(idx.begin() + lo, idx.begin() + hi, [&](int a, int b) {
    const double ca = cent[(size_t)a * D + dm], cb = cent[(size_t)b * D + dm];
    return ca != cb ? ca < cb : a < b;
  });
  const int nl = ng / 2;
  const int cut = lo + (int)(((long long)(hi - lo) * nl) / ng);
  rcb(cent, D, idx, lo, cut, first, nl, out);
  rcb(cent, D, idx, cut, hi, first + nl, ng - nl, out);
}

// Greedy distance-2 colouring of the vertex graph given by the corner ids of all elements (NK per element): two vertices
// are adjacent when they share an element; vertices of one colour are at graph distance >= 3, so the supports of E P e_v
// (elements touching an element touching v) do not meet the supports of the other hats of that colour.  Vertices are
// visited in ascending global id => every rank that holds the same id list computes the same colouring.
static void pm_colour_vertices(const std::vector<long long>& gv, int NK, std::vector<long long>& guv, std::vector<int>& gcol,
                               int& ncol) {
  guv = gv;
  std::sort(guv.begin(), guv.end());
  guv.erase(std::unique(guv.begin(), guv.end()), guv.end());
  const int gnv = (int)guv.size();
  const size_t gnel = gv.size() / NK;
  std::vector<int> gvid(gv.size());
  for (size_t i = 0; i < gv.size(); ++i) gvid[i] = (int)(std::lower_bound(guv.begin(), guv.end(), gv[i]) - guv.begin());
  std::vector<std::vector<int>> adj(gnv);
  for (size_t e = 0; e < gnel; ++e)
    for (int a = 0; a < NK; ++a)
      for (int b = 0; b < NK; ++b) adj[gvid[e * NK + a]].push_back(gvid[e * NK + b]);
  for (auto& l : adj) { std::sort(l.begin(), l.end()); l.erase(std::unique(l.begin(), l.end()), l.end()); }
  gcol.assign(gnv, -1);
  std::vector<int> stamp;
  ncol = 0;
  for (int v = 0; v < gnv; ++v) {
    stamp.assign(ncol + 1, 0);
    for (int u : adj[v])
      for (int t : adj[u])
        if (gcol[t] >= 0) stamp[gcol[t]] = 1;
    int cc = 0;
    while (cc < ncol && stamp[cc]) ++cc;
    gcol[v] = cc;
    if (cc == ncol) ++ncol;
  }
}

// FDM factors of one element direction: A = BD diag(wi) BD^T, M = BJ diag(wi) BJ^T (lx2 x lx2) with wi = 1/w inside and the two
// given end weights; generalised eigenproblem A s = lam M s, S^T M S = I (S row-major [node][mode]).
static bool pm_fdm_1d(const ConstMats& cm, int L1, double w_first, double w_last, double* S, double* lam) {
  const int L2 = L1 - 2;
  double wi[16], A[64], M[64];
  for (int l = 0; l < L1; ++l) wi[l] = 1.0 / cm.w1[l];
  wi[0] = w_first; wi[L1 - 1] = w_last;
  for (int i = 0; i < L2; ++i)
    for (int j = 0; j < L2; ++j) {
      double sa = 0.0, sm = 0.0;
      for (int l = 0; l < L1; ++l) {
        sa += cm.w2[i] * cm.D12[i * L1 + l] * wi[l] * cm.w2[j] * cm.D12[j * L1 + l];
        sm += cm.w2[i] * cm.J12[i * L1 + l] * wi[l] * cm.w2[j] * cm.J12[j * L1 + l];
      }
      A[i * L2 + j] = sa; M[i * L2 + j] = sm;
    }
  return gen_eig(L2, A, M, S, lam);
}

// ---- host-only entry points (no CUDA): the set-up logic of the preconditioner for CPU tests (declared in nekstab_b200.h)
extern "C" int nsb_pm_host_aggregates(int ldim, int nel, const double* cent, int nagg, int* agg_out) {
  if (ldim < 2 || ldim > 3 || nel <= 0 || nagg <= 0) { nsb_set_error("nsb_pm_host_aggregates: bad arguments"); return 1; }
  std::vector<double> c(cent, cent + (size_t)nel * ldim);
  std::vector<int> idx(nel), out(nel, 0);
  std::iota(idx.begin(), idx.end(), 0);
  rcb(c, ldim, idx, 0, nel, 0, std::min(nagg, nel), out);
  memcpy(agg_out, out.data(), sizeof(int) * nel);
  return 0;
}
extern "C" int nsb_pm_host_colouring(int nel, int nk, const long long* vglo, int* colour_out, int* ncolours) {
  if (nel <= 0 || (nk != 4 && nk != 8)) { nsb_set_error("nsb_pm_host_colouring: bad arguments"); return 1; }
  std::vector<long long> gv(vglo, vglo + (size_t)nel * nk), guv;
  std::vector<int> gcol;
  int ncol = 0;
  pm_colour_vertices(gv, nk, guv, gcol, ncol);
  for (size_t i = 0; i < gv.size(); ++i) colour_out[i] = gcol[(int)(std::lower_bound(guv.begin(), guv.end(), gv[i]) - guv.begin())];
  *ncolours = ncol;
  return 0;
}
extern "C" int nsb_pm_host_fdm_1d(int lx1, double w_first, double w_last, double* S, double* lam) {
  if (lx1 != 4 && lx1 != 6 && lx1 != 8) { nsb_set_error("nsb_pm_host_fdm_1d: lx1 must be 4, 6 or 8"); return 1; }
  ConstMats cm;
  sem_build_constmats(lx1, lx1 - 2, 3 * lx1 / 2, &cm);
  if (!pm_fdm_1d(cm, lx1, w_first, w_last, S, lam)) { nsb_set_error("nsb_pm_host_fdm_1d: mass matrix not positive definite"); return 1; }
  return 0;
}
extern "C" int nsb_pm_host_spd_inverse(int n, double* A) {
  if (n <= 0) { nsb_set_error("nsb_pm_host_spd_inverse: n must be positive"); return 1; }
  std::vector<double> a(A, A + (size_t)n * n);
  if (!spd_inverse(n, a)) { nsb_set_error("nsb_pm_host_spd_inverse: matrix not positive definite"); return 1; }
  memcpy(A, a.data(), sizeof(double) * (size_t)n * n);
  return 0;
}

template <class T>
static int pm_upload(T** dptr, const std::vector<T>& h) {
  const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  if (*dptr) cudaFree(*dptr);
  NSB_CUDA(cudaMalloc((void**)dptr, bytes));
  NSB_CUDA(cudaMemset(*dptr, 0, bytes));
  if (!h.empty()) NSB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
static int pm_download(Ctx* c, std::vector<double>& h, const double* d, long long n) {
  h.resize(n);
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  NSB_CUDA(cudaMemcpy(h.data(), d, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

#define PM_DISPATCH(c, CALL)                                                                                   \
  do {                                                                                                         \
    const int key_ = (c)->ldim * 10 + (c)->lx2;                                                                \
    switch (key_) {                                                                                            \
      case 22: { constexpr int D = 2, L = 2; CALL; } break;                                                    \
      case 24: { constexpr int D = 2, L = 4; CALL; } break;                                                    \
      case 26: { constexpr int D = 2, L = 6; CALL; } break;                                                    \
      case 32: { constexpr int D = 3, L = 2; CALL; } break;                                                    \
      case 34: { constexpr int D = 3, L = 4; CALL; } break;                                                    \
      case 36: { constexpr int D = 3, L = 6; CALL; } break;                                                    \
      default: nsb_set_error("pmg: unsupported (ldim, lx2) = (%d, %d)", (c)->ldim, (c)->lx2); return 1;        \
    }                                                                                                          \
    nsb_count_launch();                                                                                        \
    NSB_CUDA(cudaGetLastError());                                                                              \
  } while (0)

void pm_free(PMG& m) {
  cudaFree(m.S); cudaFree(m.lam); cudaFree(m.vid); cudaFree(m.voff); cudaFree(m.vent); cudaFree(m.d1inv); cudaFree(m.agg);
  cudaFree(m.aoff); cudaFree(m.aent); cudaFree(m.A2inv); cudaFree(m.rc); cudaFree(m.rc0); cudaFree(m.xc); cudaFree(m.hat); cudaFree(m.xv); cudaFree(m.ra); cudaFree(m.x2);
  m = PMG();
}

// ---------------------------------------------------------------------------------------------- runtime pieces
static int pm_restrict(Ctx* c, PMG& m, const double* r, const CGState* skip) {
  PM_DISPATCH(c, (k_pm_restrict<D, L><<<std::min((c->nel + 7) / 8, 148 * 8), 256, 0, c->stream>>>(r, m.rc, c->nel, skip)));
  return 0;
}
static int pm_coarse(Ctx* c, PMG& m, const double* d1inv, int vmode, int amode, const CGState* skip) {
  const int vthreads = ((std::max(m.nv, m.nagg) + 31) / 32) * 32;
  const long long nthr = vthreads + 32LL * m.nagg_loc;
  k_pm_coarse<<<(int)((nthr + 127) / 128), 128, 0, c->stream>>>(m.nv, m.voff, m.vent, d1inv, m.rc, m.xv, m.nagg, m.nagg_loc,
                                                               m.agg_first, m.aoff, m.aent, (c->ldim == 3) ? 8 : 4, m.ra, vmode,
                                                               amode, skip);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
// vertex values xv = d1inv * (sum over all copies on all ranks) and aggregate sums ra (local entries) from the corner sums rc
static int pm_coarse_levels(Ctx* c, PMG& m, const double* d1inv, const CGState* skip) {
  if (c->nranks == 1) return pm_coarse(c, m, d1inv, 0, 1, skip);
  NSB_TRY(pm_coarse(c, m, nullptr, 2, 1, skip));                                    // aggregate sums need the un-assembled rc
  NSB_TRY(gs_dssum_map(c, c->gsv, c->p2pv, m.rc, 1, 0, skip));                      // vertex sums across elements and ranks
  return pm_coarse(c, m, d1inv, 1, 0, skip);
}
static int pm_gemv(Ctx* c, PMG& m, const CGState* skip) {
  k_pm_gemv<<<(m.nagg * 32 + 127) / 128, 128, 0, c->stream>>>(m.nagg, m.A2inv, m.ra, m.x2, skip);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
static int pm_prolong(Ctx* c, PMG& m, const double* xv, const double* x2, double* out) {
  const long long n = c->n2;
  PM_DISPATCH(c, (k_pm_prolong<D, L><<<(int)((n + 255) / 256), 256, 0, c->stream>>>(xv, m.vid, x2, m.agg, out, c->nel)));
  return 0;
}
// pout = E pin for mask set `set`; uses wk[2]
static int pm_apply_E(Ctx* c, int set, const double* pin, double* pout) {
  NSB_TRY(ek_gradt(c, pin, c->wk[2]));
  NSB_TRY(gs_dssum(c, c->wk[2], c->ldim, c->n, nullptr));
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(vk_mul(c, c->wk[2] + d * c->n, c->mbinv[set][d], c->n));
  NSB_TRY(ek_div(c, c->wk[2], nullptr, pout, 1.0));
  return 0;
}

// z = M^-1 r.  mode 0: plain operator; 1: inside the pressure CG (skips when converged, updates rtz1/beta)
int pm_apply(Ctx* c, int set, const double* r, double* z, int mode, int prof_slot) {
  PMG& m = c->pmg[(set && c->has_adj_masks) ? 1 : 0];      // without separate adjoint masks both problems share one E
  if (!m.ready) { nsb_set_error("pmg: preconditioner not set up"); return 1; }
  if (mode == 0 && c->pc_kind == 1 && c->pcg_fused && c->ldim == 3) {
    // operator-level entry (nsb_op_pc_apply) through the kernels of the fused CG tail: the init form of pm_pcg_tail gives the
    // element-block part in pz and the coarse-level values xv, x2; their interpolation is added with the set-up's prolongation.
    NSB_TRY(vk_copy(c, c->pk[0], r, c->n2));
    NSB_TRY(pm_pcg_tail(c, set, 1, 0));
    NSB_TRY(pm_prolong(c, m, m.xv, m.x2, c->pk[4]));
    if (z != c->pz) NSB_TRY(vk_copy(c, z, c->pz, c->n2));
    return vk_axpy(c, z, 1.0, c->pk[4], c->n2);
  }
  CGState* sp = c->cgs + 3;
  const CGState* skip = mode ? sp : nullptr;
  NSB_TRY(pm_restrict(c, m, r, skip));
  if (prof_slot > 0) cudaEventRecord(c->prof_ev[prof_slot], c->stream);
  NSB_TRY(pm_coarse_levels(c, m, m.d1inv, skip));
  if (c->nranks > 1) NSB_TRY(vk_allreduce_sum(c, m.ra, m.nagg));
  NSB_TRY(pm_gemv(c, m, skip));
  if (prof_slot > 0) cudaEventRecord(c->prof_ev[prof_slot + 1], c->stream);
  const int kmode = mode ? (c->nranks == 1 ? 1 : 2) : 0;
  PM_DISPATCH(c, (k_pm_apply2<D, L><<<(c->nel + Pm2<D, L>::EPB - 1) / Pm2<D, L>::EPB, Pm2<D, L>::NT, 0, c->stream>>>(
                     r, z, c->nel, m.S, m.lam, m.xv, m.vid, m.x2, m.agg, sp, kmode, c->red_part, c->red_count, c->red_out)));
  if (mode && c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, sp, 1, 5));
  return 0;
}

// The fused tail of one pressure-CG iteration (3-D): CG update + norm + restriction + element-block solves (k_pcg_fused), the vertex
// and aggregate levels with their shares of z.r (k_pm_coarse_dot), the dense aggregate solve and the CG scalar update
// (k_pm_gemv_fin).  Multi-rank: ONE vertex halo exchange (assembles rc) and ONE all-reduce (aggregate sums + the three scalars)
// instead of the former norm / vertex halo / aggregate / z.r sequence.  The three scalars live behind the aggregate sums: ra[nagg..nagg+2].
int pm_pcg_tail(Ctx* c, int set, int init, int prof_slot) {
  PMG& m = c->pmg[(set && c->has_adj_masks) ? 1 : 0];
  if (!m.ready) { nsb_set_error("pmg: preconditioner not set up"); return 1; }
  CGState* sp = c->cgs + 3;
  double* sc = m.ra + m.nagg;
  const bool multi = c->nranks > 1;
  if (c->ldim == 3 && c->lx2 == 6) {                       // persistent, TMA-pipelined (lx1 = 8); other orders: one group per CTA
    constexpr int L = 6;
    static bool attr = false;
    if (!attr) {
      NSB_CUDA(cudaFuncSetAttribute(k_pcg_fused_p<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PfP<L>::smem));
      NSB_CUDA(cudaFuncSetAttribute(k_pcg_fused_p<L>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      attr = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    PfArgs a{c->pk[0], c->pk[3], c->bm2inv, m.S, m.lam, c->pk[1], c->pk[2], c->pz, m.rc, multi ? m.rc0 : nullptr, c->nel, init, 0.0};
    const int ngroups = (c->nel + PfP<L>::EPB - 1) / PfP<L>::EPB;
    k_pcg_fused_p<L><<<std::min(ngroups, 2 * sms), PfP<L>::NT, PfP<L>::smem, c->stream>>>(a, sp, c->red_part, c->red_count, sc);
    nsb_count_launch();
    NSB_CUDA(cudaGetLastError());
  } else {
    PM_DISPATCH(c, (k_pcg_fused<D, L><<<(c->nel + Pm2<D, L>::EPB - 1) / Pm2<D, L>::EPB, Pm2<D, L>::NT, 0, c->stream>>>(
                       c->pk[0], c->pk[1], c->pk[2], c->pk[3], c->bm2inv, c->pz, m.rc, multi ? m.rc0 : nullptr, c->nel, m.S, m.lam, sp, init,
                       c->red_part, c->red_count, sc)));
  }
  if (prof_slot > 0) cudaEventRecord(c->prof_ev[prof_slot], c->stream);
  // vertex sums across elements and ranks (init: the CG state may be stale -- operator-level calls -- so nothing may be skipped)
  if (multi) NSB_TRY(gs_dssum_map(c, c->gsv, c->p2pv, m.rc, 1, 0, init ? nullptr : sp));
  const int vthreads = ((std::max(m.nv, m.nagg) + 31) / 32) * 32;
  const long long nthr = vthreads + 32LL * m.nagg_loc;
  k_pm_coarse_dot<<<(int)((nthr + 127) / 128), 128, 0, c->stream>>>(m.nv, m.voff, m.vent, m.d1inv, m.rc, multi ? m.rc0 : m.rc, m.xv, m.nagg,
                                                                   m.nagg_loc, m.agg_first, m.aoff, m.aent, (c->ldim == 3) ? 8 : 4, m.ra,
                                                                   multi ? 1 : 0, sp, init, c->red_part, c->red_count, sc + 2);
  nsb_count_launch();
  if (multi) NSB_TRY(vk_allreduce_sum(c, m.ra, m.nagg + 3));
  k_pm_gemv_fin<<<(m.nagg * 32 + 127) / 128, 128, 0, c->stream>>>(m.nagg, m.A2inv, m.ra, m.x2, sc, sp, init, c->red_part, c->red_count,
                                                                c->red_out);
  {
    const int nk = (c->ldim == 3) ? 8 : 4, tot = c->nel * (nk + 1);
    k_pm_corner_values<<<(tot + 255) / 256, 256, 0, c->stream>>>(c->nel, nk, m.xv, m.vid, m.x2, m.agg, m.xc, init ? nullptr : sp);
  }
  nsb_count_launch(2);
  NSB_CUDA(cudaGetLastError());
  if (prof_slot > 0) cudaEventRecord(c->prof_ev[prof_slot + 1], c->stream);
  return 0;
}

// all ranks' corner-id lists, concatenated in rank order (NCCL all-gather of counts, then of padded data)
static int pm_allgather_ids(Ctx* c, const std::vector<long long>& mine, std::vector<long long>& all) {
  const int R = c->nranks;
  std::vector<long long> cnts(R, 0);
  long long mycnt = (long long)mine.size();
  long long* d_cnt = nullptr;
  NSB_CUDA(cudaMalloc(&d_cnt, sizeof(long long) * (R + 1)));
  NSB_CUDA(cudaMemcpy(d_cnt + R, &mycnt, sizeof(long long), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllGather(d_cnt + R, d_cnt, 1, ncclInt64, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  NSB_CUDA(cudaMemcpy(cnts.data(), d_cnt, sizeof(long long) * R, cudaMemcpyDeviceToHost));
  cudaFree(d_cnt);
  const long long mx = std::max<long long>(1, *std::max_element(cnts.begin(), cnts.end()));
  long long *d_my = nullptr, *d_all = nullptr;
  NSB_CUDA(cudaMalloc(&d_my, sizeof(long long) * mx));
  NSB_CUDA(cudaMalloc(&d_all, sizeof(long long) * mx * R));
  NSB_CUDA(cudaMemset(d_my, 0, sizeof(long long) * mx));
  NSB_CUDA(cudaMemcpy(d_my, mine.data(), sizeof(long long) * mine.size(), cudaMemcpyHostToDevice));
  NSB_NCCL(ncclAllGather(d_my, d_all, mx, ncclInt64, c->comm, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<long long> padded((size_t)mx * R);
  NSB_CUDA(cudaMemcpy(padded.data(), d_all, sizeof(long long) * padded.size(), cudaMemcpyDeviceToHost));
  cudaFree(d_my); cudaFree(d_all);
  all.clear();
  for (int r = 0; r < R; ++r) all.insert(all.end(), padded.begin() + (size_t)r * mx, padded.begin() + (size_t)r * mx + cnts[r]);
  return 0;
}

// ---------------------------------------------------------------------------------------------- setup
int pm_setup(Ctx* c, int set, int nagg_req) {
  PMG& m = c->pmg[set];
  pm_free(m);
  const int D = c->ldim, L1 = c->lx1, L2 = c->lx2, nel = c->nel, np1 = c->np1, NK = (D == 3) ? 8 : 4;
  const int N = L1 - 1, mid = L1 / 2;
  if (c->nranks > 1 && !c->gsv_ready) {    // gather-scatter over the element-vertex mesh: entries (e, corner), 2^ldim per element
    NSB_TRY(gs_build(c, c->gsv, c->p2pv, (long long)nel * NK, 2, NK, c->vglo.data()));
    c->gsv_ready = true;
  }
  if (c->vglo.size() != (size_t)nel * NK) { nsb_set_error("pmg: vertex ids missing"); return 1; }
  // ---- host copies of what the FDM factors are built from
  std::vector<double> X[3], binv, bm1, mk[3];
  for (int d = 0; d < D; ++d) NSB_TRY(pm_download(c, X[d], c->xyz[d], c->n));
  NSB_TRY(pm_download(c, binv, c->binv, c->n));
  NSB_TRY(pm_download(c, bm1, c->bm1, c->n));
  for (int d = 0; d < D; ++d) NSB_TRY(pm_download(c, mk[d], c->mask[set][d], c->n));
  // GL points and hat functions
  double zg[16], wg[16], l01[2][8];
  sem_zwgl(L2, zg, wg);
  for (int i = 0; i < L2; ++i) { l01[0][i] = 0.5 * (1.0 - zg[i]); l01[1][i] = 0.5 * (1.0 + zg[i]); }
  NSB_CUDA(cudaMemcpyToSymbol(pm_l, l01, sizeof(l01)));
  auto node = [&](int e, int i, int j, int k) { return (size_t)e * np1 + (size_t)(k * L1 + j) * L1 + i; };
  std::vector<double> hS((size_t)nel * D * L2 * L2), hlam((size_t)nel * D * L2), cent((size_t)nel * D);
  const int str1[3] = {1, L1, L1 * L1};
  for (int e = 0; e < nel; ++e) {
    double h[3] = {1, 1, 1};
    // centroid
    for (int d = 0; d < D; ++d) {
      double s = 0.0;
      for (int p = 0; p < np1; ++p) s += X[d][(size_t)e * np1 + p];
      cent[(size_t)e * D + d] = s / np1;
    }
    for (int dr = 0; dr < D; ++dr) {
      // face centroids: plain mean over the nodes with index 0 / N along direction dr
      double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      int cnt = 0;
      for (int p = 0; p < np1; ++p) {
        const int id = (p / str1[dr]) % L1;
        if (id != 0) continue;
        ++cnt;
        for (int d = 0; d < D; ++d) { lo[d] += X[d][(size_t)e * np1 + p]; hi[d] += X[d][(size_t)e * np1 + p + (size_t)N * str1[dr]]; }
      }
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (hi[d] - lo[d]) * (hi[d] - lo[d]) / ((double)cnt * cnt);
      h[dr] = std::sqrt(s);
    }
    for (int dr = 0; dr < D; ++dr) {
      int ijk[3] = {mid, mid, D == 3 ? mid : 0};
      double wend[2];
      for (int end = 0; end < 2; ++end) {
        ijk[dr] = end ? N : 0;
        const size_t g = node(e, ijk[0], ijk[1], ijk[2]);
        const double ml = 1.0 / (binv[g] * bm1[g]);
        double kk = 1.0;
        for (int d = 0; d < D; ++d) kk *= mk[d][g];
        wend[end] = kk / (c->cm.w1[end ? N : 0] * ml);
      }
      double Sm[64], lm[8];
      if (!pm_fdm_1d(c->cm, L1, wend[0], wend[1], Sm, lm)) { nsb_set_error("pmg: FDM mass matrix of element %d not positive definite", e); return 1; }
      double cd = (D == 3) ? 0.5 : 1.0;
      for (int d = 0; d < D; ++d) cd *= (d == dr) ? 1.0 / h[d] : h[d];
      for (int i = 0; i < L2 * L2; ++i) hS[((size_t)e * D + dr) * L2 * L2 + i] = Sm[i];
      for (int i = 0; i < L2; ++i) hlam[((size_t)e * D + dr) * L2 + i] = cd * lm[i];
    }
  }
  NSB_TRY(pm_upload(&m.S, hS));
  NSB_TRY(pm_upload(&m.lam, hlam));
  // ---- vertices: local numbering (ascending global id), CSR of (element, corner) entries
  std::vector<long long> uv(c->vglo);
  std::sort(uv.begin(), uv.end());
  uv.erase(std::unique(uv.begin(), uv.end()), uv.end());
  m.nv = (int)uv.size();
  std::vector<int> vid((size_t)nel * NK), voff(m.nv + 1, 0), vent((size_t)nel * NK);
  for (size_t i = 0; i < vid.size(); ++i) {
    vid[i] = (int)(std::lower_bound(uv.begin(), uv.end(), c->vglo[i]) - uv.begin());
    voff[vid[i] + 1]++;
  }
  for (int v = 0; v < m.nv; ++v) voff[v + 1] += voff[v];
  {
    std::vector<int> fill(voff.begin(), voff.end() - 1);
    for (size_t i = 0; i < vid.size(); ++i) vent[fill[vid[i]]++] = (int)i;
  }
  NSB_TRY(pm_upload(&m.vid, vid));
  NSB_TRY(pm_upload(&m.voff, voff));
  NSB_TRY(pm_upload(&m.vent, vent));
  // ---- aggregates (recursive coordinate bisection of the local elements)
  // aggregates per rank: nelv/32, at most 512 per rank and 4096 in total (r1: a fixed total of 512 made the aggregates grow with the
  // rank count under weak scaling: 55 -> 61 iterations per step at 2-8 GPUs)
  int nagg = nagg_req > 0 ? std::max(1, nagg_req / c->nranks) : std::max(1, nel / 32);
  nagg = std::max(1, std::min(std::min(nagg, nel), std::min(512, 4096 / c->nranks)));
  m.nagg_loc = nagg; m.nagg = nagg; m.agg_first = 0;
  if (c->nranks > 1) {      // global aggregate ids: rank-ordered blocks
    std::vector<double> cnt(c->nranks, 0.0);
    cnt[c->rank] = nagg;
    NSB_CUDA(cudaMemcpyAsync(c->hbuf, cnt.data(), c->nranks * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NSB_TRY(vk_allreduce_sum(c, c->hbuf, c->nranks));
    NSB_TRY(pm_download(c, cnt, c->hbuf, c->nranks));
    m.nagg = 0;
    for (int r = 0; r < c->nranks; ++r) { if (r == c->rank) m.agg_first = m.nagg; m.nagg += (int)std::lround(cnt[r]); }
  }
  std::vector<int> agg(nel), idx(nel);
  std::iota(idx.begin(), idx.end(), 0);
  rcb(cent, D, idx, 0, nel, 0, nagg, agg);
  std::vector<int> aoff(nagg + 1, 0), aent(nel);
  for (int e = 0; e < nel; ++e) aoff[agg[e] + 1]++;
  for (int a = 0; a < nagg; ++a) aoff[a + 1] += aoff[a];
  {
    std::vector<int> fill(aoff.begin(), aoff.end() - 1);
    for (int e = 0; e < nel; ++e) aent[fill[agg[e]]++] = e;
  }
  m.h_agg = agg;
  for (int e = 0; e < nel; ++e) agg[e] += m.agg_first;
  NSB_TRY(pm_upload(&m.agg, agg));
  NSB_TRY(pm_upload(&m.aoff, aoff));
  NSB_TRY(pm_upload(&m.aent, aent));
  NSB_TRY(pm_upload(&m.rc, std::vector<double>((size_t)nel * NK, 0.0)));
  NSB_TRY(pm_upload(&m.xv, std::vector<double>(m.nv, 0.0)));
  NSB_TRY(pm_upload(&m.ra, std::vector<double>(m.nagg + 3, 0.0)));      // + |r|^2, zloc.r, sum xv rv (pm_pcg_tail)
  NSB_TRY(pm_upload(&m.rc0, std::vector<double>((size_t)nel * NK, 0.0)));
  NSB_TRY(pm_upload(&m.xc, std::vector<double>((size_t)nel * (NK + 1), 0.0)));
  NSB_TRY(pm_upload(&m.hat, std::vector<double>(c->cm.hat1, c->cm.hat1 + 12)));
  NSB_TRY(pm_upload(&m.x2, std::vector<double>(m.nagg, 0.0)));
  // ---- distance-2 colouring of the GLOBAL vertex graph (adjacent = share an element): every rank gathers the corner ids
  //      of all elements and runs the same greedy colouring, so that the probing below is consistent across ranks
  std::vector<int> col(m.nv);
  int ncol = 0;
  if (c->pc_ncol > 0 && (int)c->pc_col.size() == m.nv) {
    col = c->pc_col; ncol = c->pc_ncol;           // same mesh, other mask set: reuse
  } else {
  std::vector<long long> gv;                      // corner ids of all elements of all ranks
  if (c->nranks == 1) gv = c->vglo;
  else NSB_TRY(pm_allgather_ids(c, c->vglo, gv));
  std::vector<long long> guv;
  std::vector<int> gcol;
  pm_colour_vertices(gv, NK, guv, gcol, ncol);
  for (int v = 0; v < m.nv; ++v) col[v] = gcol[(int)(std::lower_bound(guv.begin(), guv.end(), uv[v]) - guv.begin())];
  c->pc_col = col; c->pc_ncol = ncol;
  }
  m.ncolours = ncol;
  // ---- diag(P^T E P) by probing, one E application per colour
  std::vector<double> d1(m.nv, 0.0), xvh(m.nv), rvh;
  for (int cc = 0; cc < ncol; ++cc) {
    for (int v = 0; v < m.nv; ++v) xvh[v] = (col[v] == cc) ? 1.0 : 0.0;
    NSB_CUDA(cudaMemcpyAsync(m.xv, xvh.data(), m.nv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NSB_TRY(pm_prolong(c, m, m.xv, nullptr, c->pk[2]));
    NSB_TRY(pm_apply_E(c, set, c->pk[2], c->pk[3]));
    NSB_TRY(pm_restrict(c, m, c->pk[3], nullptr));
    NSB_TRY(pm_coarse_levels(c, m, nullptr, nullptr));
    NSB_TRY(pm_download(c, rvh, m.xv, m.nv));
    for (int v = 0; v < m.nv; ++v)
      if (col[v] == cc) d1[v] = rvh[v];
  }
  m.h_d1 = d1;
  for (int v = 0; v < m.nv; ++v) {
    if (!(d1[v] > 0)) { nsb_set_error("pmg: non-positive Q1 diagonal %g at vertex %d", d1[v], v); return 1; }
    d1[v] = 1.0 / d1[v];
  }
  NSB_TRY(pm_upload(&m.d1inv, d1));
  // ---- A2 = Pa^T E Pa by probing, one E application per aggregate
  std::vector<double> A2((size_t)m.nagg * m.nagg, 0.0), x2h(m.nagg, 0.0), rah;
  for (int a = 0; a < m.nagg; ++a) {
    std::fill(x2h.begin(), x2h.end(), 0.0);
    x2h[a] = 1.0;
    NSB_CUDA(cudaMemcpyAsync(m.x2, x2h.data(), m.nagg * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NSB_TRY(pm_prolong(c, m, nullptr, m.x2, c->pk[2]));
    NSB_TRY(pm_apply_E(c, set, c->pk[2], c->pk[3]));
    NSB_TRY(pm_restrict(c, m, c->pk[3], nullptr));
    NSB_TRY(pm_coarse(c, m, nullptr, 2, 1, nullptr));
    NSB_TRY(pm_download(c, rah, m.ra, m.nagg));
    for (int b = 0; b < m.nagg; ++b) A2[(size_t)b * m.nagg + a] = rah[b];      // rows of other ranks' aggregates are zero here
  }
  if (c->nranks > 1) {      // sum the row blocks of all ranks
    double* dA = nullptr;
    NSB_CUDA(cudaMalloc(&dA, A2.size() * sizeof(double)));
    NSB_CUDA(cudaMemcpy(dA, A2.data(), A2.size() * sizeof(double), cudaMemcpyHostToDevice));
    NSB_TRY(vk_allreduce_sum(c, dA, (int)A2.size()));
    NSB_TRY(pm_download(c, A2, dA, (long long)A2.size()));
    cudaFree(dA);
  }
  double tr = 0.0;
  for (int a = 0; a < m.nagg; ++a) {
    tr += A2[(size_t)a * m.nagg + a];
    for (int b = a + 1; b < m.nagg; ++b)
      A2[(size_t)a * m.nagg + b] = A2[(size_t)b * m.nagg + a] = 0.5 * (A2[(size_t)a * m.nagg + b] + A2[(size_t)b * m.nagg + a]);
  }
  if (c->ifvcor[set]) {        // E 1 = 0  =>  A2 1 = 0: shift the null vector (the CG residual stays orthogonal to 1)
    const double sh = tr / ((double)m.nagg * m.nagg);
    for (auto& v : A2) v += sh;
  }
  if (c->ifvcor[set] && m.nagg == 1) {
    A2[0] = 0.0;
  } else if (!(m.nagg > 512 ? nsb_lapack_spd_inverse(m.nagg, A2.data()) == 0 : spd_inverse(m.nagg, A2))) {
    nsb_set_error("pmg: aggregate operator not positive definite");
    return 1;
  }
  m.h_A2inv = A2;
  NSB_TRY(pm_upload(&m.A2inv, A2));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  m.ready = true;
  return 0;
}
