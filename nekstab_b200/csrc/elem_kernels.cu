// elem_kernels.cu -- element-local tensor-product kernels of the linearised Navier-Stokes step, FP64, sm_100a.
//
// Each kernel restates one Nek5000 routine reached from nekStab's `call nek_advance` (core/matvec.f:222)
// [UPSTREAM, not vendored in the reference]:
//   k_axhelm      hmholtz.f axhelm (+ local_grad3 / local_grad3_t): w = (h1 A + h2 B) u
//   k_gradt       navier1.f opgradt / cdtp : w_c = D_c^T p            (mesh 2 -> mesh 1)
//   k_div         navier1.f opdiv   / multd: q   = sum_c D_c u_c      (mesh 1 -> mesh 2)
//   k_advab       perturb.f advabp / advabp_adjoint through convect.f convect_new (dealiased on GL(lxd))
// plus the fused variants used inside the Jacobi-PCG loops (hmholtz.f cggo; pressure CG on E = D B^-1 D^T).
// One CTA = one element; see elem_common.cuh for the shared-memory/column scheme.
#include "elem_common.cuh"

template <int D, int N>
struct Cfg {
  static constexpr int N2 = N - 2;
  static constexpr int ND = 3 * N / 2;
  static constexpr int NK1 = (D == 3) ? N : 1;
  static constexpr int NK2 = (D == 3) ? N2 : 1;
  static constexpr int NKD = (D == 3) ? ND : 1;
  static constexpr int NP1 = NK1 * N * N;
  static constexpr int NP2 = NK2 * N2 * N2;
  static constexpr int NPD = NKD * ND * ND;
  static constexpr int TPB = (NP1 >= 256) ? 256 : ((NP1 + 31) / 32) * 32;
  static constexpr int TPB_ADV = (D == 3) ? 512 : 128;
  static constexpr int NG = D * (D + 1) / 2;
  using S1 = Shp<NK1, N, N>;
  using S2 = Shp<NK2, N2, N2>;
  using SD = Shp<NKD, ND, ND>;
};

__device__ __forceinline__ int gidx(int D, int i, int j) {  // symmetric G storage: 3D 11,22,33,12,13,23 ; 2D 11,22,12
  if (i == j) return i;
  if (D == 2) return 2;
  int a = i < j ? i : j, b = i < j ? j : i;
  return (a == 0) ? (b == 1 ? 3 : 4) : 5;
}

// --------------------------------------------------------------------------------------------- axhelm
// MODE 0: w = H u                      (nsb_op_axhelm)
// MODE 1: r = b + r - H u              (cresvipp residual, r holds D^T p on entry)
// MODE 2: CG direction update fused:   p = dinv*r + beta*p (stored), w = H p, rho_c partial = sum p*w
template <int D, int N, int MODE>
__global__ void __launch_bounds__(Cfg<D, N>::TPB)
k_axhelm(const double* __restrict__ u, double* __restrict__ w, const double* __restrict__ b,
         const double* __restrict__ G, const double* __restrict__ bm1, const double* __restrict__ dinv,
         double* __restrict__ pdir, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter,
         double* __restrict__ red_out, int finalize, int nfields, long long n, double h1, double h2) {
  using C = Cfg<D, N>;
  using S = typename C::S1;
  constexpr int TPB = C::TPB;
  constexpr int PPT = (C::NP1 + TPB - 1) / TPB;
  __shared__ double su[S::size];
  __shared__ double sr[D][S::size];
  __shared__ double sred[3 * 32];
  const int tid = threadIdx.x;
  const long long e0 = (long long)blockIdx.x * C::NP1;

  double g[PPT][C::NG], bm[PPT];
#pragma unroll
  for (int m = 0; m < PPT; ++m) {
    int p = tid + m * TPB;
    if (p < C::NP1) {
#pragma unroll
      for (int q = 0; q < C::NG; ++q) g[m][q] = G[(long long)q * n + e0 + p];
      bm[m] = bm1[e0 + p];
    }
  }
  double rho[3] = {0.0, 0.0, 0.0};
  for (int f = 0; f < nfields; ++f) {
    bool skip = false;
    double beta = 0.0;
    if (MODE == 2) {
      skip = cgs[f].done != 0;
      beta = cgs[f].beta;
    }
    double uo[PPT];
    if (!skip) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        int p = tid + m * TPB;
        if (p < C::NP1) {
          long long gi = (long long)f * n + e0 + p;
          double v;
          if (MODE == 2) {
            v = dinv[e0 + p] * u[gi] + beta * pdir[gi];   // u = r here
            pdir[gi] = v;
          } else {
            v = u[gi];
          }
          uo[m] = v;
          su[S::lin(p)] = v;
        }
      }
    }
    __syncthreads();
    if (!skip) {
      contract<0, N, N, C::NK1, N, N, false>(su, sr[0], cm.D, tid, TPB);
      contract<1, N, N, C::NK1, N, N, false>(su, sr[1], cm.D, tid, TPB);
      if constexpr (D == 3) contract<2, N, N, C::NK1, N, N, false>(su, sr[2], cm.D, tid, TPB);
    }
    __syncthreads();
    if (!skip) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        int p = tid + m * TPB;
        if (p < C::NP1) {
          int o = S::lin(p);
          double d[D], t[D];
#pragma unroll
          for (int i = 0; i < D; ++i) d[i] = sr[i][o];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(g[m][gidx(D, i, j)], d[j], s);
            t[i] = s;
          }
#pragma unroll
          for (int i = 0; i < D; ++i) sr[i][o] = t[i];
        }
      }
    }
    __syncthreads();
    if (!skip) contract<0, N, N, C::NK1, N, N, false>(sr[0], su, cm.Dt, tid, TPB);
    __syncthreads();
    if (!skip) contract<1, N, N, C::NK1, N, N, true>(sr[1], su, cm.Dt, tid, TPB);
    __syncthreads();
    if constexpr (D == 3) {
      if (!skip) contract<2, N, N, C::NK1, N, N, true>(sr[2], su, cm.Dt, tid, TPB);
      __syncthreads();
    }
    if (!skip) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        int p = tid + m * TPB;
        if (p < C::NP1) {
          long long gi = (long long)f * n + e0 + p;
          double hv = h1 * su[S::lin(p)] + h2 * bm[m] * uo[m];
          if (MODE == 1) {
            w[gi] = b[gi] + w[gi] - hv;
          } else {
            w[gi] = hv;
            if (MODE == 2) rho[f] += uo[m] * hv;
          }
        }
      }
    }
    __syncthreads();
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

// --------------------------------------------------------------------------------------------- gradt
// MODE 0: w_c = D_c^T p
// MODE 1: pressure-CG direction update fused: pd = dinvE*r + beta*pd (stored); w_c = D_c^T pd
template <int D, int N, int MODE>
__global__ void __launch_bounds__(Cfg<D, N>::TPB)
k_gradt(const double* __restrict__ p, double* __restrict__ w, const double* __restrict__ RW2,
        const double* __restrict__ dinvE, double* __restrict__ pdir, const CGState* __restrict__ cgs,
        long long n, long long n2) {
  using C = Cfg<D, N>;
  using S1 = typename C::S1;
  using S2 = typename C::S2;
  constexpr int N2 = C::N2, TPB = C::TPB;
  constexpr int NK2 = C::NK2;
  using SA = Shp<NK2, N2, N>;                    // after the r stage
  using SB = Shp<NK2, N, N>;                     // after the s stage
  __shared__ double sp[S2::size];
  __shared__ double sq[D][S2::size];
  __shared__ double sa[D][SA::size];
  __shared__ double sb[2][SB::size];
  __shared__ double sw[S1::size];
  const int tid = threadIdx.x;
  if (MODE == 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * C::NP2;
  const long long e1 = (long long)blockIdx.x * C::NP1;
  for (int q = tid; q < C::NP2; q += TPB) {
    double v;
    if (MODE == 1) {
      v = dinvE[e2 + q] * p[e2 + q] + cgs->beta * pdir[e2 + q];
      pdir[e2 + q] = v;
    } else {
      v = p[e2 + q];
    }
    sp[S2::lin(q)] = v;
  }
  __syncthreads();
  for (int c = 0; c < D; ++c) {
    for (int q = tid; q < C::NP2; q += TPB) {
      int o = S2::lin(q);
      double v = sp[o];
#pragma unroll
      for (int i = 0; i < D; ++i) sq[i][o] = RW2[(long long)(i * D + c) * n2 + e2 + q] * v;
    }
    __syncthreads();
    // r stage (axis 0): N2 -> N
    contract<0, N, N2, NK2, N2, N2, false>(sq[0], sa[0], cm.D12t, tid, TPB);
    contract<0, N, N2, NK2, N2, N2, false>(sq[1], sa[1], cm.J12t, tid, TPB);
    if constexpr (D == 3) contract<0, N, N2, NK2, N2, N2, false>(sq[2], sa[2], cm.J12t, tid, TPB);
    __syncthreads();
    // s stage (axis 1)
    contract<1, N, N2, NK2, N2, N, false>(sa[0], sb[0], cm.J12t, tid, TPB);
    contract<1, N, N2, NK2, N2, N, true>(sa[1], sb[0], cm.D12t, tid, TPB);
    if constexpr (D == 3) contract<1, N, N2, NK2, N2, N, false>(sa[2], sb[1], cm.J12t, tid, TPB);
    __syncthreads();
    if constexpr (D == 3) {
      contract<2, N, N2, NK2, N, N, false>(sb[0], sw, cm.J12t, tid, TPB);
      contract<2, N, N2, NK2, N, N, true>(sb[1], sw, cm.D12t, tid, TPB);
      __syncthreads();
      for (int q = tid; q < C::NP1; q += TPB) w[(long long)c * n + e1 + q] = sw[S1::lin(q)];
    } else {
      for (int q = tid; q < C::NP1; q += TPB) w[(long long)c * n + e1 + q] = sb[0][SB::lin(q)];
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------------- div
// MODE 0: q = sign * sum_c D_c (scale_c * u_c)        (scale may be null)
// MODE 1: pressure-CG: Ep = sum_c D_c (mbinv_c * w_c), rho partial = sum pd * Ep
template <int D, int N, int MODE>
__global__ void __launch_bounds__(Cfg<D, N>::TPB)
k_div(const double* __restrict__ u, const double* __restrict__ scale0, const double* __restrict__ scale1,
      const double* __restrict__ scale2, double* __restrict__ qout, const double* __restrict__ RW2,
      const double* __restrict__ pdir, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter,
      double* __restrict__ red_out, int finalize, long long n, long long n2, double sign) {
  using C = Cfg<D, N>;
  using S1 = typename C::S1;
  using S2 = typename C::S2;
  constexpr int N2 = C::N2, TPB = C::TPB, NK1 = C::NK1, NK2 = C::NK2;
  constexpr int PPT2 = (C::NP2 + TPB - 1) / TPB;
  using SA = Shp<NK2, N, N>;      // after k stage (3D) ; in 2D this is the input itself
  using SB = Shp<NK2, N2, N>;     // after j stage
  __shared__ double su[S1::size];
  __shared__ double sa[2][SA::size];
  __shared__ double sb[3][SB::size];
  __shared__ double st[D][S2::size];
  __shared__ double sred[32];
  const int tid = threadIdx.x;
  if (MODE == 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * C::NP2;
  const long long e1 = (long long)blockIdx.x * C::NP1;
  double acc[PPT2];
#pragma unroll
  for (int m = 0; m < PPT2; ++m) acc[m] = 0.0;
  for (int c = 0; c < D; ++c) {
    const double* sc = (c == 0) ? scale0 : (c == 1 ? scale1 : scale2);
    for (int q = tid; q < C::NP1; q += TPB) {
      double v = u[(long long)c * n + e1 + q];
      if (sc) v *= sc[e1 + q];
      su[S1::lin(q)] = v;
    }
    __syncthreads();
    if constexpr (D == 3) {
      contract<2, N2, N, NK1, N, N, false>(su, sa[0], cm.J12, tid, TPB);
      contract<2, N2, N, NK1, N, N, false>(su, sa[1], cm.D12, tid, TPB);
      __syncthreads();
      contract<1, N2, N, NK2, N, N, false>(sa[0], sb[0], cm.J12, tid, TPB);   // JJ
      contract<1, N2, N, NK2, N, N, false>(sa[0], sb[1], cm.D12, tid, TPB);   // D_s J_t
      contract<1, N2, N, NK2, N, N, false>(sa[1], sb[2], cm.J12, tid, TPB);   // J_s D_t
      __syncthreads();
      contract<0, N2, N, NK2, N2, N, false>(sb[0], st[0], cm.D12, tid, TPB);  // d/dr
      contract<0, N2, N, NK2, N2, N, false>(sb[1], st[1], cm.J12, tid, TPB);  // d/ds
      contract<0, N2, N, NK2, N2, N, false>(sb[2], st[2], cm.J12, tid, TPB);  // d/dt
    } else {
      contract<1, N2, N, NK2, N, N, false>(su, sb[0], cm.J12, tid, TPB);
      contract<1, N2, N, NK2, N, N, false>(su, sb[1], cm.D12, tid, TPB);
      __syncthreads();
      contract<0, N2, N, NK2, N2, N, false>(sb[0], st[0], cm.D12, tid, TPB);
      contract<0, N2, N, NK2, N2, N, false>(sb[1], st[1], cm.J12, tid, TPB);
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < PPT2; ++m) {
      int q = tid + m * TPB;
      if (q < C::NP2) {
        int o = S2::lin(q);
#pragma unroll
        for (int i = 0; i < D; ++i) acc[m] = fma(RW2[(long long)(i * D + c) * n2 + e2 + q], st[i][o], acc[m]);
      }
    }
    __syncthreads();
  }
  double rho[1] = {0.0};
#pragma unroll
  for (int m = 0; m < PPT2; ++m) {
    int q = tid + m * TPB;
    if (q < C::NP2) {
      double v = sign * acc[m];
      qout[e2 + q] = v;
      if (MODE == 1) rho[0] += pdir[e2 + q] * v;
    }
  }
  if (MODE == 1) {
    if (grid_sum_finish<1>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
      cgs->rho = red_out[0];
      cgs->alpha = cgs->rtz1 / red_out[0];
    }
  }
}

// --------------------------------------------------------------------------------------------- advection
// f_k = -( B[(u'.grad)U_k + (U.grad)u'_k] )  - bm1*spng*u'_k           (ADJ = 0, advabp)
// f_i = -( B[sum_j u'_j dU_j/dx_i - (U.grad)u'_i] ) - bm1*spng*u'_i    (ADJ = 1, advabp_adjoint)
// evaluated on the GL(lxd) dealiasing mesh and projected back (convect_new).  Rd = wd * J dr_k/dx_c on the fine mesh.
template <int D, int N>
struct AdvSmem {
  using C = Cfg<D, N>;
  static constexpr int ND = C::ND;
  using S1 = typename C::S1;
  using SD = typename C::SD;
  using SA = Shp<C::NK1, N, ND>;      // after i stage: (NK1, N, ND)
  using SB = Shp<C::NK1, ND, ND>;     // after j stage
  // doubles: su | a[2] | b[3] | F | persistent fine arrays (9 for ADJ, 6 for direct) un-pitched own-point arrays
  static constexpr int work = S1::size + 2 * SA::size + 3 * SB::size + SD::size;
  static constexpr int fine_direct = 2 * D * C::NPD;
  static constexpr int fine_adj = 3 * D * C::NPD;
};

// Computes, for the coarse field in `su`, the fine-mesh value (WANT_F) and/or the D fine-mesh r/s/t
// derivatives one at a time, handing each finished fine array (in sF) to `sink(which)`; which = -1 for the
// value, 0..D-1 for d/dr_which.  Barriers inside; all threads must call.
template <int D, int N, bool WANT_F, bool WANT_G, class Sink>
__device__ __forceinline__ void fine_expand(const double* su, double* sa, double* sb, double* sF, int tid, int nthr,
                                            Sink sink) {
  using C = Cfg<D, N>;
  using A = AdvSmem<D, N>;
  constexpr int ND = C::ND, NK1 = C::NK1;
  using SA = typename A::SA;
  using SB = typename A::SB;
  double* aJ = sa;
  double* aD = sa + SA::size;
  double* bJJ = sb;
  double* bDJ = sb + SB::size;      // D along j
  double* bJD = sb + 2 * SB::size;  // D along i
  contract<0, ND, N, NK1, N, N, false>(su, aJ, cm.Jd, tid, nthr);
  if (WANT_G) contract<0, ND, N, NK1, N, N, false>(su, aD, cm.Dd, tid, nthr);
  __syncthreads();
  contract<1, ND, N, NK1, N, ND, false>(aJ, bJJ, cm.Jd, tid, nthr);
  if (WANT_G) {
    contract<1, ND, N, NK1, N, ND, false>(aJ, bDJ, cm.Dd, tid, nthr);
    contract<1, ND, N, NK1, N, ND, false>(aD, bJD, cm.Jd, tid, nthr);
  }
  __syncthreads();
  if constexpr (D == 3) {
    if (WANT_F) {
      contract<2, ND, N, NK1, ND, ND, false>(bJJ, sF, cm.Jd, tid, nthr);
      __syncthreads();
      sink(-1);
      __syncthreads();
    }
    if (WANT_G) {
      contract<2, ND, N, NK1, ND, ND, false>(bJD, sF, cm.Jd, tid, nthr);
      __syncthreads();
      sink(0);
      __syncthreads();
      contract<2, ND, N, NK1, ND, ND, false>(bDJ, sF, cm.Jd, tid, nthr);
      __syncthreads();
      sink(1);
      __syncthreads();
      contract<2, ND, N, NK1, ND, ND, false>(bJJ, sF, cm.Dd, tid, nthr);
      __syncthreads();
      sink(2);
      __syncthreads();
    }
  } else {
    // 2-D: the j stage already produced the fine arrays (SB has the SD shape)
    if (WANT_F) {
      for (int q = tid; q < SB::size; q += nthr) sF[q] = bJJ[q];
      __syncthreads();
      sink(-1);
      __syncthreads();
    }
    if (WANT_G) {
      for (int q = tid; q < SB::size; q += nthr) sF[q] = bJD[q];
      __syncthreads();
      sink(0);
      __syncthreads();
      for (int q = tid; q < SB::size; q += nthr) sF[q] = bDJ[q];
      __syncthreads();
      sink(1);
      __syncthreads();
    }
  }
}

// SCR = true (3-D default): the persistent fine-mesh arrays (6-9 x lxd^3 doubles per element, 83-124 KB) live in a per-CTA
// global scratch slot instead of shared memory and the CTAs are persistent (grid = resident CTAs), so the slots stay in
// L2; with 256 threads this lets 2 CTAs share an SM (first generation: 512 threads x 128 registers + 156 KB = 1 CTA/SM,
// 6 % of the FP64 peak, most threads idle in the 64-column first tensor stage).
// one element; `fine` = the persistent fine-mesh arrays (shared memory or the CTA's global scratch slot).  Kept out of line:
// inlined into the persistent element loop nvcc hoists the constant-bank matrices into registers and spills (see div3q_element in pcg_kernels.cu).
template <int D, int N, int ADJ>
__device__ __noinline__ void advab_element(const double* __restrict__ up, const double* __restrict__ ub,
                                           const double* __restrict__ Rd, const double* __restrict__ bm1,
                                           const double* __restrict__ spng, double* __restrict__ fout, long long n, long long nd,
                                           double* __restrict__ fine, int e);

template <int D, int N, int ADJ, bool SCR>
__global__ void __launch_bounds__(SCR ? 256 : Cfg<D, N>::TPB_ADV, SCR ? 2 : 1)
k_advab(const double* __restrict__ up, const double* __restrict__ ub, const double* __restrict__ Rd,
        const double* __restrict__ bm1, const double* __restrict__ spng, double* __restrict__ fout, long long n,
        long long nd, double* __restrict__ scratch, int nel) {
  using C = Cfg<D, N>;
  using A = AdvSmem<D, N>;
  extern __shared__ double smem[];
  double* fine = SCR ? scratch + (size_t)blockIdx.x * (3 * D * C::NPD) : smem + A::work;
  for (int e = blockIdx.x; e < nel; e += gridDim.x) advab_element<D, N, ADJ>(up, ub, Rd, bm1, spng, fout, n, nd, fine, e);
}

template <int D, int N, int ADJ>
__device__ __noinline__ void advab_element(const double* __restrict__ up, const double* __restrict__ ub,
                                           const double* __restrict__ Rd, const double* __restrict__ bm1,
                                           const double* __restrict__ spng, double* __restrict__ fout, long long n, long long nd,
                                           double* __restrict__ fine, int e) {
  using C = Cfg<D, N>;
  using A = AdvSmem<D, N>;
  using S1 = typename C::S1;
  using SD = typename C::SD;
  using SA = typename A::SA;
  using SB = typename A::SB;
  constexpr int ND = C::ND, NPD = C::NPD, NP1 = C::NP1, NK1 = C::NK1, NKD = C::NKD;
  extern __shared__ double smem[];
  double* su = smem;
  double* sa = su + S1::size;
  double* sb = sa + 2 * SA::size;
  double* sF = sb + 3 * SB::size;
  double* crb = fine;                    // [D][NPD]   contravariant base flow  (Rd . U)
  double* crp = fine + D * NPD;          // [D][NPD]   direct: contravariant perturbation ; adjoint: u'_j on the fine mesh
  double* accf = fine + 2 * D * NPD;     // adjoint only: [D][NPD] accumulators
  const int tid = threadIdx.x, nthr = blockDim.x;
  {
  const long long e1 = (long long)e * NP1;
  const long long ed = (long long)e * NPD;

  for (int q = tid; q < 2 * D * NPD; q += nthr) fine[q] = 0.0;
  if (ADJ) for (int q = tid; q < D * NPD; q += nthr) accf[q] = 0.0;
  __syncthreads();

  // ---- phase 1: interpolate U_c and u'_c to the fine mesh; build contravariant fields
  for (int c = 0; c < D; ++c) {
    for (int q = tid; q < NP1; q += nthr) su[S1::lin(q)] = ub[(long long)c * n + e1 + q];
    __syncthreads();
    fine_expand<D, N, true, false>(su, sa, sb, sF, tid, nthr, [&](int) {
      for (int q = tid; q < NPD; q += nthr) {
        double v = sF[SD::lin(q)];
#pragma unroll
        for (int i = 0; i < D; ++i) crb[i * NPD + q] = fma(Rd[(long long)(i * D + c) * nd + ed + q], v, crb[i * NPD + q]);
      }
    });
    for (int q = tid; q < NP1; q += nthr) su[S1::lin(q)] = up[(long long)c * n + e1 + q];
    __syncthreads();
    fine_expand<D, N, true, false>(su, sa, sb, sF, tid, nthr, [&](int) {
      for (int q = tid; q < NPD; q += nthr) {
        double v = sF[SD::lin(q)];
        if (ADJ) {
          crp[c * NPD + q] = v;
        } else {
#pragma unroll
          for (int i = 0; i < D; ++i) crp[i * NPD + q] = fma(Rd[(long long)(i * D + c) * nd + ed + q], v, crp[i * NPD + q]);
        }
      }
    });
  }

  if (ADJ) {
    // ---- adjoint phase 2a: acc_i += u'_j * sum_k Rd[k][i] dU_j/dr_k   for all i, looping over j
    for (int j = 0; j < D; ++j) {
      for (int q = tid; q < NP1; q += nthr) su[S1::lin(q)] = ub[(long long)j * n + e1 + q];
      __syncthreads();
      fine_expand<D, N, false, true>(su, sa, sb, sF, tid, nthr, [&](int k) {
        for (int q = tid; q < NPD; q += nthr) {
          double v = sF[SD::lin(q)] * crp[j * NPD + q];
#pragma unroll
          for (int i = 0; i < D; ++i) accf[i * NPD + q] = fma(Rd[(long long)(k * D + i) * nd + ed + q], v, accf[i * NPD + q]);
        }
      });
    }
  }

  // ---- phase 2: per output component
  for (int k = 0; k < D; ++k) {
    // own-point accumulator lives in accf[k] (adjoint) or in a scratch slice (direct: reuse accf region = none) ->
    // for the direct form we accumulate in registers-free fashion inside sF after the loop, so keep a small array:
    // we re-use `crp`/`crb` read-only and accumulate into sa-independent buffer `accd` placed in sb[2] tail? No: use
    // a dedicated unpitched buffer carved from sF is impossible (sF is the landing buffer).  Use accf for both.
    double* acc = ADJ ? (accf + k * NPD) : (fine + 2 * D * NPD);
    if (!ADJ) {
      for (int q = tid; q < NPD; q += nthr) acc[q] = 0.0;
      __syncthreads();
      // (u'.grad) U_k
      for (int q = tid; q < NP1; q += nthr) su[S1::lin(q)] = ub[(long long)k * n + e1 + q];
      __syncthreads();
      fine_expand<D, N, false, true>(su, sa, sb, sF, tid, nthr, [&](int i) {
        for (int q = tid; q < NPD; q += nthr) acc[q] = fma(crp[i * NPD + q], sF[SD::lin(q)], acc[q]);
      });
    }
    // (U.grad) u'_k   (sign + for direct, - for adjoint)
    for (int q = tid; q < NP1; q += nthr) su[S1::lin(q)] = up[(long long)k * n + e1 + q];
    __syncthreads();
    fine_expand<D, N, false, true>(su, sa, sb, sF, tid, nthr, [&](int i) {
      for (int q = tid; q < NPD; q += nthr) {
        double t = crb[i * NPD + q] * sF[SD::lin(q)];
        acc[q] = ADJ ? (acc[q] - t) : (acc[q] + t);
      }
    });
    // project back: sF <- acc (pitched), then Jd^T along k, j, i
    for (int q = tid; q < NPD; q += nthr) sF[SD::lin(q)] = acc[q];
    __syncthreads();
    if constexpr (D == 3) {
      contract<2, N, ND, NKD, ND, ND, false>(sF, sb, cm.Jdt, tid, nthr);       // -> (N, ND, ND) uses SB shape
      __syncthreads();
      contract<1, N, ND, NK1, ND, ND, false>(sb, sa, cm.Jdt, tid, nthr);       // -> (N, N, ND) SA shape
      __syncthreads();
      contract<0, N, ND, NK1, N, ND, false>(sa, su, cm.Jdt, tid, nthr);        // -> (N, N, N)
    } else {
      contract<1, N, ND, 1, ND, ND, false>(sF, sa, cm.Jdt, tid, nthr);         // -> (1, N, ND)
      __syncthreads();
      contract<0, N, ND, 1, N, ND, false>(sa, su, cm.Jdt, tid, nthr);          // -> (1, N, N)
    }
    __syncthreads();
    for (int q = tid; q < NP1; q += nthr) {
      double v = -su[S1::lin(q)];
      if (spng) v -= bm1[e1 + q] * spng[e1 + q] * up[(long long)k * n + e1 + q];
      fout[(long long)k * n + e1 + q] = v;
    }
    __syncthreads();
  }
  }   // element loop
}

// --------------------------------------------------------------------------------------------- advection, second generation (3-D)
// Plane streaming: the CTA keeps the six coarse fields (u'_c, U_c) as k-columns in REGISTERS (thread = field x (i,j)) and
// walks over the lxd fine k-planes.  Per plane: (1) every column thread contracts its k-column with row kf of Jd and Dd
// -> twelve coarse (i,j) slices in shared memory; (2) stage i and (3) stage j expand the slices to the 12x12 plane of the
// value and the three reference derivatives of all six fields -- 144 resp. 288 column tasks, one per thread, the task
// TYPE (which matrix, which source) uniform per warp, matrices from the constant bank; (4) one thread per fine point
// combines the 24 values with the nine fine-mesh metrics Rd of the plane (prefetched one plane ahead with cp.async into
// a double buffer: they are the kernel's HBM traffic) into the three integrands; (5)(6) two transposed stages project
// the plane back to (i,j) and (7) the column threads accumulate it into their k-column of the result with row kf of Jd.
// Shared memory per CTA: 73 KB; nothing but the metrics, the six input fields and the three outputs touches HBM.
// First generation: one field and one derivative at a time through full lxd^3 shared arrays, 64 column tasks in the
// first stage for 512 threads, 6 % of the FP64 peak.
template <int N>
struct Adv2 {
  static constexpr int M = 3 * N / 2;
  static constexpr int NN = N * N, MM = M * M;
  static constexpr int c32(int x) { return (x + 31) / 32 * 32; }
  static constexpr int mx(int a, int b) { return a > b ? a : b; }
  static constexpr int SI_T = c32(6 * N);          // threads reserved per task type in stage i (6 fields x N rows)
  static constexpr int SJ_T = c32(6 * M);          // ... in stage j (6 fields x M columns)
  static constexpr int NT = mx(mx(c32(6 * NN), 3 * SI_T), mx(4 * SJ_T, c32(MM)));
  static constexpr int PN = N + 1, PM = M + 1;     // odd row pitches
  static constexpr int SL = 2 * 6 * N * PN;        // slices  [J|D][field][j][i]
  static constexpr int MID = 3 * 6 * N * PM;       // stage-i outputs [aJ|aD|cJ][field][j][if]; later aliased by P, Q1, Q
  static constexpr int FS = M * PM;                // one fine plane
  static constexpr int FINE = 6 * 4 * FS;          // [field][F|Gs|Gr|Gt][jf][if]
  static constexpr int RD = 2 * 9 * MM;            // double-buffered metric planes
  static constexpr int TOTAL = SL + MID + FINE + RD;
  static_assert(3 * FS + 3 * N * PM + 3 * N * PN <= MID, "P/Q1/Q must fit in the stage-i buffer");
};

// out[m] = sum_l Mat[m*NL + l] in[l*istr], m < NO, stored with stride ostr (Mat: constant bank, compile-time offsets)
template <int NO, int NL>
__device__ __forceinline__ void col_apply(const double* __restrict__ in, int istr, double* __restrict__ out, int ostr,
                                          const double* __restrict__ Mat) {
  double v[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) v[l] = in[l * istr];
#pragma unroll
  for (int m = 0; m < NO; ++m) {
    double sacc = 0.0;
#pragma unroll
    for (int l = 0; l < NL; ++l) sacc = fma(Mat[m * NL + l], v[l], sacc);
    out[m * ostr] = sacc;
  }
}

// stages (2)-(6) of one plane; out of line so that nvcc does not hoist the constant-bank matrices out of the plane loop
template <int N, int ADJ>
__device__ __noinline__ void adv2_plane(double* __restrict__ sm, int tid, int buf, int last) {
  using A = Adv2<N>;
  constexpr int M = A::M, PN = A::PN, PM = A::PM, FS = A::FS, MM = A::MM;
  double* sl = sm;
  double* mid = sm + A::SL;
  double* fine = mid + A::MID;
  const double* rd = fine + A::FINE + buf * 9 * MM;
  double* P = mid;                      // aliases (the stage-i arrays are dead after stage j)
  double* Q1 = mid + 3 * FS;
  double* Q = Q1 + 3 * N * PM;
  // ---- (2) stage i: rows of the slices -> aJ = Jd row, aD = Dd row (from the J-slice), cJ = Jd row (from the D-slice)
  if (tid < 3 * A::SI_T) {
    const int type = tid / A::SI_T, r = tid - type * A::SI_T;
    if (r < 6 * N) {
      const double* src = sl + (type == 2 ? 6 * N * PN : 0) + r * PN;          // r = f*N + j
      double* dst = mid + type * (6 * N * PM) + r * PM;
      if (type == 1) col_apply<M, N>(src, 1, dst, 1, cm.Dd);
      else col_apply<M, N>(src, 1, dst, 1, cm.Jd);
    }
  }
  __syncthreads();
  // ---- (3) stage j: columns -> F = Jd aJ, Gs = Dd aJ, Gr = Jd aD, Gt = Jd cJ
  if (tid < 4 * A::SJ_T) {
    const int type = tid / A::SJ_T, r = tid - type * A::SJ_T;
    if (r < 6 * M) {
      const int f = r / M, i = r - f * M;
      const int srcsel = (type == 2) ? 1 : (type == 3 ? 2 : 0);
      const double* src = mid + srcsel * (6 * N * PM) + f * (N * PM) + i;
      double* dst = fine + (f * 4 + type) * FS + i;
      if (type == 1) col_apply<M, N>(src, PM, dst, PM, cm.Dd);
      else col_apply<M, N>(src, PM, dst, PM, cm.Jd);
    }
  }
  if (last) asm volatile("cp.async.wait_group 0;" ::: "memory");
  else asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncthreads();
  // ---- (4) pointwise: fields 0..2 = u'_c, 3..5 = U_c; derivative order (r, s, t) = types (2, 1, 3)
  if (tid < MM) {
    const int q = tid, o = (q / M) * PM + (q % M);
    double R[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) R[i][c] = rd[(i * 3 + c) * MM + q];
    double val[6], gr[6][3];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      val[f] = fine[(f * 4 + 0) * FS + o];
      gr[f][0] = fine[(f * 4 + 2) * FS + o];
      gr[f][1] = fine[(f * 4 + 1) * FS + o];
      gr[f][2] = fine[(f * 4 + 3) * FS + o];
    }
    double crb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) crb[i] = R[i][0] * val[3] + R[i][1] * val[4] + R[i][2] * val[5];
    double acc[3];
    if (ADJ) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double g = R[0][i] * gr[3 + j][0] + R[1][i] * gr[3 + j][1] + R[2][i] * gr[3 + j][2];   // dU_j/dx_i (weighted)
          a = fma(val[j], g, a);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) a = fma(-crb[k], gr[i][k], a);
        acc[i] = a;
      }
    } else {
      double crp[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) crp[i] = R[i][0] * val[0] + R[i][1] * val[1] + R[i][2] * val[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) a = fma(crp[i], gr[3 + k][i], fma(crb[i], gr[k][i], a));
        acc[k] = a;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) P[c * FS + o] = acc[c];
  }
  __syncthreads();
  // ---- (5) Jd^T along j: P[c][jf][if] -> Q1[c][j][if]
  if (tid < 3 * M) {
    const int c = tid / M, i = tid - c * M;
    col_apply<N, M>(P + c * FS + i, PM, Q1 + c * (N * PM) + i, PM, cm.Jdt);
  }
  __syncthreads();
  // ---- (6) Jd^T along i: Q1[c][j][if] -> Q[c][j][i]
  if (tid < 3 * N) {
    const int c = tid / N, j = tid - c * N;
    col_apply<N, M>(Q1 + c * (N * PM) + j * PM, 1, Q + c * (N * PN) + j * PN, 1, cm.Jdt);
  }
  __syncthreads();
}

template <int N, int ADJ>
__global__ void __launch_bounds__(Adv2<N>::NT, 2)
k_advab2(const double* __restrict__ up, const double* __restrict__ ub, const double* __restrict__ Rd,
         const double* __restrict__ bm1, const double* __restrict__ spng, double* __restrict__ fout, long long n,
         long long nd) {
  using A = Adv2<N>;
  constexpr int M = A::M, NN = A::NN, MM = A::MM, PN = A::PN, PM = A::PM, NT = A::NT;
  extern __shared__ __align__(16) double sm2[];
  double* sl = sm2;
  double* Q = sm2 + A::SL + 3 * A::FS + 3 * N * PM;
  double* rdbuf = sm2 + A::SL + A::MID + A::FINE;
  const int tid = threadIdx.x;
  const long long e1 = (long long)blockIdx.x * (NN * N);
  const long long ed = (long long)blockIdx.x * (MM * M);
  // metric plane kf -> buffer kf&1 with cp.async: 16 bytes per copy when the planes are 16-byte aligned (lxd^2 even), else 8
  auto prefetch = [&](int kf) {
    const unsigned base = (unsigned)__cvta_generic_to_shared(rdbuf + (kf & 1) * 9 * MM);
    if constexpr (MM % 2 == 0) {
      for (int t = tid; t < 9 * MM / 2; t += NT) {
        const int a = t / (MM / 2), w2 = t - a * (MM / 2);
        const double* g = Rd + (long long)a * nd + ed + (long long)kf * MM + 2 * w2;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (unsigned)((a * MM + 2 * w2) * 8)), "l"(g) : "memory");
      }
    } else {
      for (int t = tid; t < 9 * MM; t += NT) {
        const int a = t / MM, w1 = t - a * MM;
        const double* g = Rd + (long long)a * nd + ed + (long long)kf * MM + w1;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(base + (unsigned)((a * MM + w1) * 8)), "l"(g) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0);
  const bool colthread = tid < 6 * NN;
  const int f = tid / NN, ij = tid - f * NN;
  const int slo = (ij / N) * PN + (ij % N);
  double col[N], res[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { col[k] = 0.0; res[k] = 0.0; }
  if (colthread) {
    const double* src = (f < 3) ? up + (long long)f * n : ub + (long long)(f - 3) * n;
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = src[e1 + k * NN + ij];
  }
  for (int kf = 0; kf < M; ++kf) {
    if (kf + 1 < M) prefetch(kf + 1);
    // ---- (1) k-contraction of the register columns with row kf of Jd / Dd
    if (colthread) {
      double wj = 0.0, wd = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        wj = fma(cm.Jd[kf * N + k], col[k], wj);
        wd = fma(cm.Dd[kf * N + k], col[k], wd);
      }
      sl[f * (N * PN) + slo] = wj;
      sl[6 * N * PN + f * (N * PN) + slo] = wd;
    }
    __syncthreads();
    adv2_plane<N, ADJ>(sm2, tid, kf & 1, kf + 1 == M);
    // ---- (7) accumulate the projected plane into the k-columns of the three outputs
    if (tid < 3 * NN) {
      const double qv = Q[f * (N * PN) + slo];
#pragma unroll
      for (int k = 0; k < N; ++k) res[k] = fma(cm.Jd[kf * N + k], qv, res[k]);
    }
  }
  if (tid < 3 * NN) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const long long gi = e1 + k * NN + ij;
      double v = -res[k];
      if (spng) v -= bm1[gi] * spng[gi] * col[k];
      fout[(long long)f * n + gi] = v;
    }
  }
}

// --------------------------------------------------------------------------------------------- geometry (setup; generic, slow, run once)
__global__ void k_metrics(int D, int N, long long n, const double* __restrict__ x, const double* __restrict__ y,
                          const double* __restrict__ z, const double* __restrict__ Dm /*N*N global*/,
                          const double* __restrict__ w1, double* __restrict__ R, double* __restrict__ jac,
                          double* __restrict__ bm1, double* __restrict__ G) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  int np = (D == 3) ? N * N * N : N * N;
  long long e0 = (gid / np) * np;
  int p = (int)(gid - e0);
  int idx[3] = {p % N, (p / N) % N, (D == 3) ? p / (N * N) : 0};
  int str[3] = {1, N, N * N};
  const double* X[3] = {x, y, z};
  double Jm[3][3];
  for (int c = 0; c < D; ++c)
    for (int i = 0; i < D; ++i) {
      double s = 0.0;
      long long base = e0 + p - (long long)idx[i] * str[i];
      for (int l = 0; l < N; ++l) s = fma(Dm[idx[i] * N + l], X[c][base + (long long)l * str[i]], s);
      Jm[c][i] = s;
    }
  double Rm[3][3], J;
  if (D == 2) {
    J = Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0];
    Rm[0][0] = Jm[1][1];  Rm[0][1] = -Jm[0][1];
    Rm[1][0] = -Jm[1][0]; Rm[1][1] = Jm[0][0];
  } else {
    // R[i][c] = cofactor: J * d r_i / d x_c
    Rm[0][0] = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1];
    Rm[0][1] = Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2];
    Rm[0][2] = Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1];
    Rm[1][0] = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2];
    Rm[1][1] = Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0];
    Rm[1][2] = Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2];
    Rm[2][0] = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
    Rm[2][1] = Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1];
    Rm[2][2] = Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0];
    J = Jm[0][0] * Rm[0][0] + Jm[1][0] * Rm[0][1] + Jm[2][0] * Rm[0][2];
  }
  double w = w1[idx[0]] * w1[idx[1]] * ((D == 3) ? w1[idx[2]] : 1.0);
  jac[gid] = J;
  bm1[gid] = J * w;
  for (int i = 0; i < D; ++i)
    for (int c = 0; c < D; ++c) R[(long long)(i * D + c) * n + gid] = Rm[i][c];
  double sc = w / J;
  for (int i = 0; i < D; ++i)
    for (int j = i; j < D; ++j) {
      double s = 0.0;
      for (int c = 0; c < D; ++c) s += Rm[i][c] * Rm[j][c];
      G[(long long)gidx(D, i, j) * n + gid] = s * sc;
    }
}

// out[e, q] = wt(q) * sum_p (M (x) M (x) M)[q,p] in[e,p]; M is (No x Ni) row-major in global memory; wt may be null
__global__ void k_interp_generic(int D, int Ni, int No, long long nout, const double* __restrict__ in,
                                 const double* __restrict__ M, const double* __restrict__ wt1d,
                                 double* __restrict__ out) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nout) return;
  int npo = (D == 3) ? No * No * No : No * No;
  int npi = (D == 3) ? Ni * Ni * Ni : Ni * Ni;
  long long e = gid / npo;
  int q = (int)(gid - e * npo);
  int qi = q % No, qj = (q / No) % No, qk = (D == 3) ? q / (No * No) : 0;
  const double* a = in + e * npi;
  double s = 0.0;
  int nk = (D == 3) ? Ni : 1;
  for (int k = 0; k < nk; ++k) {
    double mk = (D == 3) ? M[qk * Ni + k] : 1.0;
    for (int j = 0; j < Ni; ++j) {
      double mj = M[qj * Ni + j] * mk;
      double t = 0.0;
      for (int i = 0; i < Ni; ++i) t = fma(M[qi * Ni + i], a[(k * Ni + j) * Ni + i], t);
      s = fma(mj, t, s);
    }
  }
  if (wt1d) s *= wt1d[qi] * wt1d[qj] * ((D == 3) ? wt1d[qk] : 1.0);
  out[gid] = s;
}

// diag of the un-assembled stiffness matrix A (exact, including mixed terms)
__global__ void k_hdiagA(int D, int N, long long n, const double* __restrict__ G, const double* __restrict__ Dm,
                         double* __restrict__ out) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  int np = (D == 3) ? N * N * N : N * N;
  long long e0 = (gid / np) * np;
  int p = (int)(gid - e0);
  int idx[3] = {p % N, (p / N) % N, (D == 3) ? p / (N * N) : 0};
  int str[3] = {1, N, N * N};
  double s = 0.0;
  for (int i = 0; i < D; ++i) {
    long long base = e0 + p - (long long)idx[i] * str[i];
    const double* Gii = G + (long long)gidx(D, i, i) * n;
    for (int l = 0; l < N; ++l) {
      double d = Dm[l * N + idx[i]];
      s = fma(d * d, Gii[base + (long long)l * str[i]], s);
    }
  }
  for (int i = 0; i < D; ++i)
    for (int j = i + 1; j < D; ++j)
      s += 2.0 * G[(long long)gidx(D, i, j) * n + gid] * Dm[idx[i] * N + idx[i]] * Dm[idx[j] * N + idx[j]];
  out[gid] = s;
}

// exact diagonal of E = sum_c D_c diag(mbinv_c) D_c^T (no cross-element coupling on the diagonal)
__global__ void k_ediag(int D, int N, int N2, long long n, long long n2, const double* __restrict__ RW2,
                        const double* __restrict__ mb0, const double* __restrict__ mb1, const double* __restrict__ mb2,
                        const double* __restrict__ J12, const double* __restrict__ D12, double* __restrict__ out) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n2) return;
  int np2 = (D == 3) ? N2 * N2 * N2 : N2 * N2;
  int np1 = (D == 3) ? N * N * N : N * N;
  long long e = gid / np2;
  int q = (int)(gid - e * np2);
  int qi = q % N2, qj = (q / N2) % N2, qk = (D == 3) ? q / (N2 * N2) : 0;
  const double* mb[3] = {mb0, mb1, mb2};
  double rw[3][3];
  for (int i = 0; i < D; ++i)
    for (int c = 0; c < D; ++c) rw[i][c] = RW2[(long long)(i * D + c) * n2 + gid];
  double s = 0.0;
  int nk = (D == 3) ? N : 1;
  for (int k = 0; k < nk; ++k)
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) {
        double ji = J12[qi * N + i], di = D12[qi * N + i];
        double jj = J12[qj * N + j], dj = D12[qj * N + j];
        double jk = (D == 3) ? J12[qk * N + k] : 1.0, dk = (D == 3) ? D12[qk * N + k] : 0.0;
        double br = di * jj * jk, bs = ji * dj * jk, bt = ji * jj * dk;
        long long pidx = e * np1 + (k * N + j) * N + i;
        for (int c = 0; c < D; ++c) {
          double v = rw[0][c] * br + rw[1][c] * bs + ((D == 3) ? rw[2][c] * bt : 0.0);
          s = fma(v * v, mb[c][pidx], s);
        }
      }
  out[gid] = s;
}

// max over points of sum_i |u . grad r_i| / dr_i  (compute_cfl with dt = 1); one partial max per block
__global__ void k_cfl(int D, int N, long long n, const double* __restrict__ u, const double* __restrict__ R,
                      const double* __restrict__ jac, const double* __restrict__ z1, double* __restrict__ part) {
  __shared__ double smax[32];
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (gid < n) {
    int np = (D == 3) ? N * N * N : N * N;
    int p = (int)(gid % np);
    int idx[3] = {p % N, (p / N) % N, (D == 3) ? p / (N * N) : 0};
    for (int i = 0; i < D; ++i) {
      double ur = 0.0;
      for (int c = 0; c < D; ++c) ur = fma(u[(long long)c * n + gid], R[(long long)(i * D + c) * n + gid], ur);
      ur /= jac[gid];
      int a = idx[i];
      double dr = (a == 0) ? (z1[1] - z1[0]) : (a == N - 1 ? (z1[N - 1] - z1[N - 2]) : 0.5 * (z1[a + 1] - z1[a - 1]));
      v += fabs(ur / dr);
    }
  }
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = (threadIdx.x < (blockDim.x >> 5)) ? smax[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}

__global__ void k_max_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
  __shared__ double smax[32];
  double v = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) v = fmax(v, part[b]);
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = (threadIdx.x < (blockDim.x >> 5)) ? smax[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) out[0] = v;
  }
}

// ============================================================================================= launchers
int ek_upload_constants(const ConstMats& h) {
  NSB_CUDA(cudaMemcpyToSymbol(cm, &h, sizeof(ConstMats)));
  return 0;
}

#define DISPATCH_DN(c, ...)                                                       \
  do {                                                                              \
    const int key_ = (c)->ldim * 100 + (c)->lx1;                                    \
    switch (key_) {                                                                 \
      case 204: { constexpr int D = 2, N = 4; __VA_ARGS__; } break;                        \
      case 206: { constexpr int D = 2, N = 6; __VA_ARGS__; } break;                        \
      case 208: { constexpr int D = 2, N = 8; __VA_ARGS__; } break;                        \
      case 304: { constexpr int D = 3, N = 4; __VA_ARGS__; } break;                        \
      case 306: { constexpr int D = 3, N = 6; __VA_ARGS__; } break;                        \
      case 308: { constexpr int D = 3, N = 8; __VA_ARGS__; } break;                        \
      default:                                                                      \
        nsb_set_error("unsupported (ldim,lx1)=(%d,%d): built for ldim 2/3, lx1 4/6/8", (c)->ldim, (c)->lx1); \
        return 1;                                                                   \
    }                                                                               \
  } while (0)

static double* g_mats_dev = nullptr;   // D | w1 | J12 | D12 | Jd | wd | z1 | w2 in global memory for the generic setup kernels
struct MatOff { int D, w1, J12, D12, Jd, wd, z1, w2, total; };
static MatOff g_mo;

static int upload_setup_mats(Ctx* c) {
  const ConstMats& h = c->cm;
  std::vector<double> buf;
  auto push = [&](const double* p, int cnt) { int o = (int)buf.size(); buf.insert(buf.end(), p, p + cnt); return o; };
  g_mo.D = push(h.D, c->lx1 * c->lx1);
  g_mo.w1 = push(h.w1, c->lx1);
  g_mo.J12 = push(h.J12, c->lx2 * c->lx1);
  g_mo.D12 = push(h.D12, c->lx2 * c->lx1);
  g_mo.Jd = push(h.Jd, c->lxd * c->lx1);
  g_mo.wd = push(h.wd, c->lxd);
  g_mo.z1 = push(h.z1, c->lx1);
  g_mo.w2 = push(h.w2, c->lx2);
  g_mo.total = (int)buf.size();
  if (g_mats_dev) cudaFree(g_mats_dev);
  NSB_CUDA(cudaMalloc(&g_mats_dev, buf.size() * sizeof(double)));
  NSB_CUDA(cudaMemcpy(g_mats_dev, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

int ek_geometry(Ctx* c) {
  NSB_TRY(upload_setup_mats(c));
  const int D = c->ldim;
  k_metrics<<<nblk(c->n, 128), 128, 0, c->stream>>>(D, c->lx1, c->n, c->xyz[0], c->xyz[1], c->xyz[2],
                                                   g_mats_dev + g_mo.D, g_mats_dev + g_mo.w1, c->R, c->jac, c->bm1, c->G);
  nsb_count_launch();
  for (int q = 0; q < D * D; ++q) {
    k_interp_generic<<<nblk(c->n2, 128), 128, 0, c->stream>>>(D, c->lx1, c->lx2, c->n2, c->R + (long long)q * c->n,
                                                             g_mats_dev + g_mo.J12, g_mats_dev + g_mo.w2,
                                                             c->RW2 + (long long)q * c->n2);
    k_interp_generic<<<nblk(c->nd, 128), 128, 0, c->stream>>>(D, c->lx1, c->lxd, c->nd, c->R + (long long)q * c->n,
                                                             g_mats_dev + g_mo.Jd, g_mats_dev + g_mo.wd,
                                                             c->Rd + (long long)q * c->nd);
    nsb_count_launch(2);
  }
  // bm2inv <- w2 * jac interpolated (inverted by the caller)
  k_interp_generic<<<nblk(c->n2, 128), 128, 0, c->stream>>>(D, c->lx1, c->lx2, c->n2, c->jac, g_mats_dev + g_mo.J12,
                                                           g_mats_dev + g_mo.w2, c->bm2inv);
  k_hdiagA<<<nblk(c->n, 128), 128, 0, c->stream>>>(D, c->lx1, c->n, c->G, g_mats_dev + g_mo.D, c->hdiagA);
  nsb_count_launch(2);
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_ediag(Ctx* c, int adj) {
  k_ediag<<<nblk(c->n2, 64), 64, 0, c->stream>>>(c->ldim, c->lx1, c->lx2, c->n, c->n2, c->RW2, c->mbinv[adj][0],
                                                c->mbinv[adj][1], c->mbinv[adj][c->ldim == 3 ? 2 : 1],
                                                g_mats_dev + g_mo.J12, g_mats_dev + g_mo.D12, c->dinvE[adj]);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_cfl(Ctx* c, const double* u, double* cfl_dev) {
  int nb = nblk(c->n, 256);
  if (nb > NSB_MAX_BLOCKS * NSB_MAX_RED) { nsb_set_error("cfl: grid too large"); return 1; }
  k_cfl<<<nb, 256, 0, c->stream>>>(c->ldim, c->lx1, c->n, u, c->R, c->jac, g_mats_dev + g_mo.z1, c->red_part);
  k_max_final<<<1, 256, 0, c->stream>>>(c->red_part, nb, cfl_dev);
  nsb_count_launch(2);
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_axhelm(Ctx* c, const double* u, double* w, int nfields, double h1, double h2) {
  if (c->ldim == 3) return pk_axhelm(c, 0, u, w, nullptr, nfields, h1, h2);
  DISPATCH_DN(c, k_axhelm<D, N, 0><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(
                     u, w, nullptr, c->G, c->bm1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nfields, c->n, h1, h2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_axhelm_resid(Ctx* c, const double* u, const double* b, double* r, int nfields, double h1, double h2) {
  if (c->ldim == 3) return pk_axhelm(c, 1, u, r, b, nfields, h1, h2);
  DISPATCH_DN(c, k_axhelm<D, N, 1><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(
                     u, r, b, c->G, c->bm1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nfields, c->n, h1, h2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_hcg_dir_ax(Ctx* c, int ncomp, double h1, double h2) {
  // r = rk, p = wk[1], w = wk[2]
  if (c->ldim == 3) return pk_axhelm(c, 2, nullptr, nullptr, nullptr, ncomp, h1, h2);
  DISPATCH_DN(c, k_axhelm<D, N, 2><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(
                     c->rk, c->wk[2], nullptr, c->G, c->bm1, c->dinvH, c->wk[1], c->cgs, c->red_part, c->red_count,
                     c->red_out, c->nranks == 1, ncomp, c->n, h1, h2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_gradt(Ctx* c, const double* p, double* w) {
  if (c->ldim == 3) return pk_gradt(c, p, w);
  DISPATCH_DN(c, k_gradt<D, N, 0><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(p, w, c->RW2, nullptr, nullptr, nullptr, c->n, c->n2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_pcg_dir_gradt(Ctx* c, int adj) {
  // r = pk[0], pdir = pk[2], w = wk[2]
  if (c->ldim == 3) return pk_pcg_dir_gradt(c, adj);
  const double* zsrc = c->pc_kind ? c->pz : c->pk[0];           // separately applied preconditioner: z = M^-1 r replaces dinvE*r
  const double* zscale = c->pc_kind ? c->ones2 : c->dinvE[adj];
  DISPATCH_DN(c, k_gradt<D, N, 1><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(zsrc, c->wk[2], c->RW2, zscale, c->pk[2],
                                                                         c->cgs + 3, c->n, c->n2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_div(Ctx* c, const double* u, const double* scale, double* q, double sign) {
  const double* s0 = scale;
  const double* s1 = scale ? scale + c->n : nullptr;
  const double* s2 = (scale && c->ldim == 3) ? scale + 2 * c->n : nullptr;
  if (c->ldim == 3) return pk_div(c, u, s0, s1, s2, q, sign);
  DISPATCH_DN(c, k_div<D, N, 0><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(u, s0, s1, s2, q, c->RW2, nullptr, nullptr, nullptr,
                                                                       nullptr, nullptr, 0, c->n, c->n2, sign));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_div_mbinv(Ctx* c, const double* u, int adj, double* q, double sign) {
  const double* s0 = c->mbinv[adj][0];
  const double* s1 = c->mbinv[adj][1];
  const double* s2 = c->mbinv[adj][c->ldim == 3 ? 2 : 1];
  if (c->ldim == 3) return pk_div(c, u, s0, c->mask_same[adj] ? nullptr : s1, c->mask_same[adj] ? nullptr : s2, q, sign);
  DISPATCH_DN(c, k_div<D, N, 0><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(u, s0, s1, s2, q, c->RW2, nullptr, nullptr, nullptr,
                                                                       nullptr, nullptr, 0, c->n, c->n2, sign));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

int ek_pcg_div(Ctx* c, int adj) {
  // w = wk[2] (dssum'd), Ep = pk[3], pdir = pk[2]
  if (c->ldim == 3) return pk_pcg_div(c, adj, 0);
  DISPATCH_DN(c, k_div<D, N, 1><<<c->nel, Cfg<D, N>::TPB, 0, c->stream>>>(
                     c->wk[2], c->mbinv[adj][0], c->mbinv[adj][1], c->mbinv[adj][c->ldim == 3 ? 2 : 1], c->pk[3], c->RW2, c->pk[2],
                     c->cgs + 3, c->red_part, c->red_count, c->red_out, c->nranks == 1, c->n, c->n2, 1.0));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

template <int D, int N, int ADJ>
static int launch_advab(Ctx* c, const double* up, const double* ub, const double* spng, double* f) {
  if constexpr (D == 3) {          // plane-streaming second generation (r1d: 11.4 -> 2.65 ms; the first-generation 3-D path was removed in r2)
    using A2 = Adv2<N>;
    const size_t smem2 = (size_t)A2::TOTAL * sizeof(double);
    static bool attr2 = false;
    if (!attr2) {
      NSB_CUDA(cudaFuncSetAttribute(k_advab2<N, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      attr2 = true;
    }
    k_advab2<N, ADJ><<<c->nel, A2::NT, smem2, c->stream>>>(up, ub, c->Rd, c->bm1, spng, f, c->n, c->nd);
    return 0;
  } else {                         // 2-D: first-generation kernel (correctness path of the shipped 2-D configs)
    using A = AdvSmem<D, N>;
    size_t smem = (size_t)(A::work + (ADJ ? A::fine_adj : (A::fine_direct + Cfg<D, N>::NPD))) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
      NSB_CUDA(cudaFuncSetAttribute(k_advab<D, N, ADJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    k_advab<D, N, ADJ, false><<<c->nel, Cfg<D, N>::TPB_ADV, smem, c->stream>>>(up, ub, c->Rd, c->bm1, spng, f, c->n, c->nd, nullptr,
                                                                              c->nel);
    return 0;
  }
}

int ek_advab(Ctx* c, int adjoint, const double* up, const double* ub, const double* spng, double* f) {
  if (adjoint) {
    DISPATCH_DN(c, NSB_TRY((launch_advab<D, N, 1>(c, up, ub, spng, f))));
  } else {
    DISPATCH_DN(c, NSB_TRY((launch_advab<D, N, 0>(c, up, ub, spng, f))));
  }
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
