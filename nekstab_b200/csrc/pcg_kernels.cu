// pcg_kernels.cu -- second-generation 3-D kernels of the pressure operator E = D (mask B~^-1 QQ^T) D^T, the
// operator applied thousands of times per time step by the Jacobi-PCG pressure solve (>95 % of a step).
//
//   k_gradt3 : w_c = D_c^T p      mesh 2 (GL, lx2^3) -> mesh 1 (GLL, lx1^3)   [UPSTREAM navier1.f opgradt/cdtp]
//   k_div3   : q = sum_c D_c u_c  mesh 1 -> mesh 2                              [UPSTREAM navier1.f opdiv/multd]
// optionally fused with the CG vector work (direction update before gradt; p.Ep partial after div) and with the
// gather-scatter: in FUSED mode k_div3 sums the copies of every element-surface node itself from a per-element
// gather table (all copies, ascending dof order => every copy sees the bit-identical sum), so the separate
// dssum kernel and its write-back disappear from the pressure loop.
//
// r1a ncu (profiles/r1a_*): the first-generation kernels were bound by the shared-memory pipe (LSU wavefronts 89 %,
// barrier stalls), not by DRAM.  Changes here: (1) all three components go through each tensor stage together
// (4 barriers per launch instead of 12); (2) the two matrices that feed one output are applied in one pass with
// register accumulation (no shared-memory read-modify-write); (3) the metric multiply is folded into the first
// (gradt) / last (div) stage with the metrics read straight from global memory in column order; (4) gradt's last
// stage stores to global memory directly.  Shared-memory accesses per element drop by ~half.
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
template <int N, int MODE>
__global__ void __launch_bounds__(PK_TPB, 3)
k_gradt3(const double* __restrict__ p, double* __restrict__ w, const double* __restrict__ RW2,
         const double* __restrict__ dinvE, double* __restrict__ pdir, const CGState* __restrict__ cgs, long long n,
         long long n2) {
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
  if (MODE == 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * NP2;
  const long long e1 = (long long)blockIdx.x * NP1;
  static_assert(NP2 <= TPB, "one mesh-2 point per thread");
  if (tid < NP2) {
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
// FUSED 1: the copies of every surface node are summed here from the gather table (replaces dssum)
template <int N, int MODE, int FUSED>
__global__ void __launch_bounds__(PK_TPB, 3)
k_div3(const double* __restrict__ u, const double* __restrict__ scale0, const double* __restrict__ scale1,
       const double* __restrict__ scale2, double* __restrict__ qout, const double* __restrict__ RW2,
       const double* __restrict__ pdir, CGState* __restrict__ cgs, double* __restrict__ part, unsigned* counter,
       double* __restrict__ red_out, int finalize, long long n, long long n2, double sign,
       const int* __restrict__ surf_pts, int ns, const int* __restrict__ nb_off, const int* __restrict__ nb_idx) {
  constexpr int N2 = N - 2, TPB = PK_TPB;
  constexpr int NP1 = N * N * N, NP2 = N2 * N2 * N2;
  using S1 = Shp<N, N, N>;
  using SA = Shp<N2, N, N>;      // after the t stage
  using SB = Shp<N2, N2, N>;     // after the s stage
  using C2 = ColIn<2, N, N, N>;
  using C1 = ColIn<1, N2, N, N>;
  using C0 = ColIn<0, N2, N2, N>;
  // su (3 x S1) is dead after the t stage and sbuf (9 x SB) is first written in the s stage: they share storage
  constexpr int SU = 3 * S1::size, SBB = 9 * SB::size;
  __shared__ double sus[(SU > SBB) ? SU : SBB];
  __shared__ double sa[6][SA::size];       // [c*2 + {J,D}]; dead after the s stage => reused for the partial sums
  __shared__ double sred[32];
  double (*spart)[NP2] = reinterpret_cast<double (*)[NP2]>(&sa[0][0]);
  static_assert(9 * NP2 <= 6 * SA::size, "spart alias");
  double* su = sus;
  double* sbuf = sus;
  const int tid = threadIdx.x;
  if (MODE == 1 && cgs->done) return;
  const long long e2 = (long long)blockIdx.x * NP2;
  const long long e1 = (long long)blockIdx.x * NP1;
  if (FUSED) {
    // ---- stage the three (scaled) components, then overwrite the surface nodes with the gathered sums
    for (int q = tid; q < NP1; q += TPB) {
      const int o = S1::lin(q);
      const double s0 = scale0 ? scale0[e1 + q] : 1.0;
      const double s1 = scale1 ? scale1[e1 + q] : s0;
      const double s2 = scale2 ? scale2[e1 + q] : s0;
      su[o] = u[e1 + q] * s0;
      su[S1::size + o] = u[n + e1 + q] * s1;
      su[2 * S1::size + o] = u[2 * n + e1 + q] * s2;
    }
    __syncthreads();
    const int* off = nb_off + (long long)blockIdx.x * (ns + 1);
    for (int s = tid; s < ns; s += TPB) {
      const int q = surf_pts[s];
      const int a = off[s], b = off[s + 1];
      double g0 = 0.0, g1 = 0.0, g2 = 0.0;
      for (int j = a; j < b; ++j) {
        const int idx = nb_idx[j];
        g0 += u[idx];
        g1 += u[n + idx];
        g2 += u[2 * n + idx];
      }
      const int o = S1::lin(q);
      const double s0 = scale0 ? scale0[e1 + q] : 1.0;
      const double s1 = scale1 ? scale1[e1 + q] : s0;
      const double s2 = scale2 ? scale2[e1 + q] : s0;
      su[o] = g0 * s0;
      su[S1::size + o] = g1 * s1;
      su[2 * S1::size + o] = g2 * s2;
    }
    __syncthreads();
  }
  // ---- t stage: aJ = J12 u, aD = D12 u along t (same input column, two outputs).  Without the fused gather the
  //      column is read straight from global memory: lanes run over (j,i), so each of the N loads is coalesced.
  static_assert(3 * C2::ncol <= TPB, "one task per thread");
  if (tid < 3 * C2::ncol) {
    const int t = tid;
    const int c = t / C2::ncol, col = t - c * C2::ncol;
    double v[N];
    if (FUSED) {
      const double* pin = su + c * S1::size + C2::base(col);
#pragma unroll
      for (int l = 0; l < N; ++l) v[l] = pin[l * C2::stride];
    } else {
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
  __syncthreads();
  double rho[1] = {0.0};
  for (int q = tid; q < NP2; q += TPB) {
    double acc = 0.0;
#pragma unroll
    for (int g = 0; g < 9; ++g) acc = fma(RW2[(long long)g * n2 + e2 + q], spart[g][q], acc);   // metrics: coalesced
    const double v = sign * acc;
    qout[e2 + q] = v;
    if (MODE == 1) rho[0] += pdir[e2 + q] * v;
  }
  if (MODE == 1) {
    if (grid_sum_finish<1>(rho, part, counter, red_out, sred) && finalize && tid == 0) {
      cgs->rho = red_out[0];
      cgs->alpha = cgs->rtz1 / red_out[0];
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
  DISPATCH_N(c, k_gradt3<N, 0><<<c->nel, PK_TPB, 0, c->stream>>>(p, w, c->RW2, nullptr, nullptr, nullptr, c->n, c->n2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int pk_pcg_dir_gradt(Ctx* c, int adj) {
  DISPATCH_N(c, k_gradt3<N, 1><<<c->nel, PK_TPB, 0, c->stream>>>(c->pk[0], c->wk[2], c->RW2, c->dinvE[adj], c->pk[2], c->cgs + 3,
                                                              c->n, c->n2));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int pk_div(Ctx* c, const double* u, const double* s0, const double* s1, const double* s2, double* q, double sign) {
  DISPATCH_N(c, k_div3<N, 0, 0><<<c->nel, PK_TPB, 0, c->stream>>>(u, s0, s1, s2, q, c->RW2, nullptr, nullptr, nullptr, nullptr,
                                                               nullptr, 0, c->n, c->n2, sign, nullptr, 0, nullptr, nullptr));
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int pk_pcg_div(Ctx* c, int adj, int fused) {
  const double* s0 = c->mbinv[adj][0];
  const double* s1 = c->mask_same[adj] ? nullptr : c->mbinv[adj][1];
  const double* s2 = c->mask_same[adj] ? nullptr : c->mbinv[adj][2];
  const GSMap& m = c->gs;
  if (fused) {
    DISPATCH_N(c, k_div3<N, 1, 1><<<c->nel, PK_TPB, 0, c->stream>>>(c->wk[2], s0, s1, s2, c->pk[3], c->RW2, c->pk[2], c->cgs + 3,
                                                                 c->red_part, c->red_count, c->red_out, c->nranks == 1, c->n,
                                                                 c->n2, 1.0, m.surf_pts, m.ns, m.nb_off, m.nb_idx));
  } else {
    DISPATCH_N(c, k_div3<N, 1, 0><<<c->nel, PK_TPB, 0, c->stream>>>(c->wk[2], s0, s1, s2, c->pk[3], c->RW2, c->pk[2], c->cgs + 3,
                                                                 c->red_part, c->red_count, c->red_out, c->nranks == 1, c->n,
                                                                 c->n2, 1.0, nullptr, 0, nullptr, nullptr));
  }
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
