// vec_kernels.cu -- streaming (HBM-bound) kernels: BLAS-1 on Krylov vectors, mass-weighted inner products,
// the fused vector updates of the two Jacobi-PCG loops, EXT/BDF right-hand-side assembly and the
// tall-skinny GEMV pair of the Gram-Schmidt step.
//
// Reference routines restated: core/krylov_subspace.f:24-258 (krylov_inner_product .. krylov_matmul),
// core/krylov_decomposition.f:116-202 (update_hessenberg_matrix), core/eigensolvers.f:466-474 (basis rotation);
// [UPSTREAM Nek5000] math.f glsc3/add2s2/..., perturb.f makextp/makebdfp/lagfieldp, hmholtz.f cggo vector updates.
// All reductions are two-stage and deterministic (per-block partials combined in block order by the last block).
#include "elem_common.cuh"

static inline int grid_for(long long n, int tpb = 256, int per_thread = 4) {
  long long b = (n + (long long)tpb * per_thread - 1) / ((long long)tpb * per_thread);
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}
#define GSTRIDE(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void k_fill(double* a, double v, long long n) { GSTRIDE(i, n) a[i] = v; }
__global__ void k_scale(double* a, double s, long long n) { GSTRIDE(i, n) a[i] *= s; }
__global__ void k_axpy(double* y, double a, const double* x, long long n) { GSTRIDE(i, n) y[i] = fma(a, x[i], y[i]); }
__global__ void k_mul(double* a, const double* b, long long n) { GSTRIDE(i, n) a[i] *= b[i]; }
__global__ void k_inv(double* a, long long n) { GSTRIDE(i, n) a[i] = 1.0 / a[i]; }
__global__ void k_lin2(double* o, double a, const double* x, double b, const double* y, long long n) {
  GSTRIDE(i, n) o[i] = a * x[i] + b * y[i];
}
__global__ void k_add_scalar(double* a, const double* s, double factor, long long n) {
  const double v = factor * s[0];
  GSTRIDE(i, n) a[i] += v;
}

// sum a*b*w (b, w optional)
__global__ void k_dot3(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ w,
                       long long n, double* part, unsigned* counter, double* out) {
  __shared__ double sred[32];
  double v[1] = {0.0};
  GSTRIDE(i, n) {
    double t = a[i];
    if (b) t *= b[i];
    if (w) t *= w[i];
    v[0] += t;
  }
  grid_sum_finish<1>(v, part, counter, out, sred);
}

#define LAUNCH1(kernel, n, ...)                                              \
  do {                                                                       \
    kernel<<<grid_for(n), 256, 0, c->stream>>>(__VA_ARGS__);                 \
    nsb_count_launch();                                                      \
    NSB_CUDA(cudaGetLastError());                                            \
  } while (0)

int vk_fill(Ctx* c, double* a, double v, long long n) { LAUNCH1(k_fill, n, a, v, n); return 0; }
int vk_copy(Ctx* c, double* dst, const double* src, long long n) {
  NSB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int vk_scale(Ctx* c, double* a, double s, long long n) { LAUNCH1(k_scale, n, a, s, n); return 0; }
int vk_axpy(Ctx* c, double* y, double a, const double* x, long long n) { LAUNCH1(k_axpy, n, y, a, x, n); return 0; }
int vk_mul(Ctx* c, double* a, const double* b, long long n) { LAUNCH1(k_mul, n, a, b, n); return 0; }
int vk_inv(Ctx* c, double* a, long long n) { LAUNCH1(k_inv, n, a, n); return 0; }
int vk_lin2(Ctx* c, double* o, double a, const double* x, double b, const double* y, long long n) {
  LAUNCH1(k_lin2, n, o, a, x, b, y, n);
  return 0;
}
int vk_add_scalar_from_dev(Ctx* c, double* a, const double* s, double f, long long n) {
  LAUNCH1(k_add_scalar, n, a, s, f, n);
  return 0;
}
int vk_dot3(Ctx* c, const double* a, const double* b, const double* w, long long n, double* out_dev) {
  LAUNCH1(k_dot3, n, a, b, w, n, c->red_part, c->red_count, out_dev);
  return 0;
}
int vk_sum(Ctx* c, const double* a, long long n, double* out_dev) { return vk_dot3(c, a, nullptr, nullptr, n, out_dev); }

int vk_allreduce_sum(Ctx* c, double* dev, int count) {
  if (c->p2p.on) return p2p_allreduce(c, dev, count, 0, nullptr, 0, 0);
  if (c->nranks > 1) NSB_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, ncclSum, c->comm, c->stream));
  return 0;
}
int vk_allreduce_max(Ctx* c, double* dev, int count) {
  if (c->p2p.on) return p2p_allreduce(c, dev, count, 1, nullptr, 0, 0);
  if (c->nranks > 1) NSB_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, ncclMax, c->comm, c->stream));
  return 0;
}

// full Navier-Stokes step with the sponge on (jp = 0 branch of nekStab_forcing, core/utils.f:166-171):
// f_c += bm1 * spng_str * spng_fun * (spng_vr_c - u_c)
__global__ void k_sponge_dns(double* __restrict__ f, const double* __restrict__ u, const double* __restrict__ ref,
                             const double* __restrict__ spng, const double* __restrict__ bm1, double str, long long n, int D) {
  const long long tot = n * D;
  GSTRIDE(i, tot) {
    const long long ip = i % n;
    f[i] = fma(bm1[ip] * str * spng[ip], ref[i] - u[i], f[i]);
  }
}
int vk_sponge_dns(Ctx* c, double* f, const double* u) {
  if (c->spng_str_dns == 0.0 || !c->spng || !c->spng_ref) return 0;
  LAUNCH1(k_sponge_dns, c->n * c->ldim, f, u, c->spng_ref, c->spng, c->bm1, c->spng_str_dns, c->n, c->ldim);
  return 0;
}

// ---------------------------------------------------------------------------------------- stepper pointwise
// b_c = sum_j ab_j f_j,c + (bm1/dt) sum_j bd_j u_j,c     (makextp + makebdfp; vtrans = rho folded into bdr)
struct RhsArgs {
  const double* f[3];
  const double* u[3];
  double ab[3];
  double bdr[3];   // rho*bd(j+1)/dt
  int k;
};
__global__ void k_make_rhs(double* __restrict__ b, RhsArgs a, const double* __restrict__ bm1, long long n, int D) {
  const long long tot = n * D;
  GSTRIDE(i, tot) {
    double s = 0.0, h = 0.0;
    for (int j = 0; j < a.k; ++j) {
      s = fma(a.ab[j], a.f[j][i], s);
      h = fma(a.bdr[j], a.u[j][i], h);
    }
    b[i] = fma(bm1[i % n], h, s);
  }
}
int vk_make_rhs(Ctx* c, double* b, int k, const double* ab, const double* bd) {
  RhsArgs a;
  const double* us[3] = {c->u, c->ulag[0], c->ulag[1]};
  for (int j = 0; j < 3; ++j) {
    a.f[j] = c->f[j];
    a.u[j] = us[j];
    a.ab[j] = (j < k) ? ab[j] : 0.0;
    a.bdr[j] = (j < k) ? c->rho * bd[j + 1] / c->dt : 0.0;
  }
  a.k = k;
  LAUNCH1(k_make_rhs, c->n * c->ldim, b, a, c->bm1, c->n, c->ldim);
  return 0;
}

__global__ void k_mask3(double* __restrict__ r, const double* __restrict__ m0, const double* __restrict__ m1,
                        const double* __restrict__ m2, long long n, int D) {
  GSTRIDE(i, n) {
    r[i] *= m0[i];
    r[n + i] *= m1[i];
    if (D == 3) r[2 * n + i] *= m2[i];
  }
}
int vk_mask_fields(Ctx* c, double* r, int adj) {
  LAUNCH1(k_mask3, c->n, r, c->mask[adj][0], c->mask[adj][1], c->mask[adj][c->ldim == 3 ? 2 : 1], c->n, c->ldim);
  return 0;
}

__global__ void k_press_extrap(double* __restrict__ pt, const double* __restrict__ pr, const double* __restrict__ prlag,
                               int k, long long n2) {
  GSTRIDE(i, n2) pt[i] = (k == 3) ? 2.0 * pr[i] - prlag[i] : pr[i];
}
int vk_press_extrap(Ctx* c, int k) {
  LAUNCH1(k_press_extrap, c->n2, c->pt, c->pr, c->prlag, k, c->n2);
  return 0;
}

// u_new_c += mbinv_c * w_c   (u_new already holds u + du);   p_new = pt + h2*phi written over prlag
__global__ void k_final_update(double* __restrict__ un, const double* __restrict__ w, const double* __restrict__ m0,
                               const double* __restrict__ m1, const double* __restrict__ m2, long long n, int D,
                               double* __restrict__ pnew, const double* __restrict__ pt, const double* __restrict__ phi,
                               double h2, long long n2) {
  GSTRIDE(i, n) {
    un[i] = fma(m0[i], w[i], un[i]);
    un[n + i] = fma(m1[i], w[n + i], un[n + i]);
    if (D == 3) un[2 * n + i] = fma(m2[i], w[2 * n + i], un[2 * n + i]);
  }
  GSTRIDE(i, n2) pnew[i] = fma(h2, phi[i], pt[i]);
}
int vk_final_update(Ctx* c, int adj, double h2) {
  // un = ulag[1] (new velocity buffer), w = wk[2], phi = pk[1], pnew -> prlag
  LAUNCH1(k_final_update, c->n, c->ulag[1], c->wk[2], c->mbinv[adj][0], c->mbinv[adj][1],
          c->mbinv[adj][c->ldim == 3 ? 2 : 1], c->n, c->ldim, c->prlag, c->pt, c->pk[1], h2, c->n2);
  return 0;
}

__global__ void k_dinvH(double* __restrict__ o, const double* __restrict__ a, const double* __restrict__ b, double h1,
                        double h2, long long n) {
  GSTRIDE(i, n) o[i] = 1.0 / (h1 * a[i] + h2 * b[i]);
}
int vk_dinvH(Ctx* c, double h1, double h2) {
  if (c->dinvH_h1 == h1 && c->dinvH_h2 == h2) return 0;
  LAUNCH1(k_dinvH, c->n, c->dinvH, c->hdiagA, c->hdiagB, h1, h2, c->n);
  c->dinvH_h1 = h1;
  c->dinvH_h2 = h2;
  return 0;
}

// ---------------------------------------------------------------------------------------- CG scalar logic
__device__ void hcg_finalize_init(CGState* s, const double* sums, int ncomp) {
  for (int f = 0; f < ncomp; ++f) {
    s[f].rtz1 = sums[2 * f];
    s[f].rtz2 = 1.0;
    s[f].beta = 0.0;
    s[f].alpha = 0.0;
    s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
    s[f].iter = 0;
    s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].maxit <= 0);
  }
}
__device__ void hcg_finalize_update(CGState* s, const double* sums, int ncomp) {
  for (int f = 0; f < ncomp; ++f) {
    if (s[f].done) continue;
    s[f].rtz2 = s[f].rtz1;
    s[f].rtz1 = sums[2 * f];
    s[f].beta = s[f].rtz1 / s[f].rtz2;
    s[f].rnorm = sqrt(fmax(sums[2 * f + 1], 0.0) / s[f].vol);
    s[f].iter += 1;
    s[f].done = (s[f].rnorm <= s[f].tol) || (s[f].iter >= s[f].maxit) || !(s[f].rnorm == s[f].rnorm);
  }
}
__device__ void cg_finalize_rho(CGState* s, const double* sums, int ncomp) {
  for (int f = 0; f < ncomp; ++f)
    if (!s[f].done) {
      s[f].rho = sums[f];
      s[f].alpha = s[f].rtz1 / sums[f];
    }
}
// variants for a preconditioner applied by separate kernels (pmg.cu): the update kernels only produce the residual norm,
// z^T r arrives later (kind 5)
__device__ void pcg_finalize_init_norm(CGState* s, const double* sums) {
  s->rtz1 = 1.0; s->rtz2 = 1.0; s->beta = 0.0; s->alpha = 0.0;
  s->rnorm = sqrt(fmax(sums[1], 0.0) / s->vol);
  s->iter = 0;
  s->done = (s->rnorm <= s->tol) || (s->maxit <= 0);
}
__device__ void pcg_finalize_update_norm(CGState* s, const double* sums) {
  if (s->done) return;
  s->rtz2 = s->rtz1;
  s->rnorm = sqrt(fmax(sums[1], 0.0) / s->vol);
  s->iter += 1;
  s->done = (s->rnorm <= s->tol) || (s->iter >= s->maxit) || !(s->rnorm == s->rnorm);
}
__device__ void pcg_finalize_rtz(CGState* s, const double* sums) {
  if (s->done) return;
  s->rtz1 = sums[0];
  s->beta = (s->iter == 0) ? 0.0 : sums[0] / s->rtz2;
}
// kind 0: init (2 sums/comp), 1: update (2 sums/comp), 2: rho (1 sum/comp), 3/4: init/update without z^T r, 5: z^T r
__global__ void k_cg_finalize(CGState* s, const double* sums, int ncomp, int kind) {
  if (threadIdx.x || blockIdx.x) return;
  if (kind == 0) hcg_finalize_init(s, sums, ncomp);
  else if (kind == 1) hcg_finalize_update(s, sums, ncomp);
  else if (kind == 2) cg_finalize_rho(s, sums, ncomp);
  else if (kind == 3) pcg_finalize_init_norm(s, sums);
  else if (kind == 4) pcg_finalize_update_norm(s, sums);
  else pcg_finalize_rtz(s, sums);
}
int vk_cg_finalize_multi(Ctx* c, CGState* s, int ncomp, int kind) {   // multi-rank path: allreduce then finalize
  int cnt = (kind == 2 || kind == 5) ? ncomp : 2 * ncomp;
  if (kind == 2 && s == c->cgs) cnt = 3;
  if (c->p2p.on) return p2p_allreduce(c, c->red_out, cnt, 0, s, ncomp, kind);   // all-reduce + scalar update in one kernel
  NSB_TRY(vk_allreduce_sum(c, c->red_out, cnt));
  k_cg_finalize<<<1, 32, 0, c->stream>>>(s, c->red_out, ncomp, kind);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

// Helmholtz CG (up to 3 components batched; r already assembled and masked)
__global__ void k_hcg_init(const double* __restrict__ r, double* __restrict__ x, double* __restrict__ p,
                           const double* __restrict__ dinv, const double* __restrict__ mult,
                           const double* __restrict__ binv, long long n, int ncomp, CGState* cgs, double* part,
                           unsigned* counter, double* out, int finalize) {
  __shared__ double sred[6 * 32];
  double v[6] = {0, 0, 0, 0, 0, 0};
  GSTRIDE(i, n) {
    const double m = mult[i], di = dinv[i] * m, bi = binv[i] * m;
    for (int f = 0; f < ncomp; ++f) {
      const double rr = r[f * n + i];
      x[f * n + i] = 0.0;
      p[f * n + i] = 0.0;
      v[2 * f] = fma(rr * rr, di, v[2 * f]);
      v[2 * f + 1] = fma(rr * rr, bi, v[2 * f + 1]);
    }
  }
  if (grid_sum_finish<6>(v, part, counter, out, sred) && finalize && threadIdx.x == 0) hcg_finalize_init(cgs, out, ncomp);
}
template <int PN>   // PN = lx1 when w arrives in the surface-first element layout (compile-time: the index map is all shifts and masks), else 0
__global__ void k_hcg_update(double* __restrict__ r, double* __restrict__ x, const double* __restrict__ p,
                             const double* __restrict__ w, const double* __restrict__ m0, const double* __restrict__ m1,
                             const double* __restrict__ m2, const double* __restrict__ dinv,
                             const double* __restrict__ mult, const double* __restrict__ binv, long long n, int ncomp,
                             CGState* cgs, double* part, unsigned* counter, double* out, int finalize) {
  __shared__ double sred[6 * 32];
  double v[6] = {0, 0, 0, 0, 0, 0};
  bool active[3];
  double alpha[3];
  bool any = false;
  for (int f = 0; f < 3; ++f) {
    active[f] = (f < ncomp) && !cgs[f].done;
    alpha[f] = active[f] ? cgs[f].alpha : 0.0;
    any |= active[f];
  }
  if (!any) return;
  constexpr int np = PN * PN * PN;
  GSTRIDE(i, n) {
    const double m = mult[i], di = dinv[i] * m, bi = binv[i] * m;
    // w = H p arrives from k_axhelm3p / dssum in the surface-first element layout (PN = lx1), everything else is natural
    long long iw = i;
    if (PN) {
      const long long e = i / np;
      iw = e * np + SurfFirst<(PN ? PN : 4)>::pos_lin((int)(i - e * np));
    }
    for (int f = 0; f < ncomp; ++f) {
      if (!active[f]) continue;
      const double mk = (f == 0) ? m0[i] : (f == 1 ? m1[i] : m2[i]);
      const long long gi = f * n + i;
      x[gi] = fma(alpha[f], p[gi], x[gi]);
      const double rr = fma(-alpha[f], mk * w[f * n + iw], r[gi]);
      r[gi] = rr;
      v[2 * f] = fma(rr * rr, di, v[2 * f]);
      v[2 * f + 1] = fma(rr * rr, bi, v[2 * f + 1]);
    }
  }
  if (grid_sum_finish<6>(v, part, counter, out, sred) && finalize && threadIdx.x == 0) hcg_finalize_update(cgs, out, ncomp);
}
int vk_hcg_init(Ctx* c, int ncomp) {
  LAUNCH1(k_hcg_init, c->n, c->rk, c->wk[3], c->wk[1], c->dinvH, c->mult, c->binv, c->n, ncomp, c->cgs, c->red_part,
          c->red_count, c->red_out, c->nranks == 1);
  if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, c->cgs, ncomp, 0));
  return 0;
}
int vk_hcg_update(Ctx* c, int ncomp, int adj) {
  if (ncomp == 3 && perm_h_active(c)) {           // lx1 = 8 (perm_h_active)
    LAUNCH1(k_hcg_update<8>, c->n, c->rk, c->wk[3], c->wk[1], c->wk[2], c->mask[adj][0], c->mask[adj][1],
            c->mask[adj][c->ldim == 3 ? 2 : 1], c->dinvH, c->mult, c->binv, c->n, ncomp, c->cgs, c->red_part, c->red_count,
            c->red_out, c->nranks == 1);
  } else {
    LAUNCH1(k_hcg_update<0>, c->n, c->rk, c->wk[3], c->wk[1], c->wk[2], c->mask[adj][0], c->mask[adj][1],
            c->mask[adj][c->ldim == 3 ? 2 : 1], c->dinvH, c->mult, c->binv, c->n, ncomp, c->cgs, c->red_part, c->red_count,
            c->red_out, c->nranks == 1);
  }
  if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, c->cgs, ncomp, 1));
  return 0;
}

// pressure CG
__global__ void k_pcg_init(const double* __restrict__ r, double* __restrict__ x, double* __restrict__ p,
                           const double* __restrict__ dinv, const double* __restrict__ bm2inv, long long n2, CGState* cgs,
                           double* part, unsigned* counter, double* out, int finalize) {
  __shared__ double sred[2 * 32];
  double v[2] = {0, 0};
  GSTRIDE(i, n2) {
    const double rr = r[i];
    x[i] = 0.0;
    p[i] = 0.0;
    v[0] = fma(rr * rr, dinv[i], v[0]);
    v[1] = fma(rr * rr, bm2inv[i], v[1]);
  }
  if (grid_sum_finish<2>(v, part, counter, out, sred) && finalize && threadIdx.x == 0) {
    if (finalize == 1) hcg_finalize_init(cgs, out, 1);
    else pcg_finalize_init_norm(cgs, out);
  }
}
__global__ void k_pcg_update(double* __restrict__ r, double* __restrict__ x, const double* __restrict__ p,
                             const double* __restrict__ ep, const double* __restrict__ dinv,
                             const double* __restrict__ bm2inv, long long n2, CGState* cgs, double* part,
                             unsigned* counter, double* out, int finalize) {
  __shared__ double sred[2 * 32];
  if (cgs->done) return;
  const double alpha = cgs->alpha;
  double v[2] = {0, 0};
  GSTRIDE(i, n2) {
    x[i] = fma(alpha, p[i], x[i]);
    const double rr = fma(-alpha, ep[i], r[i]);
    r[i] = rr;
    v[0] = fma(rr * rr, dinv[i], v[0]);
    v[1] = fma(rr * rr, bm2inv[i], v[1]);
  }
  if (grid_sum_finish<2>(v, part, counter, out, sred) && finalize && threadIdx.x == 0) {
    if (finalize == 1) hcg_finalize_update(cgs, out, 1);
    else pcg_finalize_update_norm(cgs, out);
  }
}
// finalize argument of the two kernels: 0 = multi-rank (scalars updated after the all-reduce), 1 = Jacobi (z^T r fused),
// 2 = separate preconditioner (norm only; pm_apply delivers z^T r)
int vk_pcg_init(Ctx* c, int adj) {
  const int pc = c->pc_kind != 0;
  LAUNCH1(k_pcg_init, c->n2, c->pk[0], c->pk[1], c->pk[2], c->dinvE[adj], c->bm2inv, c->n2, c->cgs + 3, c->red_part,
          c->red_count, c->red_out, c->nranks == 1 ? 1 + pc : 0);
  if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, c->cgs + 3, 1, pc ? 3 : 0));
  return 0;
}
int vk_pcg_update(Ctx* c, int adj) {
  const int pc = c->pc_kind != 0;
  LAUNCH1(k_pcg_update, c->n2, c->pk[0], c->pk[1], c->pk[2], c->pk[3], c->dinvE[adj], c->bm2inv, c->n2, c->cgs + 3,
          c->red_part, c->red_count, c->red_out, c->nranks == 1 ? 1 + pc : 0);
  if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, c->cgs + 3, 1, pc ? 4 : 1));
  return 0;
}

// ---------------------------------------------------------------------------------------- Krylov basis kernels
// h_j = sum_i Q_j[i] * W[i mod n] * f[i], i < nw (= ldim*n: pressure carries no weight, core/krylov_subspace.f:37-45)
// Each block owns a contiguous chunk of rows and streams the k basis vectors over it; f*W stays in registers.
template <int RPT>
__global__ void __launch_bounds__(256)
k_multidot(const double* __restrict__ Q, long long vlen, int k, const double* __restrict__ f,
           const double* __restrict__ W, long long n, long long nw, double* __restrict__ part /*[grid][k]*/) {
  extern __shared__ double sacc[];   // [k][8 warps]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < k * 8; j += blockDim.x) sacc[j] = 0.0;
  __syncthreads();
  const long long chunk = 256LL * RPT;
  for (long long base = (long long)blockIdx.x * chunk; base < nw; base += (long long)gridDim.x * chunk) {
    double wf[RPT];
    long long idx[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      long long i = base + r * 256 + threadIdx.x;
      idx[r] = i;
      wf[r] = (i < nw) ? (W ? f[i] * W[i % n] : f[i]) : 0.0;
    }
    for (int j = 0; j < k; ++j) {
      const double* q = Q + (long long)j * vlen;
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r)
        if (idx[r] < nw) s = fma(q[idx[r]], wf[r], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) sacc[j * 8 + wid] += s;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sacc[j * 8 + w];
    part[(long long)blockIdx.x * k + j] = s;
  }
}
__global__ void k_multidot_final(const double* __restrict__ part, int nb, int k, double* __restrict__ h) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  double s = 0.0;
  for (int b = 0; b < nb; ++b) s += part[(long long)b * k + j];
  h[j] = s;
}
// generic form: h_j = sum_{i<nw} Q[j*vlen+i] * W[i%n] * f[i]  (W may be null), all-reduced over ranks
int vk_multidot_raw(Ctx* c, int k, const double* Q, long long vlen, const double* f, const double* W, long long n, long long nw,
                    double* h_dev) {
  int nb = grid_for(nw, 256, 4);
  if ((long long)nb * k > c->hpart_cap) {
    if (c->hpart) cudaFree(c->hpart);
    c->hpart_cap = (long long)nb * k * 2;
    NSB_CUDA(cudaMalloc(&c->hpart, c->hpart_cap * sizeof(double)));
  }
  size_t smem = (size_t)k * 8 * sizeof(double);
  if (smem > 200 * 1024) { nsb_set_error("multidot: k too large"); return 1; }
  if (smem > 48 * 1024) NSB_CUDA(cudaFuncSetAttribute(k_multidot<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_multidot<4><<<nb, 256, smem, c->stream>>>(Q, vlen, k, f, W, n, nw, c->hpart);
  k_multidot_final<<<(k + 127) / 128, 128, 0, c->stream>>>(c->hpart, nb, k, h_dev);
  nsb_count_launch(2);
  NSB_CUDA(cudaGetLastError());
  NSB_TRY(vk_allreduce_sum(c, h_dev, k));
  return 0;
}
// out[i] = a*f[i] + sign * sum_j h_j Q[j*vlen+i], i < vlen
int vk_multiaxpy_raw(Ctx* c, int k, const double* Q, long long vlen, const double* f, double a, const double* h_dev, double sign,
                     double* out);

int vk_multidot(Ctx* c, int k, int first_slot, int slot_f, double* h_dev) {
  const long long nw = c->n * c->ldim;
  int nb = grid_for(nw, 256, 4);
  if ((long long)nb * k > c->hpart_cap) {
    if (c->hpart) cudaFree(c->hpart);
    c->hpart_cap = (long long)nb * k * 2;
    NSB_CUDA(cudaMalloc(&c->hpart, c->hpart_cap * sizeof(double)));
  }
  size_t smem = (size_t)k * 8 * sizeof(double);
  if (smem > 200 * 1024) { nsb_set_error("multidot: k too large"); return 1; }
  if (smem > 48 * 1024) NSB_CUDA(cudaFuncSetAttribute(k_multidot<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_multidot<4><<<nb, 256, smem, c->stream>>>(slot_ptr(c, first_slot), c->vlen, k, slot_ptr(c, slot_f), c->bm1s, c->n, nw, c->hpart);
  k_multidot_final<<<(k + 127) / 128, 128, 0, c->stream>>>(c->hpart, nb, k, h_dev);
  nsb_count_launch(2);
  NSB_CUDA(cudaGetLastError());
  NSB_TRY(vk_allreduce_sum(c, h_dev, k));
  return 0;
}

// out[i] = a*f[i] + sign * sum_j h_j Q_j[i] over the whole vector (velocity and pressure)
__global__ void __launch_bounds__(256)
k_multiaxpy(const double* __restrict__ Q, long long vlen, int k, const double* f, double a,
            const double* __restrict__ h, double sign, double* out) {      // f and out may be the same vector: no __restrict__
  extern __shared__ double sh[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) sh[j] = sign * h[j];
  __syncthreads();
  for (long long base = (long long)blockIdx.x * 1024; base < vlen; base += (long long)gridDim.x * 1024) {
    double acc[4];
    long long idx[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      idx[r] = base + r * 256 + threadIdx.x;
      acc[r] = (idx[r] < vlen && a != 0.0) ? a * f[idx[r]] : 0.0;
    }
    for (int j = 0; j < k; ++j) {
      const double* q = Q + (long long)j * vlen;
      const double hj = sh[j];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (idx[r] < vlen) acc[r] = fma(hj, q[idx[r]], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (idx[r] < vlen) out[idx[r]] = acc[r];
  }
}
int vk_multiaxpy_raw(Ctx* c, int k, const double* Q, long long vlen, const double* f, double a, const double* h_dev, double sign,
                     double* out) {
  size_t smem = (size_t)k * sizeof(double);
  k_multiaxpy<<<grid_for(vlen, 256, 4), 256, smem, c->stream>>>(Q, vlen, k, f, a, h_dev, sign, out);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int vk_multiaxpy(Ctx* c, int k, int first_slot, int slot_f, const double* h_dev, double sign) {
  size_t smem = (size_t)k * sizeof(double);
  double* f = slot_ptr(c, slot_f);
  k_multiaxpy<<<grid_for(c->vlen, 256, 4), 256, smem, c->stream>>>(slot_ptr(c, first_slot), c->vlen, k, f, 1.0, h_dev, sign, f);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
int vk_gemv_out(Ctx* c, int k, int first_slot, const double* y_dev, int slot_out) {
  size_t smem = (size_t)k * sizeof(double);
  double* o = slot_ptr(c, slot_out);
  k_multiaxpy<<<grid_for(c->vlen, 256, 4), 256, smem, c->stream>>>(slot_ptr(c, first_slot), c->vlen, k, o, 0.0, y_dev, 1.0, o);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

// in-place Q(:,0:k) <- Q(:,0:k) * S  (S column-major, lds); each block stages a [k][32 rows] tile in shared memory
__global__ void __launch_bounds__(256)
k_rotate(double* __restrict__ Q, long long vlen, int k, const double* __restrict__ S, int lds) {
  extern __shared__ double tile[];   // [k][32]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long base = (long long)blockIdx.x * 32; base < vlen; base += (long long)gridDim.x * 32) {
    const long long row = base + lane;
    for (int j = wid; j < k; j += 8) tile[j * 32 + lane] = (row < vlen) ? Q[(long long)j * vlen + row] : 0.0;
    __syncthreads();
    for (int jo = wid; jo < k; jo += 8) {
      const double* sc = S + (long long)jo * lds;
      double s = 0.0;
      for (int j = 0; j < k; ++j) s = fma(tile[j * 32 + lane], __ldg(&sc[j]), s);
      if (row < vlen) Q[(long long)jo * vlen + row] = s;
    }
    __syncthreads();
  }
}
int vk_rotate(Ctx* c, int k, int first_slot, const double* S_dev, int lds) {
  size_t smem = (size_t)k * 32 * sizeof(double);
  if (smem > 200 * 1024) { nsb_set_error("rotate: k too large (max 800)"); return 1; }
  if (smem > 48 * 1024) NSB_CUDA(cudaFuncSetAttribute(k_rotate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long nb = (c->vlen + 31) / 32;
  if (nb > 148 * 8) nb = 148 * 8;
  k_rotate<<<(int)nb, 256, smem, c->stream>>>(slot_ptr(c, first_slot), c->vlen, k, S_dev, lds);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------- direct / adjoint mode post-processing
// biorthogonalize (core/sensitivity.f:486-501): (a_re, a_im) <- ((g a_re - d a_im), (g a_im + d a_re)) / (g^2 + d^2), whole vectors
__global__ void k_rotate_pair(double* __restrict__ are, double* __restrict__ aim, double g, double d, double inv_den, long long n) {
  GSTRIDE(i, n) {
    const double r = are[i], m = aim[i];
    are[i] = (g * r - d * m) * inv_den;
    aim[i] = (g * m + d * r) * inv_den;
  }
}
// wave_maker (core/sensitivity.f:69-71): |u_direct| * |u_adjoint| pointwise over the velocity components
__global__ void k_wavemaker(const double* __restrict__ dre, const double* __restrict__ dim, const double* __restrict__ are,
                            const double* __restrict__ aim, double* __restrict__ wm, long long n, int D) {
  GSTRIDE(i, n) {
    double s1 = 0.0, s2 = 0.0;
    for (int c = 0; c < D; ++c) {
      const long long j = (long long)c * n + i;
      s1 += dre[j] * dre[j] + dim[j] * dim[j];
      s2 += are[j] * are[j] + aim[j] * aim[j];
    }
    wm[i] = sqrt(s1) * sqrt(s2);
  }
}
int vk_rotate_pair(Ctx* c, double* are, double* aim, double g, double d, long long n) {
  LAUNCH1(k_rotate_pair, n, are, aim, g, d, 1.0 / (g * g + d * d), n);
  return 0;
}
int vk_wavemaker(Ctx* c, const double* dre, const double* dim, const double* are, const double* aim, double* wm) {
  LAUNCH1(k_wavemaker, c->n, dre, dim, are, aim, wm, c->n, c->ldim);
  return 0;
}

// ---------------------------------------------------------------------------------------- FP64 FMA peak (measurement aid)
// Dependent-chain-free DFMA loop: 8 independent accumulators per thread, 2 flops per DFMA.  Gives the denominator for the FP64-pipe
// utilisation of the dealiased advection kernel ("measure with an FMA microbenchmark before quoting utilisation").
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;                  // never true: keeps the loop alive
}
int vk_fp64_peak(Ctx* c, double* tflops) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const int blocks = sms * 8, iters = 1 << 16;
  cudaEvent_t e0, e1;
  NSB_CUDA(cudaEventCreate(&e0));
  NSB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    NSB_CUDA(cudaEventRecord(e0, c->stream));
    k_fp64_peak<<<blocks, 256, 0, c->stream>>>(c->red_out + 12, iters, 0.999999, 1e-9);
    NSB_CUDA(cudaEventRecord(e1, c->stream));
    NSB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;         // rep 0 = warm-up
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  nsb_count_launch(5);
  *tflops = best;
  return 0;
}
