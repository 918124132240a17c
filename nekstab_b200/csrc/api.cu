// api.cu -- the extern "C" surface declared in include/nekstab_b200.h.
#include <cmath>
#include <cstdarg>

#include "nsb_internal.h"

Ctx* g_ctx = nullptr;
static char g_err[1024] = "";
static int g_rank = 0, g_nranks = 1;
static ncclComm_t g_comm = nullptr;
static long long g_launches = 0;

void nsb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void nsb_count_launch(int n) { g_launches += n; if (g_ctx) g_ctx->stats.kernel_launches += n; }

extern "C" const char* nsb_last_error(void) { return g_err; }

#define REQUIRE_CTX()                                              \
  Ctx* c = g_ctx;                                                  \
  if (!c) { nsb_set_error("nsb_init has not been called"); return 1; }

template <class T>
static int dalloc(T** p, long long count) {
  NSB_CUDA(cudaMalloc((void**)p, std::max<long long>(count, 1) * sizeof(T)));
  NSB_CUDA(cudaMemset(*p, 0, std::max<long long>(count, 1) * sizeof(T)));
  return 0;
}
static int h2d(Ctx* c, double* d, const double* h, long long n) {
  NSB_CUDA(cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return 0;
}
static int d2h(Ctx* c, double* h, const double* d, long long n) {
  NSB_CUDA(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------ communicator
extern "C" int nsb_comm_unique_id(char id_out[128]) {
  ncclUniqueId id;
  NSB_NCCL(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId larger than 128 bytes");
  memset(id_out, 0, 128);
  memcpy(id_out, &id, sizeof(id));
  return 0;
}
extern "C" int nsb_comm_init(int rank, int nranks, const char id_in[128], int device) {
  if (g_comm) { ncclCommDestroy(g_comm); g_comm = nullptr; }
  g_rank = rank; g_nranks = nranks;
  if (nranks > 1) {
    NSB_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof(id));
    NSB_NCCL(ncclCommInitRank(&g_comm, nranks, id, rank));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ setup
static void drop_graphs(Ctx* c) {
  for (auto& g : c->graph_p) { if (g.exec) cudaGraphExecDestroy(g.exec); g = Ctx::GraphEntry(); }
  for (auto& g : c->graph_h) { if (g.exec) cudaGraphExecDestroy(g.exec); g = Ctx::GraphEntry(); }
}
static int setup_masks(Ctx* c, int set, const double* m0, const double* m1, const double* m2) {
  drop_graphs(c);
  const double* hm[3] = {m0, m1, m2};
  c->mask_same[set] = true;
  for (int d = 1; d < c->ldim; ++d)
    if (memcmp(hm[0], hm[d], c->n * sizeof(double)) != 0) c->mask_same[set] = false;
  for (int d = 0; d < c->ldim; ++d) {
    NSB_TRY(dalloc(&c->mask[set][d], c->n));
    NSB_TRY(dalloc(&c->mbinv[set][d], c->n));
    NSB_TRY(h2d(c, c->mask[set][d], hm[d], c->n));
    NSB_TRY(vk_copy(c, c->mbinv[set][d], c->mask[set][d], c->n));
    NSB_TRY(vk_mul(c, c->mbinv[set][d], c->binv, c->n));
  }
  NSB_TRY(dalloc(&c->dinvE[set], c->n2));
  NSB_TRY(ek_ediag(c, set));
  NSB_TRY(vk_dot3(c, c->dinvE[set], c->dinvE[set], nullptr, c->n2, c->red_out + 9));   // sum diag(E)^2 (scale for the test below)
  NSB_TRY(vk_inv(c, c->dinvE[set], c->n2));
  // all-Dirichlet / periodic velocity => E * 1 = 0 [UPSTREAM ifvcor]: test numerically
  NSB_TRY(vk_fill(c, c->pk[4], 1.0, c->n2));
  NSB_TRY(ek_gradt(c, c->pk[4], c->wk[2]));
  NSB_TRY(gs_dssum(c, c->wk[2], c->ldim, c->n, nullptr));
  double* sc = nullptr;
  NSB_TRY(dalloc(&sc, c->n * 3));
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(vk_copy(c, sc + d * c->n, c->mbinv[set][d], c->n));
  NSB_TRY(ek_div(c, c->wk[2], sc, c->pk[3], 1.0));
  NSB_TRY(vk_dot3(c, c->pk[3], c->pk[3], nullptr, c->n2, c->red_out + 8));
  NSB_TRY(vk_allreduce_sum(c, c->red_out + 8, 2));
  double hv[2];
  NSB_TRY(d2h(c, hv, c->red_out + 8, 2));
  cudaFree(sc);
  // ||E 1|| against ||diag(E)||: O(1e-2) with an outflow boundary, rounding level without
  c->ifvcor[set] = std::sqrt(hv[0]) < 1e-9 * std::sqrt(hv[1]);
  return 0;
}

extern "C" int nsb_finalize(void) {
  Ctx* c = g_ctx;
  if (!c) return 0;
  cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  gs_free(c);
  double* ptrs[] = {c->xyz[0], c->xyz[1], c->xyz[2], c->R, c->jac, c->bm1, c->binv, c->mult, c->bm1s, c->G, c->RW2, c->bm2inv,
                    c->Rd, c->hdiagA, c->hdiagB, c->dinvH, c->ub0, c->pb0, c->orbit, c->spng_ref, c->spng, c->u, c->ulag[0], c->ulag[1], c->f[0], c->f[1],
                    c->f[2], c->pr, c->prlag, c->pt, c->wk[0], c->wk[1], c->wk[2], c->wk[3], c->rk, c->pk[0], c->pk[1],
                    c->pk[2], c->pk[3], c->pk[4], c->red_part, c->red_out, c->slab, c->hbuf, c->hpart};
  for (double* p : ptrs) if (p) cudaFree(p);
  for (int s = 0; s < 2; ++s) {
    if (s == 1 && !c->has_adj_masks) break;
    for (int d = 0; d < 3; ++d) { if (c->mask[s][d]) cudaFree(c->mask[s][d]); if (c->mbinv[s][d]) cudaFree(c->mbinv[s][d]); }
    if (c->dinvE[s]) cudaFree(c->dinvE[s]);
  }
  if (c->projX) { cudaFree(c->projX); cudaFree(c->projEX); }
  {
    Ctx::StepState& b = c->base_state;
    double* bp[] = {b.u, b.ulag[0], b.ulag[1], b.f[0], b.f[1], b.f[2], b.pr, b.prlag};
    for (double* q : bp) if (q) cudaFree(q);
  }
  pm_free(c->pmg[0]); pm_free(c->pmg[1]);
  {
    Ctx::Scalar& z = c->scal;
    double* sp[] = {z.tmask, z.tb, z.th, z.thlag[0], z.thlag[1], z.q[0], z.q[1], z.q[2], z.wk[0], z.wk[1], z.mats, c->bvec};
    for (double* q : sp) if (q) cudaFree(q);
  }
  if (c->adv_scratch) cudaFree(c->adv_scratch);
  if (c->pz) cudaFree(c->pz);
  if (c->ones2) cudaFree(c->ones2);
  if (c->cgs) cudaFree(c->cgs);
  if (c->cgs_host) cudaFreeHost(c->cgs_host);
  if (c->red_host) cudaFreeHost(c->red_host);
  if (c->red_count) cudaFree(c->red_count);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  g_ctx = nullptr;
  return 0;
}

extern "C" int nsb_init(int ldim, int lx1, int lxd, int lx2, int nelv, long long nelgv, const double* xm1,
                        const double* ym1, const double* zm1, const double* v1mask, const double* v2mask,
                        const double* v3mask, const long long* glo_num, int device) {
  if (g_ctx) nsb_finalize();
  if (ldim != 2 && ldim != 3) { nsb_set_error("ldim must be 2 or 3"); return 1; }
  if (lx2 != lx1 - 2) { nsb_set_error("only the P_N-P_{N-2} formulation (lx2 = lx1-2) is supported, got lx1=%d lx2=%d", lx1, lx2); return 1; }
  if (lxd != 3 * lx1 / 2) { nsb_set_error("lxd must be 3*lx1/2 (3/2-rule dealiasing), got lx1=%d lxd=%d", lx1, lxd); return 1; }
  if (lx1 != 4 && lx1 != 6 && lx1 != 8) { nsb_set_error("lx1 must be 4, 6 or 8 (compiled instantiations)"); return 1; }
  if (nelv <= 0) { nsb_set_error("nelv must be positive"); return 1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    nsb_set_error("no CUDA device available (%s): nekstab_b200 has no CPU fallback", cudaGetErrorString(e));
    return 1;
  }
  NSB_CUDA(cudaSetDevice(device));
  Ctx* c = new Ctx();
  g_ctx = c;
  c->device = device;
  c->ldim = ldim; c->lx1 = lx1; c->lxd = lxd; c->lx2 = lx2; c->nel = nelv; c->nelg = nelgv;
  c->np1 = (ldim == 3) ? lx1 * lx1 * lx1 : lx1 * lx1;
  c->np2 = (ldim == 3) ? lx2 * lx2 * lx2 : lx2 * lx2;
  c->npd = (ldim == 3) ? lxd * lxd * lxd : lxd * lxd;
  c->n = (long long)nelv * c->np1; c->n2 = (long long)nelv * c->np2; c->nd = (long long)nelv * c->npd;
  c->n2_glob = nelgv * c->np2;
  c->vlen = c->n * ldim + c->n2;
  c->poff = c->n * ldim;
  c->rank = g_rank; c->nranks = g_nranks; c->comm = g_comm;
  NSB_CUDA(cudaStreamCreate(&c->stream));
  NSB_CUDA(cudaEventCreate(&c->ev0));
  NSB_CUDA(cudaEventCreate(&c->ev1));
  sem_build_constmats(lx1, lx2, lxd, &c->cm);
  NSB_TRY(ek_upload_constants(c->cm));
  NSB_TRY(pk_upload_constants(c->cm));

  const int D = ldim;
  const double* hx[3] = {xm1, ym1, zm1};
  for (int d = 0; d < D; ++d) {
    if (!hx[d]) { nsb_set_error("coordinate array %d is NULL", d); return 1; }
    NSB_TRY(dalloc(&c->xyz[d], c->n));
    NSB_TRY(h2d(c, c->xyz[d], hx[d], c->n));
  }
  NSB_TRY(dalloc(&c->R, c->n * D * D));
  NSB_TRY(dalloc(&c->jac, c->n));
  NSB_TRY(dalloc(&c->bm1, c->n));
  NSB_TRY(dalloc(&c->binv, c->n));
  NSB_TRY(dalloc(&c->mult, c->n));
  NSB_TRY(dalloc(&c->bm1s, c->n));
  NSB_TRY(dalloc(&c->G, c->n * (D * (D + 1) / 2)));
  NSB_TRY(dalloc(&c->RW2, c->n2 * D * D));
  NSB_TRY(dalloc(&c->bm2inv, c->n2));
  NSB_TRY(dalloc(&c->Rd, c->nd * D * D));
  NSB_TRY(dalloc(&c->hdiagA, c->n));
  NSB_TRY(dalloc(&c->hdiagB, c->n));
  NSB_TRY(dalloc(&c->dinvH, c->n));
  long long nred = std::max<long long>(nelv, NSB_MAX_BLOCKS) * NSB_MAX_RED;
  NSB_TRY(dalloc(&c->red_part, nred));
  NSB_TRY(dalloc(&c->red_out, NSB_MAX_RED * 4));
  NSB_TRY(dalloc(&c->red_count, 4));
  NSB_CUDA(cudaMallocHost((void**)&c->red_host, NSB_MAX_RED * 4 * sizeof(double)));
  NSB_TRY(dalloc(&c->hbuf, 1 << 16));
  NSB_TRY(st_alloc(c));

  NSB_TRY(gs_setup(c, glo_num));
  {   // global ids of the element corners: the vertex (Q1) mesh of the pressure preconditioner (pmg.cu)
    const int NK = (D == 3) ? 8 : 4, N = lx1 - 1;
    c->vglo.resize((size_t)nelv * NK);
    for (int e = 0; e < nelv; ++e)
      for (int k = 0; k < NK; ++k) {
        const int i = (k & 1) ? N : 0, j = ((k >> 1) & 1) ? N : 0, kk = ((k >> 2) & 1) ? N : 0;
        c->vglo[(size_t)e * NK + k] = glo_num[(size_t)e * c->np1 + (size_t)(kk * lx1 + j) * lx1 + i];
      }
  }
  NSB_TRY(ek_geometry(c));
  // positive Jacobian check + volumes
  NSB_TRY(vk_dot3(c, c->bm1, nullptr, nullptr, c->n, c->red_out + 8));
  NSB_TRY(vk_dot3(c, c->bm2inv, nullptr, nullptr, c->n2, c->red_out + 9));   // bm2inv holds bm2 at this point
  NSB_TRY(vk_allreduce_sum(c, c->red_out + 8, 2));
  double hv[2];
  NSB_TRY(d2h(c, hv, c->red_out + 8, 2));
  c->vol = hv[0]; c->vol2 = hv[1];
  if (!(c->vol > 0) || !(c->vol2 > 0)) { nsb_set_error("non-positive mesh volume (%g, %g): check element orientation", c->vol, c->vol2); return 1; }
  NSB_TRY(vk_inv(c, c->bm2inv, c->n2));
  // multiplicity, assembled mass and its inverse, assembled diagonal of A
  NSB_TRY(vk_fill(c, c->mult, 1.0, c->n));
  NSB_TRY(gs_dssum(c, c->mult, 1, c->n, nullptr));
  NSB_TRY(vk_inv(c, c->mult, c->n));
  NSB_TRY(vk_copy(c, c->hdiagB, c->bm1, c->n));
  NSB_TRY(gs_dssum(c, c->hdiagB, 1, c->n, nullptr));
  NSB_TRY(vk_copy(c, c->binv, c->hdiagB, c->n));
  NSB_TRY(vk_inv(c, c->binv, c->n));
  NSB_TRY(gs_dssum(c, c->hdiagA, 1, c->n, nullptr));
  NSB_TRY(vk_copy(c, c->bm1s, c->bm1, c->n));
  if (!v1mask || !v2mask || (D == 3 && !v3mask)) { nsb_set_error("velocity masks must not be NULL"); return 1; }
  NSB_TRY(setup_masks(c, 0, v1mask, v2mask, v3mask));
  for (int d = 0; d < 3; ++d) { c->mask[1][d] = c->mask[0][d]; c->mbinv[1][d] = c->mbinv[0][d]; }
  c->dinvE[1] = c->dinvE[0];
  c->ifvcor[1] = c->ifvcor[0];
  c->mask_same[1] = c->mask_same[0];
  c->has_adj_masks = false;
  // (r1b, measured: a per-element gather fused into k_div3 lost to the separate segment-sorted dssum, 0.465 vs 0.115 + 0.238 ms; removed)
  const char* ng = getenv("NSB_GRAPHS");
  c->use_graphs = !(ng && ng[0] == '0');
  const char* np_ = getenv("NSB_PERSISTENT");
  c->persistent_pcg = !(np_ && np_[0] == '0');
  const char* nt = getenv("NSB_PCG_FUSED");
  c->pcg_fused = !(nt && nt[0] == '0');
  const char* na = getenv("NSB_AX_PERSISTENT");
  c->ax_persistent = !(na && na[0] == '0');
  c->perm_h = c->gsp_ready;                      // surface-first layout of the Helmholtz loop vector (gs.cu builds the second map; NSB_PERM=0: off)
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  const char* pcv = getenv("NSB_PRECOND");
  if (pcv && pcv[0] == '1') NSB_TRY(nsb_set_pressure_preconditioner(1, 0));
  return 0;
}

extern "C" int nsb_set_adjoint_masks(const double* m0, const double* m1, const double* m2) {
  REQUIRE_CTX();
  drop_graphs(c);
  if (c->has_adj_masks) {
    for (int d = 0; d < c->ldim; ++d) { cudaFree(c->mask[1][d]); cudaFree(c->mbinv[1][d]); }
    cudaFree(c->dinvE[1]);
    c->has_adj_masks = false;
  }
  if (!m0) {
    for (int d = 0; d < 3; ++d) { c->mask[1][d] = c->mask[0][d]; c->mbinv[1][d] = c->mbinv[0][d]; }
    c->dinvE[1] = c->dinvE[0];
    c->ifvcor[1] = c->ifvcor[0];
    c->mask_same[1] = c->mask_same[0];
    pm_free(c->pmg[1]);
    return 0;
  }
  for (int d = 0; d < 3; ++d) { c->mask[1][d] = nullptr; c->mbinv[1][d] = nullptr; }
  c->dinvE[1] = nullptr;
  NSB_TRY(setup_masks(c, 1, m0, m1, m2));
  c->has_adj_masks = true;
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->pc_kind) NSB_TRY(pm_setup(c, 1, c->pc_nagg));
  return 0;
}

extern "C" int nsb_set_params(double viscosity, double density, double tol_v, double tol_p, int maxit_v, int maxit_p) {
  REQUIRE_CTX();
  if (!(viscosity > 0) || !(density > 0)) { nsb_set_error("viscosity and density must be positive"); return 1; }
  c->visc = viscosity; c->rho = density; c->tol_v = tol_v; c->tol_p = tol_p;
  if (maxit_v > 0) c->maxit_v = maxit_v;
  if (maxit_p > 0) c->maxit_p = maxit_p;
  return 0;
}
extern "C" int nsb_set_weights(const double* bm1s) {
  REQUIRE_CTX();
  if (bm1s) NSB_TRY(h2d(c, c->bm1s, bm1s, c->n));
  else NSB_TRY(vk_copy(c, c->bm1s, c->bm1, c->n));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int nsb_set_baseflow(const double* u, const double* v, const double* w) {
  REQUIRE_CTX();
  if (!c->ub0) NSB_TRY(dalloc(&c->ub0, c->n * c->ldim));
  c->ub = c->ub0;
  const double* h[3] = {u, v, w};
  for (int d = 0; d < c->ldim; ++d) {
    if (!h[d]) { nsb_set_error("base-flow component %d is NULL", d); return 1; }
    NSB_TRY(h2d(c, c->ub0 + d * c->n, h[d], c->n));
  }
  c->orbit_ready = false;                       // a new base flow invalidates a stored orbit
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
// Sponge forcing of the FULL Navier-Stokes stepper (nonlinear_forward_map, the co-evolving Floquet base flow): the jp = 0 branch of
// nekStab_forcing (core/utils.f:166-171), f += spng_str * spng_fun * (spng_vr - u), spng_vr = the field held at nekStab_init
// (core/utils.f:240).  ur/vr/wr NULL: the current base flow is taken as the reference.  spng_str = 0 switches it off (the default).
extern "C" int nsb_set_dns_sponge(double spng_str, const double* ur, const double* vr, const double* wr) {
  REQUIRE_CTX();
  c->spng_str_dns = spng_str;
  if (spng_str == 0.0) return 0;
  if (!c->spng) { nsb_set_error("nsb_set_dns_sponge: call nsb_set_sponge first (spng_fun)"); return 1; }
  if (!c->spng_ref) NSB_TRY(dalloc(&c->spng_ref, c->n * c->ldim));
  const double* h[3] = {ur, vr, wr};
  if (!ur) {
    if (!c->ub0) { nsb_set_error("nsb_set_dns_sponge: no reference field and no base flow"); return 1; }
    NSB_TRY(vk_copy(c, c->spng_ref, c->ub0, c->n * c->ldim));
  } else {
    for (int d = 0; d < c->ldim; ++d) {
      if (!h[d]) { nsb_set_error("nsb_set_dns_sponge: reference component %d is NULL", d); return 1; }
      NSB_TRY(h2d(c, c->spng_ref + d * c->n, h[d], c->n));
    }
  }
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
// Floquet / UPO analysis (uparam(1) = 3.11 / 3.21, core/matvec.f:192,278): the base flow is advanced with the full Navier-Stokes
// stepper next to the perturbation during the first matvec (Nek's `ifbase`), its orbit uor, vor, wor(lv, nsteps) is kept in HBM and
// replayed by every later matvec (`ifstorebase = .true.`, core/usr_extra.f:24).  pbase (mesh 2, may be NULL = 0): pressure the base
// flow starts from (the P field of the UPO file).  enable = 0 switches back to a frozen base flow and frees the orbit.
extern "C" int nsb_set_floquet(int enable, const double* pbase) {
  REQUIRE_CTX();
  c->floquet = enable != 0;
  c->orbit_ready = false;
  if (pbase) {
    if (!c->pb0) NSB_TRY(dalloc(&c->pb0, c->n2));
    NSB_TRY(h2d(c, c->pb0, pbase, c->n2));
    NSB_CUDA(cudaStreamSynchronize(c->stream));
  } else if (c->pb0) { cudaFree(c->pb0); c->pb0 = nullptr; }
  if (!enable && c->orbit) { cudaFree(c->orbit); c->orbit = nullptr; c->orbit_steps = 0; }
  return 0;
}
// Base-flow orbit snapshot U^{istep} (1 <= istep <= nsteps) after the first Floquet matvec -- what the reference keeps in uor, vor, wor
extern "C" int nsb_get_orbit(int istep, double* u, double* v, double* w) {
  REQUIRE_CTX();
  if (!c->orbit || !c->orbit_ready || istep < 1 || istep > c->orbit_steps) { nsb_set_error("nsb_get_orbit: no stored orbit / step out of range"); return 1; }
  double* h[3] = {u, v, w};
  for (int d = 0; d < c->ldim; ++d)
    if (h[d]) NSB_TRY(d2h(c, h[d], c->orbit + ((long long)(istep - 1) * c->ldim + d) * c->n, c->n));
  return 0;
}
extern "C" int nsb_set_sponge(const double* spng) {
  REQUIRE_CTX();
  if (!spng) {
    if (c->spng) { cudaFree(c->spng); c->spng = nullptr; }
    return 0;
  }
  if (!c->spng) NSB_TRY(dalloc(&c->spng, c->n));
  NSB_TRY(h2d(c, c->spng, spng, c->n));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int nsb_set_timestep(double dt, int nsteps) {
  REQUIRE_CTX();
  if (!(dt > 0) || nsteps <= 0) { nsb_set_error("dt and nsteps must be positive"); return 1; }
  c->dt = dt; c->nsteps = nsteps;
  return 0;
}
extern "C" int nsb_set_step_callback(nsb_step_callback cb, void* user) {
  REQUIRE_CTX();
  c->step_cb = cb;
  c->step_cb_user = user;
  return 0;
}
extern "C" int nsb_set_projection(int mxprev) {
  REQUIRE_CTX();
  if (mxprev < 0 || mxprev > 200) { nsb_set_error("nsb_set_projection: mxprev out of range"); return 1; }
  if (c->projX) { cudaFree(c->projX); cudaFree(c->projEX); c->projX = c->projEX = nullptr; }
  c->proj_max = mxprev; c->proj_m = 0; c->proj_adj = -1;
  if (mxprev > 0) {
    NSB_TRY(dalloc(&c->projX, (long long)mxprev * c->n2));
    NSB_TRY(dalloc(&c->projEX, (long long)mxprev * c->n2));
  }
  return 0;
}
extern "C" int nsb_set_pressure_preconditioner(int kind, int nagg) {
  REQUIRE_CTX();
  if (kind < 0 || kind > 1) { nsb_set_error("nsb_set_pressure_preconditioner: kind must be 0 (Jacobi) or 1 (three-level Schwarz/multilevel)"); return 1; }
  if (nagg < 0 || nagg > 4096) { nsb_set_error("nsb_set_pressure_preconditioner: nagg must be in [0, 4096]"); return 1; }
  drop_graphs(c);
  c->pc_kind = 0;
  if (kind == 0) return 0;
  if (!c->pz) {
    NSB_TRY(dalloc(&c->pz, c->n2));
    NSB_TRY(dalloc(&c->ones2, c->n2));
    NSB_TRY(vk_fill(c, c->ones2, 1.0, c->n2));
  }
  c->pc_nagg = nagg;
  NSB_TRY(pm_setup(c, 0, nagg));
  if (c->has_adj_masks) NSB_TRY(pm_setup(c, 1, nagg));     // separate adjoint masks => a different E => its own factors
  c->pc_kind = kind;
  return 0;
}
extern "C" int nsb_op_pc_apply(int adjoint, const double* r, double* z) {
  REQUIRE_CTX();
  if (!c->pz || !c->pmg[0].ready) { nsb_set_error("nsb_op_pc_apply: call nsb_set_pressure_preconditioner(1, ..) first"); return 1; }
  NSB_TRY(h2d(c, c->pk[3], r, c->n2));
  NSB_TRY(pm_apply(c, adjoint ? 1 : 0, c->pk[3], c->pz, 0));
  return d2h(c, z, c->pz, c->n2);
}
extern "C" int nsb_pc_get(int which, double* out, long long* count) {
  REQUIRE_CTX();
  const PMG& m = c->pmg[0];
  if (!m.ready) { nsb_set_error("nsb_pc_get: preconditioner not set up"); return 1; }
  long long n = 0;
  switch (which) {
    case 0: n = (long long)m.h_agg.size(); if (out) for (long long i = 0; i < n; ++i) out[i] = m.h_agg[i]; break;
    case 1: n = (long long)m.h_d1.size(); if (out) for (long long i = 0; i < n; ++i) out[i] = m.h_d1[i]; break;
    case 2: n = (long long)m.h_A2inv.size(); if (out) for (long long i = 0; i < n; ++i) out[i] = m.h_A2inv[i]; break;
    case 3: n = 5; if (out) { out[0] = m.nv; out[1] = m.nagg; out[2] = m.ncolours; out[3] = m.nagg_loc; out[4] = m.agg_first; } break;
    default: nsb_set_error("nsb_pc_get: bad selector %d", which); return 1;
  }
  if (count) *count = n;
  return 0;
}
extern "C" int nsb_set_ifvcor(int direct, int adjoint) {
  REQUIRE_CTX();
  if (direct >= 0) c->ifvcor[0] = direct != 0;
  if (adjoint >= 0) c->ifvcor[1] = adjoint != 0;
  else if (!c->has_adj_masks && direct >= 0) c->ifvcor[1] = c->ifvcor[0];
  return 0;
}
extern "C" int nsb_prepare_linearized_solver(double end_time, double cfl_target, double* dt, int* nsteps, double* ctarg) {
  REQUIRE_CTX();
  if (!c->ub) { nsb_set_error("base flow not set"); return 1; }
  if (cfl_target > 1.0) cfl_target = 0.5;   // core/matvec.f:21-24
  NSB_TRY(ek_cfl(c, c->ub, c->red_out + 8));
  NSB_TRY(vk_allreduce_max(c, c->red_out + 8, 1));
  double ct;
  NSB_TRY(d2h(c, &ct, c->red_out + 8, 1));
  if (!(ct > 0)) { nsb_set_error("compute_cfl returned %g", ct); return 1; }
  double dt0 = cfl_target / ct;
  int ns = (int)std::ceil(end_time / dt0);
  c->dt = end_time / ns; c->nsteps = ns;
  if (dt) *dt = c->dt;
  if (nsteps) *nsteps = ns;
  if (ctarg) *ctarg = ct;
  return 0;
}

// ------------------------------------------------------------------------------------------ vectors
#define CHECK_SLOT(s)                                                                     \
  if ((s) < 0 || (s) >= c->nslots) { nsb_set_error("slot %d out of range [0,%d)", (s), c->nslots); return 1; }

// h_j = <Q_j, f> under bm1s over the velocity -- and the scalar when it travels with the vector (core/krylov_subspace.f:37-45)
static int multidot(Ctx* c, int k, int first, int slot_f, double* h_dev) {
  if (!c->scal.on) return vk_multidot(c, k, first, slot_f, h_dev);
  return vk_multidot_raw(c, k, slot_ptr(c, first), c->vlen, slot_ptr(c, slot_f), c->bm1s, c->n, c->n * (c->ldim + 1), h_dev);
}

extern "C" int nsb_vec_alloc(int nslots) {
  REQUIRE_CTX();
  if (nslots <= 0) { nsb_set_error("nslots must be positive"); return 1; }
  if (c->slab) cudaFree(c->slab);
  c->slab = nullptr;
  NSB_TRY(dalloc(&c->slab, (long long)nslots * c->vlen));
  c->nslots = nslots;
  c->slot_time.assign(nslots, 0.0);
  return 0;
}
extern "C" int nsb_vec_upload(int slot, const double* vx, const double* vy, const double* vz, const double* pr) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  double* v = slot_ptr(c, slot);
  const double* h[3] = {vx, vy, vz};
  for (int d = 0; d < c->ldim; ++d) {
    if (!h[d]) { nsb_set_error("vec_upload: component %d is NULL", d); return 1; }
    NSB_TRY(h2d(c, v + d * c->n, h[d], c->n));
  }
  if (pr) NSB_TRY(h2d(c, v + c->poff, pr, c->n2));
  else NSB_TRY(vk_fill(c, v + c->poff, 0.0, c->n2));
  if (c->scal.on) NSB_TRY(vk_fill(c, v + c->ldim * c->n, 0.0, c->n));       // theta: nsb_vec_upload_scalar
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  c->slot_time[slot] = 0.0;
  return 0;
}
// q%time of a slot (core/krylov_subspace.f:14): the period unknown of the UPO Newton; part of the inner product only when nsb_set_upo(1)
extern "C" int nsb_vec_set_time(int slot, double t) { REQUIRE_CTX(); CHECK_SLOT(slot); c->slot_time[slot] = t; return 0; }
extern "C" int nsb_vec_get_time(int slot, double* t) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  if (!t) { nsb_set_error("nsb_vec_get_time: NULL output"); return 1; }
  *t = c->slot_time[slot];
  return 0;
}
// uparam(1) = 2.1 (Newton-GMRES for unstable periodic orbits): the time component enters krylov_inner_product (core/krylov_subspace.f:47-50),
// nsb_nonlinear_forward_map keeps the orbit uor, vor, wor in HBM (core/newton_krylov.f:364-368) together with the two border vectors
// of compute_bvec, and the Newton matvec adds the border terms (core/matvec.f:407-419).
extern "C" int nsb_set_upo(int enable) {
  REQUIRE_CTX();
  c->upo = enable != 0;
  c->bvec_ready = false;
  c->orbit_ready = false;
  if (!enable && c->bvec) { cudaFree(c->bvec); c->bvec = nullptr; }
  if (!enable && !c->floquet && c->orbit) { cudaFree(c->orbit); c->orbit = nullptr; c->orbit_steps = 0; }
  return 0;
}
bool upo_active() { return g_ctx && g_ctx->upo; }

// Scalar transport (`ifheat`, ldimt = 1).  After nsb_set_scalar(1, ..) every Krylov vector is [vx|vy|(vz)|theta|pr]
// (core/krylov_subspace.f:8-15), theta enters the inner product with the weight bm1s (:41-45) and is advanced by the direct and the
// full Navier-Stokes maps [UPSTREAM perturb.f heatp / convabp; heat / convab].  conductivity = param(8), rhocp = param(7); tmask = the
// scalar's Dirichlet mask; ri: buoyancy f_gdir += ri * theta (uparam(6) in the shipped Boussinesq .usr files), gdir 0-based.
// The layout change discards the slots: call nsb_vec_alloc afterwards.
extern "C" int nsb_set_scalar(int enable, double conductivity, double rhocp, const double* tmask, double ri, int gdir) {
  REQUIRE_CTX();
  Ctx::Scalar& z = c->scal;
  drop_graphs(c);
  if (c->slab) { cudaFree(c->slab); c->slab = nullptr; c->nslots = 0; c->slot_time.clear(); }
  z.on = enable != 0;
  c->poff = c->n * (c->ldim + (z.on ? 1 : 0));
  c->vlen = c->poff + c->n2;
  if (c->bvec) { cudaFree(c->bvec); c->bvec = nullptr; c->bvec_ready = false; }
  if (!z.on) return 0;
  if (!tmask) { nsb_set_error("nsb_set_scalar: tmask is NULL"); return 1; }
  if (gdir < 0 || gdir >= c->ldim) { nsb_set_error("nsb_set_scalar: gdir %d out of range", gdir); return 1; }
  if (!(rhocp > 0) || conductivity < 0) { nsb_set_error("nsb_set_scalar: rhocp must be positive, conductivity non-negative"); return 1; }
  z.cond = conductivity; z.rhocp = rhocp; z.ri = ri; z.gdir = gdir;
  if (!z.tmask) {
    NSB_TRY(dalloc(&z.tmask, c->n)); NSB_TRY(dalloc(&z.th, c->n));
    for (int j = 0; j < 2; ++j) { NSB_TRY(dalloc(&z.thlag[j], c->n)); NSB_TRY(dalloc(&z.wk[j], c->n)); }
    for (int j = 0; j < 3; ++j) NSB_TRY(dalloc(&z.q[j], c->n));
  }
  NSB_TRY(h2d(c, z.tmask, tmask, c->n));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return sk_setup(c);
}
// tbase (core/NEKSTAB /nStab_bflows/, loaded at core/eigensolvers.f:195-199): the scalar field the perturbation is linearised about
extern "C" int nsb_set_scalar_base(const double* tbase) {
  REQUIRE_CTX();
  if (!c->scal.on) { nsb_set_error("nsb_set_scalar_base: call nsb_set_scalar(1, ..) first"); return 1; }
  if (!tbase) { nsb_set_error("nsb_set_scalar_base: NULL field"); return 1; }
  if (!c->scal.tb) NSB_TRY(dalloc(&c->scal.tb, c->n));
  NSB_TRY(h2d(c, c->scal.tb, tbase, c->n));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int nsb_vec_upload_scalar(int slot, const double* theta) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  if (!c->scal.on || !theta) { nsb_set_error("nsb_vec_upload_scalar: scalar transport is off or NULL field"); return 1; }
  NSB_TRY(h2d(c, slot_ptr(c, slot) + c->ldim * c->n, theta, c->n));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int nsb_vec_download_scalar(int slot, double* theta) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  if (!c->scal.on || !theta) { nsb_set_error("nsb_vec_download_scalar: scalar transport is off or NULL field"); return 1; }
  return d2h(c, theta, slot_ptr(c, slot) + c->ldim * c->n, c->n);
}
extern "C" int nsb_vec_download(int slot, double* vx, double* vy, double* vz, double* pr) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  double* v = slot_ptr(c, slot);
  double* h[3] = {vx, vy, vz};
  for (int d = 0; d < c->ldim; ++d)
    if (h[d]) NSB_TRY(d2h(c, h[d], v + d * c->n, c->n));
  if (pr) NSB_TRY(d2h(c, pr, v + c->poff, c->n2));
  return 0;
}
extern "C" int nsb_vec_copy(int dst, int src) {
  REQUIRE_CTX(); CHECK_SLOT(dst); CHECK_SLOT(src);
  if (dst != src) NSB_TRY(vk_copy(c, slot_ptr(c, dst), slot_ptr(c, src), c->vlen));
  c->slot_time[dst] = c->slot_time[src];
  return 0;
}
extern "C" int nsb_vec_zero(int slot) { REQUIRE_CTX(); CHECK_SLOT(slot); c->slot_time[slot] = 0.0; return vk_fill(c, slot_ptr(c, slot), 0.0, c->vlen); }
extern "C" int nsb_vec_cmult(int slot, double a) { REQUIRE_CTX(); CHECK_SLOT(slot); c->slot_time[slot] *= a; return vk_scale(c, slot_ptr(c, slot), a, c->vlen); }
extern "C" int nsb_vec_add2(int p, int q) {
  REQUIRE_CTX(); CHECK_SLOT(p); CHECK_SLOT(q);
  c->slot_time[p] += c->slot_time[q];
  return vk_axpy(c, slot_ptr(c, p), 1.0, slot_ptr(c, q), c->vlen);
}
extern "C" int nsb_vec_sub2(int p, int q) {
  REQUIRE_CTX(); CHECK_SLOT(p); CHECK_SLOT(q);
  c->slot_time[p] -= c->slot_time[q];
  return vk_axpy(c, slot_ptr(c, p), -1.0, slot_ptr(c, q), c->vlen);
}

extern "C" int nsb_vec_inner_product(int p, int q, double* alpha) {
  REQUIRE_CTX(); CHECK_SLOT(p); CHECK_SLOT(q);
  NSB_TRY(multidot(c, 1, q, p, c->hbuf));
  NSB_TRY(d2h(c, alpha, c->hbuf, 1));
  if (c->upo) *alpha += c->slot_time[p] * c->slot_time[q];          // time component (core/krylov_subspace.f:47-50)
  if (std::isnan(*alpha)) { nsb_set_error("krylov_inner_product: NaN (core/krylov_subspace.f:53 -> nek_end)"); return 2; }
  return 0;
}
extern "C" int nsb_vec_norm(int p, double* alpha) {
  NSB_TRY(nsb_vec_inner_product(p, p, alpha));
  *alpha = std::sqrt(*alpha);
  return 0;
}
extern "C" int nsb_vec_normalize(int p, double* alpha) {
  NSB_TRY(nsb_vec_norm(p, alpha));
  return nsb_vec_cmult(p, 1.0 / *alpha);
}
extern "C" int nsb_basis_gemv(int k, int first, const double* y, int out) {
  REQUIRE_CTX(); CHECK_SLOT(first); CHECK_SLOT(first + k - 1); CHECK_SLOT(out);
  if (out >= first && out < first + k) { nsb_set_error("basis_gemv: output slot inside the basis range"); return 1; }
  NSB_TRY(h2d(c, c->hbuf, y, k));
  double t = 0.0;
  for (int i = 0; i < k; ++i) t += y[i] * c->slot_time[first + i];
  c->slot_time[out] = t;
  return vk_gemv_out(c, k, first, c->hbuf, out);
}
extern "C" int nsb_basis_gemv_complex(int k, int first, const double* yre, const double* yim, int sre, int sim) {
  NSB_TRY(nsb_basis_gemv(k, first, yre, sre));
  return nsb_basis_gemv(k, first, yim, sim);
}
extern "C" int nsb_basis_rotate(int k, int first, const double* S, int lds) {
  REQUIRE_CTX(); CHECK_SLOT(first); CHECK_SLOT(first + k - 1);
  double* dS = nullptr;
  NSB_TRY(dalloc(&dS, (long long)lds * k));
  NSB_TRY(h2d(c, dS, S, (long long)lds * k));
  int rc = vk_rotate(c, k, first, dS, lds);
  cudaStreamSynchronize(c->stream);
  cudaFree(dS);
  std::vector<double> t(k, 0.0);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) t[j] += c->slot_time[first + i] * S[(size_t)j * lds + i];
  for (int j = 0; j < k; ++j) c->slot_time[first + j] = t[j];
  return rc;
}
extern "C" int nsb_orthonormalize(int k, int first, int slot_f, double* hcol) {
  REQUIRE_CTX(); CHECK_SLOT(first); CHECK_SLOT(first + k - 1); CHECK_SLOT(slot_f);
  if (k + 1 > (1 << 15)) { nsb_set_error("orthonormalize: k too large"); return 1; }
  double* h1 = c->hbuf;
  double* h2 = c->hbuf + (1 << 15);
  std::vector<double> a(k), b(k);
  if (c->upo) {
    // vectors with a time component: the coefficients need t_i * t_f from the host before the update, so the two passes make a
    // host round trip each (the UPO Newton basis is small next to the ~1e3-step matvec it orthogonalises)
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<double>& h = pass ? b : a;
      double* hd = pass ? h2 : h1;
      NSB_TRY(multidot(c, k, first, slot_f, hd));
      NSB_TRY(d2h(c, h.data(), hd, k));
      double tf = c->slot_time[slot_f];
      for (int i = 0; i < k; ++i) h[i] += c->slot_time[first + i] * tf;
      for (int i = 0; i < k; ++i) tf -= h[i] * c->slot_time[first + i];
      NSB_TRY(h2d(c, hd, h.data(), k));
      NSB_TRY(vk_multiaxpy(c, k, first, slot_f, hd, -1.0));
      c->slot_time[slot_f] = tf;
    }
    for (int i = 0; i < k; ++i) {
      hcol[i] = a[i] + b[i];
      if (std::isnan(hcol[i])) { nsb_set_error("orthonormalize: NaN coefficient"); return 2; }
    }
    return nsb_vec_normalize(slot_f, &hcol[k]);
  }
  const bool pr = c->prof_on != 0;                          // sampling profiler: kinds 11 (multidot) and 12 (multiaxpy)
  if (pr) cudaEventRecord(c->prof_ev[16], c->stream);
  NSB_TRY(multidot(c, k, first, slot_f, h1));            // h1 = Q^T W f
  if (pr) cudaEventRecord(c->prof_ev[17], c->stream);
  NSB_TRY(vk_multiaxpy(c, k, first, slot_f, h1, -1.0));     // f -= Q h1
  if (pr) cudaEventRecord(c->prof_ev[18], c->stream);
  NSB_TRY(multidot(c, k, first, slot_f, h2));            // re-orthogonalisation (DGKS)
  if (pr) cudaEventRecord(c->prof_ev[19], c->stream);
  NSB_TRY(vk_multiaxpy(c, k, first, slot_f, h2, -1.0));
  if (pr) cudaEventRecord(c->prof_ev[20], c->stream);
  NSB_TRY(d2h(c, a.data(), h1, k));
  if (pr) {
    for (int i = 0; i < 4; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, c->prof_ev[16 + i], c->prof_ev[17 + i]) == cudaSuccess) { c->prof_ms[11 + (i & 1)] += ms; c->prof_cnt[11 + (i & 1)] += 1; }
    }
  }
  NSB_TRY(d2h(c, b.data(), h2, k));
  for (int i = 0; i < k; ++i) {
    hcol[i] = a[i] + b[i];
    if (std::isnan(hcol[i])) { nsb_set_error("orthonormalize: NaN coefficient"); return 2; }
  }
  return nsb_vec_normalize(slot_f, &hcol[k]);
}

// ------------------------------------------------------------------------------------------ direct / adjoint mode post-processing
// biorthogonalize (core/sensitivity.f:428-504) on four device-resident modes: the direct mode is scaled to unit norm, the adjoint
// mode rotated / scaled so that <a, d> = 1 under the bm1s inner product (core/eigensolvers.f:7-58).
extern "C" int nsb_biorthogonalize(int dre, int dim, int are, int aim) {
  REQUIRE_CTX(); CHECK_SLOT(dre); CHECK_SLOT(dim); CHECK_SLOT(are); CHECK_SLOT(aim);
  if (dre == dim || are == aim || dre == are || dre == aim || dim == are || dim == aim) { nsb_set_error("biorthogonalize: the four slots must differ"); return 1; }
  double alpha = 0, beta = 0;
  NSB_TRY(nsb_vec_inner_product(dre, dre, &alpha));                    // :463-467
  NSB_TRY(nsb_vec_inner_product(dim, dim, &beta));
  if (!(alpha + beta > 0)) { nsb_set_error("biorthogonalize: zero direct mode"); return 1; }
  double gamma = 1.0 / std::sqrt(alpha + beta);
  NSB_TRY(nsb_vec_cmult(dre, gamma));                                  // :471-472 (the reference's opcmult leaves the pressure as is;
  NSB_TRY(nsb_vec_cmult(dim, gamma));                                  //  the pressure part is not used downstream)
  NSB_TRY(nsb_vec_inner_product(are, dre, &alpha));                    // :475-477
  NSB_TRY(nsb_vec_inner_product(aim, dim, &beta));
  gamma = alpha + beta;
  NSB_TRY(nsb_vec_inner_product(are, dim, &alpha));                    // :479-481
  NSB_TRY(nsb_vec_inner_product(aim, dre, &beta));
  const double delta = alpha - beta;
  if (!(gamma * gamma + delta * delta > 0)) { nsb_set_error("biorthogonalize: direct and adjoint modes are orthogonal"); return 1; }
  return vk_rotate_pair(c, slot_ptr(c, are), slot_ptr(c, aim), gamma, delta, c->vlen);      // :484-501
}
// wave_maker (core/sensitivity.f:7-81): bi-orthonormalise, then |u_direct| |u_adjoint| pointwise; the field (n values, the array the
// reference outposts as temperature in wm_<session>0.f00001) is returned to the host.
extern "C" int nsb_wave_maker(int dre, int dim, int are, int aim, double* wavemaker) {
  NSB_TRY(nsb_biorthogonalize(dre, dim, are, aim));
  Ctx* c = g_ctx;
  if (!wavemaker) { nsb_set_error("wave_maker: output array is NULL"); return 1; }
  NSB_TRY(vk_wavemaker(c, slot_ptr(c, dre), slot_ptr(c, dim), slot_ptr(c, are), slot_ptr(c, aim), c->wk[0]));
  return d2h(c, wavemaker, c->wk[0], c->n);
}

// ------------------------------------------------------------------------------------------ matvec
extern "C" int nsb_matvec(int mode, int sin, int sout) {
  REQUIRE_CTX(); CHECK_SLOT(sin); CHECK_SLOT(sout);
  if (sin == sout) { nsb_set_error("matvec: input and output slots must differ"); return 1; }
  double* q = slot_ptr(c, sin);
  double* f = slot_ptr(c, sout);
  switch (mode) {
    case NSB_DIRECT: return st_linearized_map(c, 0, q, f);
    case NSB_ADJOINT: return st_linearized_map(c, 1, q, f);
    case NSB_DIRECT_ADJOINT: {                       // transient_growth_map: core/matvec.f:332-349
      NSB_TRY(st_linearized_map(c, 0, q, f));
      // wrk lives in the stepper's own state: copy f to a scratch vector first
      double* tmp = nullptr;
      NSB_TRY(dalloc(&tmp, c->vlen));
      NSB_TRY(vk_copy(c, tmp, f, c->vlen));
      int rc = st_linearized_map(c, 1, tmp, f);
      cudaStreamSynchronize(c->stream);
      cudaFree(tmp);
      return rc;
    }
    case NSB_NEWTON:                                 // (exp(TL) - I) q : core/matvec.f:397-400
      NSB_TRY(st_linearized_map(c, 0, q, f));
      NSB_TRY(vk_axpy(c, f, -1.0, q, c->vlen));
      c->slot_time[sout] = 0.0;                      // :421
      if (c->upo) {                                  // Newton for UPOs, :407-419
        if (!c->bvec_ready) { nsb_set_error("UPO Newton matvec: call nsb_nonlinear_forward_map first (it computes the border vectors)"); return 1; }
        NSB_TRY(vk_axpy(c, f, c->slot_time[sin], c->bvec, c->vlen));                                  // f += bvec(fc_nwt) * q%time
        NSB_TRY(vk_multidot_raw(c, 1, c->bvec + c->vlen, c->vlen, q, c->bm1s, c->n, c->n * c->ldim, c->hbuf));
        NSB_TRY(d2h(c, &c->slot_time[sout], c->hbuf, 1));                                             // f%time = <bvec(ic_nwt), q>
      }
      return 0;
    case NSB_FORCE_SENS:                             // (I - exp(TL+)) q : core/matvec.f:366-371
      NSB_TRY(st_linearized_map(c, 1, q, f));
      NSB_TRY(vk_axpy(c, f, -1.0, q, c->vlen));
      return vk_scale(c, f, -1.0, c->vlen);
  }
  nsb_set_error("matvec: unknown mode %d", mode);
  return 1;
}

// compute_bvec (core/matvec.f:435-475): one first-order step of the full Navier-Stokes stepper from qbase approximates its time derivative
static int compute_bvec(Ctx* c, const double* qbase, double* out) {
  const int ns = c->nsteps;
  const bool upo = c->upo;
  c->nsteps = 1; c->upo = false;                    // a single step, no orbit storage
  int rc = st_linearized_map(c, 2, qbase, out);
  c->nsteps = ns; c->upo = upo;
  if (rc) return rc;
  NSB_TRY(vk_axpy(c, out, -1.0, qbase, c->vlen));
  return vk_scale(c, out, 1.0 / c->dt, c->vlen);
}
// nonlinear_forward_map (core/newton_krylov.f:336-378): f = phi_T(q) - q with the full Navier-Stokes stepper (q carries its
// Dirichlet data); afterwards q becomes the base flow of the linearised maps (ubase <- q, :374-375).
extern "C" int nsb_nonlinear_forward_map(int sq, int sf) {
  REQUIRE_CTX(); CHECK_SLOT(sq); CHECK_SLOT(sf);
  if (sq == sf) { nsb_set_error("nonlinear_forward_map: input and output slots must differ"); return 1; }
  double* q = slot_ptr(c, sq);
  double* f = slot_ptr(c, sf);
  NSB_TRY(st_linearized_map(c, 2, q, f));           // upo: the orbit U^1..U^nsteps is stored on the way (:364-368)
  if (c->upo) {                                     // ic_nwt = q, fc_nwt = phi_T(q) (:346, :373): their border vectors, once per Newton iterate
    if (!c->bvec) NSB_TRY(dalloc(&c->bvec, 2 * c->vlen));
    NSB_TRY(compute_bvec(c, f, c->bvec));
    NSB_TRY(compute_bvec(c, q, c->bvec + c->vlen));
    c->bvec_ready = true;
  }
  NSB_TRY(vk_axpy(c, f, -1.0, q, c->vlen));
  c->slot_time[sf] = 0.0;                           // :375
  if (!c->ub0) NSB_TRY(dalloc(&c->ub0, c->n * c->ldim));
  c->ub = c->ub0;
  if (!c->upo) c->orbit_ready = false;
  if (c->scal.on) {                                 // tbase <- q%theta (:375)
    if (!c->scal.tb) NSB_TRY(dalloc(&c->scal.tb, c->n));
    NSB_TRY(vk_copy(c, c->scal.tb, q + c->n * c->ldim, c->n));
  }
  return vk_copy(c, c->ub0, q, c->n * c->ldim);
}
// prepare_linearized_solver on the velocity of a Krylov vector instead of the stored base flow (newton_krylov calls it on
// the current Newton iterate, core/newton_krylov.f:69 with vx,vy,vz = q)
extern "C" int nsb_prepare_solver_from_slot(int slot, double end_time, double cfl_target, double* dt, int* nsteps, double* ctarg) {
  REQUIRE_CTX(); CHECK_SLOT(slot);
  if (cfl_target > 1.0) cfl_target = 0.5;
  NSB_TRY(ek_cfl(c, slot_ptr(c, slot), c->red_out + 8));
  NSB_TRY(vk_allreduce_max(c, c->red_out + 8, 1));
  double ct;
  NSB_TRY(d2h(c, &ct, c->red_out + 8, 1));
  if (!(ct > 0)) { nsb_set_error("compute_cfl returned %g", ct); return 1; }
  int ns = (int)std::ceil(end_time / (cfl_target / ct));
  c->dt = end_time / ns; c->nsteps = ns;
  if (dt) *dt = c->dt;
  if (nsteps) *nsteps = ns;
  if (ctarg) *ctarg = ct;
  return 0;
}
extern "C" int nsb_get_stats(nsb_stats* out, int reset) {
  REQUIRE_CTX();
  if (out) *out = c->stats;
  if (reset) c->stats = nsb_stats{0, 0, 0, 0, 0.0};
  return 0;
}
extern "C" int nsb_profile(int enable, double* ms_sum, long long* count) {
  REQUIRE_CTX();
  if (enable > 0 && !c->prof_ev[0])
    for (int i = 0; i < 24; ++i) NSB_CUDA(cudaEventCreate(&c->prof_ev[i]));
  if (ms_sum) for (int i = 0; i < 16; ++i) ms_sum[i] = c->prof_ms[i];
  if (count) for (int i = 0; i < 16; ++i) count[i] = c->prof_cnt[i];
  if (enable >= 0) {
    c->prof_on = enable;
    for (int i = 0; i < 16; ++i) { c->prof_ms[i] = 0; c->prof_cnt[i] = 0; }
  }
  return 0;
}
// FP64 FMA throughput of this GPU (TFLOP/s), measured with a register-resident DFMA loop: the roofline denominator of the advection kernel
extern "C" int nsb_fp64_peak(double* tflops) {
  REQUIRE_CTX();
  if (!tflops) { nsb_set_error("nsb_fp64_peak: NULL output"); return 1; }
  return vk_fp64_peak(c, tflops);
}
extern "C" long long nsb_n(void) { return g_ctx ? g_ctx->n : 0; }
extern "C" long long nsb_n2(void) { return g_ctx ? g_ctx->n2 : 0; }

// ------------------------------------------------------------------------------------------ operator-level entry points
// out = B (a.grad) phi, dealiased (convop; needs nsb_set_scalar(1, ..) for the kernel's matrices)
extern "C" int nsb_op_conv_scalar(const double* ax, const double* ay, const double* az, const double* phi, double* out) {
  REQUIRE_CTX();
  if (!c->scal.on) { nsb_set_error("nsb_op_conv_scalar: call nsb_set_scalar(1, ..) first"); return 1; }
  const double* h[3] = {ax, ay, az};
  for (int d = 0; d < c->ldim; ++d) {
    if (!h[d]) { nsb_set_error("nsb_op_conv_scalar: component %d is NULL", d); return 1; }
    NSB_TRY(h2d(c, c->wk[0] + d * c->n, h[d], c->n));
  }
  NSB_TRY(h2d(c, c->scal.wk[0], phi, c->n));
  NSB_TRY(sk_conv(c, c->wk[0], c->scal.wk[0], nullptr, nullptr, c->scal.wk[1]));
  return d2h(c, out, c->scal.wk[1], c->n);
}
extern "C" int nsb_op_axhelm(const double* u, double h1, double h2, double* w) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->wk[0], u, c->n));
  NSB_TRY(ek_axhelm(c, c->wk[0], c->wk[1], 1, h1, h2));
  return d2h(c, w, c->wk[1], c->n);
}
extern "C" int nsb_op_dssum(double* u) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->wk[0], u, c->n));
  NSB_TRY(gs_dssum(c, c->wk[0], 1, c->n, nullptr));
  return d2h(c, u, c->wk[0], c->n);
}
extern "C" int nsb_op_glsc3(const double* a, const double* b, const double* w, double* out) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->wk[0], a, c->n));
  NSB_TRY(h2d(c, c->wk[0] + c->n, b, c->n));
  NSB_TRY(h2d(c, c->wk[1], w, c->n));
  NSB_TRY(vk_dot3(c, c->wk[0], c->wk[0] + c->n, c->wk[1], c->n, c->red_out + 8));
  NSB_TRY(vk_allreduce_sum(c, c->red_out + 8, 1));
  return d2h(c, out, c->red_out + 8, 1);
}
extern "C" int nsb_op_opgradt(const double* p, double* wx, double* wy, double* wz) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->pk[4], p, c->n2));
  NSB_TRY(ek_gradt(c, c->pk[4], c->wk[0]));
  double* h[3] = {wx, wy, wz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(d2h(c, h[d], c->wk[0] + d * c->n, c->n));
  return 0;
}
extern "C" int nsb_op_opdiv(const double* ux, const double* uy, const double* uz, double* q) {
  REQUIRE_CTX();
  const double* h[3] = {ux, uy, uz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(h2d(c, c->wk[0] + d * c->n, h[d], c->n));
  NSB_TRY(ek_div(c, c->wk[0], nullptr, c->pk[4], 1.0));
  return d2h(c, q, c->pk[4], c->n2);
}
extern "C" int nsb_op_cdabdtp(const double* p, double* ep) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->pk[4], p, c->n2));
  NSB_TRY(ek_gradt(c, c->pk[4], c->wk[2]));
  NSB_TRY(gs_dssum(c, c->wk[2], c->ldim, c->n, nullptr));
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(vk_mul(c, c->wk[2] + d * c->n, c->mbinv[0][d], c->n));
  NSB_TRY(ek_div(c, c->wk[2], nullptr, c->pk[3], 1.0));
  return d2h(c, ep, c->pk[3], c->n2);
}
extern "C" int nsb_op_advab(int adjoint, const double* upx, const double* upy, const double* upz, double* fx, double* fy,
                            double* fz) {
  REQUIRE_CTX();
  if (!c->ub) { nsb_set_error("base flow not set"); return 1; }
  const double* h[3] = {upx, upy, upz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(h2d(c, c->wk[0] + d * c->n, h[d], c->n));
  NSB_TRY(ek_advab(c, adjoint, c->wk[0], c->ub, c->spng, c->wk[1]));
  NSB_TRY(vk_scale(c, c->wk[1], -1.0, c->n * c->ldim));    // report +B[...] (+ sponge term) like advabp's ta arrays
  double* o[3] = {fx, fy, fz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(d2h(c, o[d], c->wk[1] + d * c->n, c->n));
  return 0;
}
extern "C" int nsb_op_hmholtz(double* ux, double* uy, double* uz, const double* rx, const double* ry, const double* rz,
                              double h1, double h2, int* iters) {
  REQUIRE_CTX();
  const double* h[3] = {rx, ry, rz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(h2d(c, c->rk + d * c->n, h[d], c->n));
  NSB_TRY(gs_dssum(c, c->rk, c->ldim, c->n, nullptr));
  NSB_TRY(vk_mask_fields(c, c->rk, 0));
  NSB_TRY(st_helmholtz(c, 0, h1, h2, iters));
  double* o[3] = {ux, uy, uz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(d2h(c, o[d], c->wk[3] + d * c->n, c->n));
  return 0;
}
extern "C" int nsb_op_esolver(const double* gin, double* phi, int* iters) {
  REQUIRE_CTX();
  NSB_TRY(h2d(c, c->pk[0], gin, c->n2));
  NSB_TRY(st_pressure(c, 0, iters));
  return d2h(c, phi, c->pk[1], c->n2);
}
extern "C" int nsb_op_cfl(const double* ux, const double* uy, const double* uz, double dt, double* cfl) {
  REQUIRE_CTX();
  const double* h[3] = {ux, uy, uz};
  for (int d = 0; d < c->ldim; ++d) NSB_TRY(h2d(c, c->wk[0] + d * c->n, h[d], c->n));
  NSB_TRY(ek_cfl(c, c->wk[0], c->red_out + 8));
  NSB_TRY(vk_allreduce_max(c, c->red_out + 8, 1));
  NSB_TRY(d2h(c, cfl, c->red_out + 8, 1));
  *cfl *= dt;
  return 0;
}
extern "C" int nsb_get_field(const char* name, double* out, long long* count) {
  REQUIRE_CTX();
  std::string s(name);
  const double* src = nullptr;
  long long cnt = c->n;
  bool invert = false;
  if (s == "bm1") src = c->bm1;
  else if (s == "binvm1") src = c->binv;
  else if (s == "jacm1") src = c->jac;
  else if (s == "vmult") src = c->mult;
  else if (s == "bm1s") src = c->bm1s;
  else if (s == "hdiagA") src = c->hdiagA;
  else if (s == "bm2") { src = c->bm2inv; cnt = c->n2; invert = true; }
  else if (s == "ediag") { src = c->dinvE[0]; cnt = c->n2; invert = true; }
  else if (s == "ediag_adj") { src = c->dinvE[1]; cnt = c->n2; invert = true; }
  else if (s.size() == 2 && s[0] == 'g' && s[1] >= '1' && s[1] <= '6') {
    int q = s[1] - '1';
    if (q >= c->ldim * (c->ldim + 1) / 2) { nsb_set_error("get_field: %s not defined in %dD", name, c->ldim); return 1; }
    src = c->G + (long long)q * c->n;
  } else if (s == "ifvcor") {
    if (out) out[0] = c->ifvcor[0] ? 1.0 : 0.0;
    if (count) *count = 1;
    return 0;
  } else if (s == "p2p") {          // 1: NVLink peer-memory data plane (csrc/p2p.cu), 0: NCCL / single rank
    if (out) out[0] = c->p2p.on ? 1.0 : 0.0;
    if (count) *count = 1;
    return 0;
  } else if (s == "vol") {
    if (out) { out[0] = c->vol; }
    if (count) *count = 1;
    return 0;
  }
  if (!src) { nsb_set_error("get_field: unknown field '%s'", name); return 1; }
  if (count) *count = cnt;
  if (out) {
    NSB_TRY(d2h(c, out, src, cnt));
    if (invert) for (long long i = 0; i < cnt; ++i) out[i] = 1.0 / out[i];
  }
  return 0;
}
