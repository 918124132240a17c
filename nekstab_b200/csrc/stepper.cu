// stepper.cu -- the linearised Navier-Stokes time stepper driven entirely from device-resident state.
//
// Restates one `nek_advance` in perturbation mode (P_N-P_{N-2}, BDF3/EXT3, order ramp 1,2,3 restarted by every
// matvec: core/matvec.f:216) [UPSTREAM drive1.f, perturb.f fluidp/perturbv/makefp/cresvipp/incomprp] and the maps
// nekStab builds on it: forward_linearized_map (core/matvec.f:163-241), adjoint_linearized_map (:249-325).
// Step n -> n+1:
//   f   = -B sigma u' - [C(u')U + C(U)u']                    (adjoint: -[(grad U)^T u' - C(U)u'])     k_advab
//   b   = sum_j ab_j f^{n+1-j} + (rho/dt) B sum_j bd_{j+1} u^{n+1-j}                                  k_make_rhs
//   pt  = p^n (order<3) | 2p^n - p^{n-1}                                                                k_press_extrap
//   r   = b + D^T pt - H u^n ; r <- mask QQ^T r                                                         k_gradt, k_axhelm<1>, dssum
//   H du = r  (Jacobi-PCG, 3 components batched) ; uh = u^n + du                                        st_helmholtz
//   E phi = -D uh (Jacobi-PCG) ; u^{n+1} = uh + mask B~^-1 QQ^T D^T phi ; p^{n+1} = pt + (bd1/dt) phi   st_pressure
// The CG loops run without host round trips: scalars, convergence flags and iteration counters live in
// device memory (CGState); the host only polls the flag every `check_every` iterations, and kernels of
// iterations issued past convergence exit immediately, so results do not depend on the polling interval.
#include <cmath>
#include <utility>

#include "nsb_internal.h"

static const double BD[4][4] = {{0, 0, 0, 0}, {1.0, 1.0, 0, 0}, {1.5, 2.0, -0.5, 0}, {11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0}};
static const double AB[4][3] = {{0, 0, 0}, {1.0, 0, 0}, {2.0, -1.0, 0}, {3.0, -3.0, 1.0}};

int vk_cg_finalize_multi(Ctx* c, CGState* s, int ncomp, int kind);

template <class T>
static int dalloc(T** p, long long count) {
  NSB_CUDA(cudaMalloc((void**)p, std::max<long long>(count, 1) * sizeof(T)));
  NSB_CUDA(cudaMemset(*p, 0, std::max<long long>(count, 1) * sizeof(T)));
  return 0;
}

int st_alloc(Ctx* c) {
  const long long dn = c->n * c->ldim;
  NSB_TRY(dalloc(&c->u, dn));
  NSB_TRY(dalloc(&c->ulag[0], dn));
  NSB_TRY(dalloc(&c->ulag[1], dn));
  for (int j = 0; j < 3; ++j) NSB_TRY(dalloc(&c->f[j], dn));
  NSB_TRY(dalloc(&c->pr, c->n2));
  NSB_TRY(dalloc(&c->prlag, c->n2));
  NSB_TRY(dalloc(&c->pt, c->n2));
  for (int j = 0; j < 4; ++j) NSB_TRY(dalloc(&c->wk[j], dn));
  NSB_TRY(dalloc(&c->rk, dn));
  for (int j = 0; j < 5; ++j) NSB_TRY(dalloc(&c->pk[j], c->n2));
  NSB_TRY(dalloc(&c->cgs, 4));
  NSB_CUDA(cudaMallocHost((void**)&c->cgs_host, 4 * sizeof(CGState)));
  memset(c->cgs_host, 0, 4 * sizeof(CGState));
  return 0;
}

static int cg_state_setup(Ctx* c, int first, int count, double tol, double vol, int maxit) {
  for (int f = first; f < first + count; ++f) {
    CGState& s = c->cgs_host[f];
    memset(&s, 0, sizeof(s));
    s.tol = tol; s.vol = vol; s.maxit = maxit; s.rtz2 = 1.0;
  }
  NSB_CUDA(cudaMemcpyAsync(c->cgs + first, c->cgs_host + first, count * sizeof(CGState), cudaMemcpyHostToDevice, c->stream));
  return 0;
}

static int cg_state_poll(Ctx* c, int first, int count, bool* all_done) {
  NSB_CUDA(cudaMemcpyAsync(c->cgs_host + first, c->cgs + first, count * sizeof(CGState), cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  *all_done = true;
  for (int f = first; f < first + count; ++f)
    if (!c->cgs_host[f].done) *all_done = false;
  return 0;
}

// ---- sampling profiler: kinds 0 pcg_gradt, 1 dssum(3 fields), 2 pcg_div, 3 pcg_update, 4 hcg_axhelm, 5 hcg_update,
//      6 advab, 7 helmholtz dssum, 8-10 preconditioner pieces, 11-12 Gram-Schmidt, 13 fused pressure-CG tail.  Events bracket ONE
//      launch each; elapsed times are read after the next host poll.
static inline void prof_mark(Ctx* c, bool on, int slot) {
  if (on) cudaEventRecord(c->prof_ev[slot], c->stream);
}
static void prof_collect(Ctx* c, const int* kinds, int nk, int first_slot) {
  for (int i = 0; i < nk; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->prof_ev[first_slot + i], c->prof_ev[first_slot + i + 1]) == cudaSuccess) {
      c->prof_ms[kinds[i]] += ms;
      c->prof_cnt[kinds[i]] += 1;
    }
  }
}

// ---- CUDA-graph replay of one CG batch (profiler off; multi-rank when the data plane is peer memory).  The first batches of a context run un-captured so
//      that one-time function attributes are set outside any capture.
static bool graphs_ok(Ctx* c) {
  // multi-rank: every collective inside a batch must be a peer-memory kernel (device-resident epochs, csrc/p2p.cu); the NCCL fallback
  // is enqueued from the host and is not captured
  const bool all_p2p = c->p2p.on && (!c->gsv_ready || c->p2pv.on) && (!c->gsp_ready || c->p2pp.on);
  return c->use_graphs && (c->nranks == 1 || all_p2p) && !c->prof_on && c->graph_warm;
}
template <class F>
static int run_batch_graph(Ctx* c, Ctx::GraphEntry* ge, F&& batch) {
  if (!ge->exec) {
    const long long l0 = c->stats.kernel_launches;
    cudaGraph_t g = nullptr;
    NSB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = batch(false);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc || e != cudaSuccess || !g) {
      nsb_set_error("CUDA graph capture of a CG batch failed (%s)", cudaGetErrorString(e));
      return rc ? rc : 1;
    }
    NSB_CUDA(cudaGraphInstantiate(&ge->exec, g, 0));
    cudaGraphDestroy(g);
    ge->launches = (int)(c->stats.kernel_launches - l0);
    c->stats.kernel_launches = l0;                 // captured, not executed
  }
  NSB_CUDA(cudaGraphLaunch(ge->exec, c->stream));
  c->stats.kernel_launches += ge->launches;
  return 0;
}

// Solve (h1 A + h2 B) x_c = r_c for the ldim velocity components at once (independent CG recurrences sharing
// every kernel launch) [UPSTREAM hmholtz.f hmholtz/cggo].  In: c->rk (assembled, masked). Out: c->wk[3].
// ncomp = 0: the ldim velocity components; ncomp = 1 solves a single field held in component 0 of rk / wk (the scalar, whose mask the
// caller has swapped into mask[adj][0]); graph_key keeps its captured batches apart from the velocity's.
int st_helmholtz(Ctx* c, int adj, double h1, double h2, int* iters, int ncomp, int graph_key) {
  const int nc = ncomp > 0 ? ncomp : c->ldim;
  const int gkey = graph_key >= 0 ? graph_key : adj;
  NSB_TRY(vk_dinvH(c, h1, h2));
  NSB_TRY(cg_state_setup(c, 0, nc, c->tol_v, c->vol, c->maxit_v));
  NSB_TRY(vk_hcg_init(c, nc));
  bool done = false;
  int issued = 0;
  Ctx::GraphEntry* ge = nullptr;
  if (graphs_ok(c)) {
    for (auto& g : c->graph_h)
      if (g.exec && g.adj == gkey && g.h1 == h1 && g.h2 == h2) ge = &g;
    if (!ge)
      for (auto& g : c->graph_h)
        if (!g.exec) { ge = &g; ge->adj = gkey; ge->h1 = h1; ge->h2 = h2; break; }
  }
  auto batch = [&](bool sample) -> int {
    for (int it = 0; it < c->check_every_v; ++it) {
      const bool sm = sample && it == 0;
      prof_mark(c, sm, 0);
      NSB_TRY(ek_hcg_dir_ax(c, nc, h1, h2));
      if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, c->cgs, nc, 2));
      prof_mark(c, sm, 1);
      NSB_TRY(gs_dssum_w(c, c->wk[2], nc, nc == 3 && perm_h_active(c), nullptr));
      prof_mark(c, sm, 2);
      NSB_TRY(vk_hcg_update(c, nc, adj));
      prof_mark(c, sm, 3);
    }
    return 0;
  };
  while (!done) {
    const bool sample = c->prof_on && issued == 0;
    if (ge) NSB_TRY(run_batch_graph(c, ge, batch));
    else NSB_TRY(batch(sample));
    issued += c->check_every_v;
    NSB_TRY(cg_state_poll(c, 0, nc, &done));
    if (sample) { const int kinds[3] = {4, 7, 5}; prof_collect(c, kinds, 3, 0); }
    if (issued > c->maxit_v + c->check_every_v) break;
  }
  int tot = 0;
  for (int f = 0; f < nc; ++f) {
    tot += c->cgs_host[f].iter;
    if (!(c->cgs_host[f].rnorm == c->cgs_host[f].rnorm)) { nsb_set_error("Helmholtz CG produced NaN (component %d)", f); return 2; }
  }
  NSB_TRY(p2p_check_error(c));
  c->stats.helm_iters += tot;
  if (iters) *iters = tot;
  return 0;
}

// out = E in  (in, out: mesh-2 vectors; uses wk[2] as scratch)
static int apply_E(Ctx* c, int adj, const double* in, double* out) {
  NSB_TRY(ek_gradt(c, in, c->wk[2]));
  NSB_TRY(gs_dssum(c, c->wk[2], c->ldim, c->n, nullptr));
  return ek_div_mbinv(c, c->wk[2], adj, out, 1.0);
}

// Residual projection [UPSTREAM navier4.f setrhsp]: xbar = X X^T g (X is E-orthonormal), g <- g - (E X) X^T g.
// xbar is left in pk[4].
static int proj_pre(Ctx* c, int adj) {
  if (c->proj_adj != adj) { c->proj_m = 0; c->proj_adj = adj; }      // the basis belongs to one operator E
  const int m = c->proj_m;
  if (m == 0) return vk_fill(c, c->pk[4], 0.0, c->n2);
  double* al = c->hbuf + 40000;
  NSB_TRY(vk_multidot_raw(c, m, c->projX, c->n2, c->pk[0], nullptr, c->n2, c->n2, al));
  NSB_TRY(vk_multiaxpy_raw(c, m, c->projX, c->n2, c->pk[4], 0.0, al, 1.0, c->pk[4]));
  return vk_multiaxpy_raw(c, m, c->projEX, c->n2, c->pk[0], 1.0, al, -1.0, c->pk[0]);
}
// [UPSTREAM navier4.f gensolnp]: phi = xbar + delta; E-orthonormalise delta against the basis and append it (restart
// from the full solution when the basis is full).  delta = pk[1] on entry; phi in pk[1] on exit.
static int proj_post(Ctx* c, int adj) {
  double* al = c->hbuf + 40000;
  double* x = c->pk[1];
  int m = c->proj_m;
  double* xn;
  double* exn;
  if (m == c->proj_max) {                      // restart: keep only the newest full solution
    NSB_TRY(vk_axpy(c, x, 1.0, c->pk[4], c->n2));
    xn = c->projX; exn = c->projEX;
    NSB_TRY(vk_copy(c, xn, x, c->n2));
    NSB_TRY(apply_E(c, adj, xn, exn));
    m = 0;
  } else {
    xn = c->projX + (long long)m * c->n2; exn = c->projEX + (long long)m * c->n2;
    NSB_TRY(vk_copy(c, xn, x, c->n2));
    NSB_TRY(apply_E(c, adj, xn, exn));
    if (m > 0) {                               // xn -= X (X^T E xn) ; E xn likewise
      NSB_TRY(vk_multidot_raw(c, m, c->projX, c->n2, exn, nullptr, c->n2, c->n2, al));
      NSB_TRY(vk_multiaxpy_raw(c, m, c->projX, c->n2, xn, 1.0, al, -1.0, xn));
      NSB_TRY(vk_multiaxpy_raw(c, m, c->projEX, c->n2, exn, 1.0, al, -1.0, exn));
    }
    NSB_TRY(vk_axpy(c, x, 1.0, c->pk[4], c->n2));
  }
  // normalise in the E inner product
  NSB_TRY(vk_dot3(c, xn, exn, nullptr, c->n2, c->red_out + 10));
  NSB_TRY(vk_allreduce_sum(c, c->red_out + 10, 1));
  double nrm2 = 0;
  NSB_CUDA(cudaMemcpyAsync(&nrm2, c->red_out + 10, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  if (nrm2 > 0 && nrm2 == nrm2) {
    const double sc = 1.0 / std::sqrt(nrm2);
    NSB_TRY(vk_scale(c, xn, sc, c->n2));
    NSB_TRY(vk_scale(c, exn, sc, c->n2));
    c->proj_m = m + 1;
  } else {
    c->proj_m = m;                             // zero update (already converged): nothing to add
  }
  return 0;
}

// Solve E x = g with PCG, E = D (mask B~^-1 QQ^T) D^T [UPSTREAM navier1.f esolver/cdabdtp; the reference runs
// GMRES+multigrid here, the north-star prescribes Jacobi-PCG].  Preconditioner: Jacobi (pc_kind 0) or the three-level
// additive Schwarz/multilevel operator of pmg.cu (pc_kind 1).  In: c->pk[0] = g (destroyed). Out: c->pk[1].
int st_pressure(Ctx* c, int adj, int* iters) {
  CGState* sp = c->cgs + 3;
  if (c->ifvcor[adj]) {   // remove the constant null-space component from the right-hand side [UPSTREAM ortho]
    NSB_TRY(vk_sum(c, c->pk[0], c->n2, c->red_out + 8));
    NSB_TRY(vk_allreduce_sum(c, c->red_out + 8, 1));
    NSB_TRY(vk_add_scalar_from_dev(c, c->pk[0], c->red_out + 8, -1.0 / (double)c->n2_glob, c->n2));
  }
  const bool proj = c->proj_max > 0;
  if (proj) NSB_TRY(proj_pre(c, adj));
  NSB_TRY(cg_state_setup(c, 3, 1, c->tol_p, c->vol2, c->maxit_p));
  const bool pc = c->pc_kind != 0;
  // 3-D + three-level preconditioner: everything between the E application and the next direction is one fused tail
  // (pm_pcg_tail, csrc/pmg.cu); otherwise the separate update / preconditioner kernels
  const bool fused = c->pc_kind == 1 && c->pcg_fused && c->ldim == 3;
  if (fused) {
    NSB_TRY(pm_pcg_tail(c, adj, 1, 0));
  } else {
    NSB_TRY(vk_pcg_init(c, adj));
    if (pc) NSB_TRY(pm_apply(c, adj, c->pk[0], c->pz, 1));
  }
  bool done = false;
  int issued = 0;
  Ctx::GraphEntry* ge = graphs_ok(c) ? &c->graph_p[adj] : nullptr;
  auto batch = [&](bool sample) -> int {
    for (int it = 0; it < c->check_every_p; ++it) {
      const bool sm = sample && it == 0;       // first iteration of the batch: cannot be a skipped (converged) one
      prof_mark(c, sm, 4);
      NSB_TRY(ek_pcg_dir_gradt(c, adj));
      prof_mark(c, sm, 5);
      NSB_TRY(gs_dssum(c, c->wk[2], c->ldim, c->n, sp));
      prof_mark(c, sm, 6);
      NSB_TRY(ek_pcg_div(c, adj));
      if (c->nranks > 1) NSB_TRY(vk_cg_finalize_multi(c, sp, 1, 2));
      prof_mark(c, sm, 7);
      if (fused) {
        NSB_TRY(pm_pcg_tail(c, adj, 0, sm ? 8 : 0));        // records events 8 (after k_pcg_fused) and 9 (after the coarse levels)
      } else {
        NSB_TRY(vk_pcg_update(c, adj));
        prof_mark(c, sm, 8);
        if (pc) NSB_TRY(pm_apply(c, adj, c->pk[0], c->pz, 1, sm ? 9 : 0));
        prof_mark(c, sm, 11);
      }
    }
    return 0;
  };
  while (!done) {
    const bool sample = c->prof_on != 0;
    if (ge) NSB_TRY(run_batch_graph(c, ge, batch));
    else NSB_TRY(batch(sample));
    issued += c->check_every_p;
    NSB_TRY(cg_state_poll(c, 3, 1, &done));
    if (sample) {
      if (fused) { const int kinds[5] = {0, 1, 2, 13, 9}; prof_collect(c, kinds, 5, 4); }
      else { const int kinds[7] = {0, 1, 2, 3, 8, 9, 10}; prof_collect(c, kinds, pc ? 7 : 4, 4); }
    }
    if (issued > c->maxit_p + c->check_every_p) break;
  }
  if (fused && c->cgs_host[3].iter > 0) NSB_TRY(pm_pcg_xfix(c));       // the last x += alpha p (see k_gradt3 MODE 2)
  NSB_TRY(p2p_check_error(c));
  if (!(c->cgs_host[3].rnorm == c->cgs_host[3].rnorm)) { nsb_set_error("pressure CG produced NaN"); return 2; }
  if (proj) NSB_TRY(proj_post(c, adj));
  if (c->ifvcor[adj]) {
    NSB_TRY(vk_sum(c, c->pk[1], c->n2, c->red_out + 8));
    NSB_TRY(vk_allreduce_sum(c, c->red_out + 8, 1));
    NSB_TRY(vk_add_scalar_from_dev(c, c->pk[1], c->red_out + 8, -1.0 / (double)c->n2_glob, c->n2));
  }
  c->stats.pres_iters += c->cgs_host[3].iter;
  if (iters) *iters = c->cgs_host[3].iter;
  return 0;
}

// ---- scalar transport (ifheat): theta' advanced next to the velocity [UPSTREAM perturb.f heatp / cdscalp / makeqp / convabp; full
//      equation: heat / cdscal / makeq / convab].  Explicit term of step n (all fields at level n):
//        q = -rhocp * B [ (U.grad) theta' + (u'.grad) Theta ] - B spng_fun theta'      perturbation (nekStab_forcing_temp jp = 1, core/utils.f:199)
//        q = -rhocp * B (u.grad) theta                                                 full equation
//      and the momentum feels  f_g += B ri theta  (userf of the shipped Boussinesq cases: ffy = temp * uparam(6)).
static int scalar_explicit(Ctx* c, int kind, double* fvel) {
  Ctx::Scalar& z = c->scal;
  double* qnew = z.q[2];
  if (kind == 2) NSB_TRY(sk_conv(c, c->u, z.th, nullptr, nullptr, qnew));
  else NSB_TRY(sk_conv(c, c->ub, z.th, c->u, z.tb, qnew));
  NSB_TRY(vk_scale(c, qnew, -z.rhocp, c->n));
  NSB_TRY(vk_copy(c, z.wk[0], z.th, c->n));
  NSB_TRY(vk_mul(c, z.wk[0], c->bm1, c->n));                         // B theta
  if (z.ri != 0.0) NSB_TRY(vk_axpy(c, fvel + (long long)z.gdir * c->n, z.ri, z.wk[0], c->n));
  if (kind != 2 && c->spng) {
    NSB_TRY(vk_mul(c, z.wk[0], c->spng, c->n));
    NSB_TRY(vk_axpy(c, qnew, -1.0, z.wk[0], c->n));
  }
  z.q[2] = z.q[1]; z.q[1] = z.q[0]; z.q[0] = qnew;
  return 0;
}
// implicit part: (cond A + rhocp bd0/dt B) dtheta = b - H theta^n, theta^{n+1} = theta^n + dtheta (cdscalp), Jacobi-PCG of st_helmholtz
static int scalar_solve(Ctx* c, int k) {
  Ctx::Scalar& z = c->scal;
  const double h1 = z.cond, h2 = z.rhocp * BD[k][0] / c->dt;
  double* b = z.wk[1];
  NSB_TRY(vk_fill(c, z.wk[0], 0.0, c->n));
  NSB_TRY(vk_fill(c, b, 0.0, c->n));
  const double* lag[3] = {z.th, z.thlag[0], z.thlag[1]};
  for (int j = 0; j < k; ++j) {
    NSB_TRY(vk_axpy(c, b, AB[k][j], z.q[j], c->n));
    NSB_TRY(vk_axpy(c, z.wk[0], BD[k][j + 1], lag[j], c->n));
  }
  NSB_TRY(vk_mul(c, z.wk[0], c->bm1, c->n));
  NSB_TRY(vk_axpy(c, b, z.rhocp / c->dt, z.wk[0], c->n));
  NSB_TRY(vk_fill(c, c->rk, 0.0, c->n));
  NSB_TRY(ek_axhelm_resid(c, z.th, b, c->rk, 1, h1, h2));             // r = b + r - H theta^n
  NSB_TRY(gs_dssum(c, c->rk, 1, c->n, nullptr));
  NSB_TRY(vk_mul(c, c->rk, z.tmask, c->n));
  std::swap(c->mask[0][0], z.tmask);                                  // the CG update masks component 0 with mask[adj][0]
  int rc = st_helmholtz(c, 0, h1, h2, nullptr, 1, 2);
  std::swap(c->mask[0][0], z.tmask);
  if (rc) return rc;
  double* tn = z.thlag[1];
  NSB_TRY(vk_lin2(c, tn, 1.0, z.th, 1.0, c->wk[3], c->n));
  z.thlag[1] = z.thlag[0]; z.thlag[0] = z.th; z.th = tn;
  return 0;
}

// kind: 0 direct, 1 adjoint perturbation step; 2 full Navier-Stokes step (nonlinear_forward_map, core/newton_krylov.f:336-378)
static int one_step(Ctx* c, int istep, int kind) {
  const int adj = (kind == 1) ? 1 : 0;
  const int D = c->ldim;
  const long long dn = c->n * D;
  const int k = istep < 3 ? istep : 3;
  const double h1 = c->visc, h2 = c->rho * BD[k][0] / c->dt;
  // explicit term into the oldest ring slot, then rotate so that f[0] is current
  double* fnew = c->f[2];
  prof_mark(c, c->prof_on, 10);
  if (kind == 2) {
    // f = -B (u.grad)u : the direct perturbation form with U = u' = u gives 2 C(u)u  [UPSTREAM navier1.f makef -> advab -> convop]
    NSB_TRY(ek_advab(c, 0, c->u, c->u, nullptr, fnew));
    NSB_TRY(vk_scale(c, fnew, 0.5, dn));
    NSB_TRY(vk_sponge_dns(c, fnew, c->u));        // + B spng_str spng_fun (spng_vr - u) when the DNS sponge is set (core/utils.f:166-171)
  } else {
    NSB_TRY(ek_advab(c, adj, c->u, c->ub, c->spng, fnew));
  }
  prof_mark(c, c->prof_on, 11);
  if (c->scal.on) NSB_TRY(scalar_explicit(c, kind, fnew));
  c->f[2] = c->f[1]; c->f[1] = c->f[0]; c->f[0] = fnew;
  double* b = c->wk[0];
  NSB_TRY(vk_make_rhs(c, b, k, AB[k], BD[k]));
  NSB_TRY(vk_press_extrap(c, k));
  double* r = c->rk;
  NSB_TRY(ek_gradt(c, c->pt, r));
  NSB_TRY(ek_axhelm_resid(c, c->u, b, r, D, h1, h2));
  NSB_TRY(gs_dssum(c, r, D, c->n, nullptr));
  NSB_TRY(vk_mask_fields(c, r, adj));
  NSB_TRY(st_helmholtz(c, adj, h1, h2, nullptr));
  if (c->prof_on) { const int kinds[1] = {6}; prof_collect(c, kinds, 1, 10); }   // Helmholtz polls => advab events are complete
  // uh = u + du into the oldest velocity buffer (ulag[1]); it becomes the new current field below
  double* un = c->ulag[1];
  NSB_TRY(vk_lin2(c, un, 1.0, c->u, 1.0, c->wk[3], dn));
  NSB_TRY(ek_div(c, un, nullptr, c->pk[0], -1.0));
  NSB_TRY(st_pressure(c, adj, nullptr));
  NSB_TRY(ek_gradt(c, c->pk[1], c->wk[2]));
  NSB_TRY(gs_dssum(c, c->wk[2], D, c->n, nullptr));
  NSB_TRY(vk_final_update(c, adj, h2));
  if (c->scal.on) NSB_TRY(scalar_solve(c, k));
  // rotate: velocity (u -> ulag0 -> ulag1), pressure (pr <-> prlag)
  double* t = c->ulag[1]; c->ulag[1] = c->ulag[0]; c->ulag[0] = c->u; c->u = t;
  double* tp = c->prlag; c->prlag = c->pr; c->pr = tp;
  c->stats.steps += 1;
  c->graph_warm = true;
  return 0;
}

// ---- Floquet / UPO support: the base flow co-evolves with the full Navier-Stokes stepper (Nek's `ifbase`), the orbit is stored in HBM on
//      the first matvec and replayed afterwards (`ifstorebase`): forward_linearized_map core/matvec.f:187-236, adjoint :277-320.
static void swap_state(Ctx* c, Ctx::StepState& s) {
  std::swap(c->u, s.u); std::swap(c->ulag[0], s.ulag[0]); std::swap(c->ulag[1], s.ulag[1]);
  std::swap(c->f[0], s.f[0]); std::swap(c->f[1], s.f[1]); std::swap(c->f[2], s.f[2]);
  std::swap(c->pr, s.pr); std::swap(c->prlag, s.prlag);
}
static int orbit_alloc(Ctx* c) {
  const long long dn = c->n * c->ldim;
  if (c->orbit_steps != c->nsteps || !c->orbit) {
    if (c->orbit) cudaFree(c->orbit);
    c->orbit = nullptr;
    NSB_TRY(dalloc(&c->orbit, (long long)c->nsteps * dn));      // "ALLOCATING ORBIT WITH NSTEPS" core/matvec.f:201-209, core/newton_krylov.f:77-86
    c->orbit_steps = c->nsteps;
    c->orbit_ready = false;
  }
  return 0;
}
static int floquet_prepare(Ctx* c) {
  const long long dn = c->n * c->ldim;
  Ctx::StepState& b = c->base_state;
  if (!b.u) {
    NSB_TRY(dalloc(&b.u, dn)); NSB_TRY(dalloc(&b.ulag[0], dn)); NSB_TRY(dalloc(&b.ulag[1], dn));
    for (int j = 0; j < 3; ++j) NSB_TRY(dalloc(&b.f[j], dn));
    NSB_TRY(dalloc(&b.pr, c->n2)); NSB_TRY(dalloc(&b.prlag, c->n2));
  }
  NSB_TRY(orbit_alloc(c));
  if (!c->orbit_ready) {                                       // the base flow starts from the given field (opcopy(vx.. <- ubase), core/matvec.f:103)
    NSB_TRY(vk_copy(c, b.u, c->ub0, dn));
    if (c->pb0) NSB_TRY(vk_copy(c, b.pr, c->pb0, c->n2));
    else NSB_TRY(vk_fill(c, b.pr, 0.0, c->n2));
  }
  return 0;
}

// vin / vout: device Krylov vectors [vx|vy|(vz)|pr]
int st_linearized_map(Ctx* c, int adjoint, const double* vin, double* vout) {
  if (c->nsteps <= 0 || c->dt <= 0) { nsb_set_error("time step not set: call nsb_prepare_linearized_solver / nsb_set_timestep"); return 1; }
  if (!c->ub && adjoint != 2) { nsb_set_error("base flow not set: call nsb_set_baseflow"); return 1; }
  const long long dn = c->n * c->ldim;
  const bool flq = c->floquet && adjoint != 2;
  // UPO Newton (uparam(1) = 2.1, ifstorebase): the nonlinear map stores the orbit (core/newton_krylov.f:364-368), the linearised maps replay it
  const bool upo_store = c->upo && adjoint == 2, upo_replay = c->upo && adjoint != 2 && !flq;
  if (flq) NSB_TRY(floquet_prepare(c));
  if (upo_store) { NSB_TRY(orbit_alloc(c)); c->orbit_ready = false; }
  if (upo_replay && (!c->orbit_ready || c->orbit_steps != c->nsteps)) {
    nsb_set_error("UPO: no stored orbit for %d steps (call nsb_nonlinear_forward_map with the current time step first)", c->nsteps);
    return 1;
  }
  NSB_CUDA(cudaEventRecord(c->ev0, c->stream));
  if (c->scal.on) {
    if (adjoint == 1) { nsb_set_error("scalar transport: the adjoint equations are not built (direct and full Navier-Stokes maps only)"); return 1; }
    if (flq || c->upo) { nsb_set_error("scalar transport: Floquet / UPO orbit storage does not carry the scalar yet (tor)"); return 1; }
    if (adjoint != 2 && !c->scal.tb) { nsb_set_error("scalar transport: base scalar field not set (nsb_set_scalar_base)"); return 1; }
    NSB_TRY(vk_copy(c, c->scal.th, vin + dn, c->n));
  }
  NSB_TRY(vk_copy(c, c->u, vin, dn));            // nopcopy(vxp,..,prp <- q)   core/matvec.f:212
  NSB_TRY(vk_copy(c, c->pr, vin + c->poff, c->n2));
  int rc = 0;
  for (int istep = 1; istep <= c->nsteps && !rc; ++istep) {
    if (c->step_cb) c->step_cb(istep, (istep - 1) * c->dt, c->step_cb_user);     // nekstab_usrchk(), core/matvec.f:221,304
    if (flq) {
      if (!c->orbit_ready) {                     // first matvec: advance the base flow (full NS step) and store U^{istep} (:224-227)
        swap_state(c, c->base_state);
        rc = one_step(c, istep, 2);
        swap_state(c, c->base_state);
        if (rc) break;
        NSB_TRY(vk_copy(c, c->orbit + (long long)(istep - 1) * dn, c->base_state.u, dn));
      }
      // the perturbation's explicit terms of step istep see U^{istep-1}: the given field at step 1, then the stored orbit (:228-231)
      c->ub = (istep == 1) ? c->ub0 : c->orbit + (long long)(istep - 2) * dn;
    }
    if (upo_replay) c->ub = (istep == 1) ? c->ub0 : c->orbit + (long long)(istep - 2) * dn;     // "using stored baseflow" (:228-231)
    rc = one_step(c, istep, adjoint);
    if (upo_store && !rc) NSB_TRY(vk_copy(c, c->orbit + (long long)(istep - 1) * dn, c->u, dn));
  }
  if (flq) {
    c->ub = c->ub0;
    if (!rc) c->orbit_ready = true;              // ifbase = .false.; init = .true.  (:234-236)
  }
  if (upo_replay) c->ub = c->ub0;
  if (upo_store && !rc) c->orbit_ready = true;
  if (rc) return rc;
  NSB_TRY(vk_copy(c, vout, c->u, dn));           // nopcopy(f <- vxp,..,prp)   core/matvec.f:239
  if (c->scal.on) NSB_TRY(vk_copy(c, vout + dn, c->scal.th, c->n));
  NSB_TRY(vk_copy(c, vout + c->poff, c->pr, c->n2));
  NSB_CUDA(cudaEventRecord(c->ev1, c->stream));
  NSB_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  NSB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stats.step_ms += ms;
  return 0;
}
