/* nekstab_b200.h -- C ABI of the B200-native hot path of nekStab.
 *
 * The hot path is nekStab's matrix-free Krylov loop whose matvec is one linearised Navier-Stokes
 * time-stepper integration.  In the reference this is Fortran: external procedures with implicit
 * interfaces that share state through COMMON blocks (core/NEKSTAB:6-55) and call Nek5000's
 * `nek_advance` (core/matvec.f:222).  This header is what a thin ISO_C_BINDING module binds
 * (nekstab_b200/fortran/nekstab_b200_c.f90; INTEGRATION.md shows the replaced Fortran bodies).
 *
 * Conventions
 *  - plain C types only; all arrays are HOST pointers owned by the caller unless stated otherwise;
 *    element data are Nek-ordered: a(ix,iy,iz,e), ix fastest, e = local element (Fortran column major);
 *  - every function returns 0 on success, non-zero on error (message: nsb_last_error()); the Fortran
 *    shim maps non-zero to `call nek_end`, the reference's only error convention
 *    (core/krylov_subspace.f:53, core/krylov_decomposition.f:64-67);
 *  - one context per process (= per GPU, = per MPI rank), not re-entrant, like the reference's
 *    `save`d locals (core/matvec.f:98-99);
 *  - Krylov vectors live in device memory and are addressed by integer *slots*; host copies are
 *    made only by nsb_vec_upload / nsb_vec_download (the field-access sites listed in SURVEY.md 8b).
 *  - there is NO CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef NEKSTAB_B200_H
#define NEKSTAB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Marks a pointer parameter that addresses ONE value (an output or in/out scalar), as opposed to an array: the Fortran binding
 * (nekstab_b200/fortran/nekstab_b200_c.f90, generated from this header by tools/gen_fortran_bindings.py) declares it as a scalar
 * passed by reference, every other pointer as an assumed-size array. */
#define NSB_SCALAR

/* ------------------------------------------------------------------ communicator (NCCL over NVLink)
 * Replaces Nek5000's MPI wrappers reached from the path: gop/glsc3 (core/krylov_subspace.f:37-43),
 * gslib gs_op under dssum, bcast/nekgsync (core/matvec.f:7,18).  Rank 0 creates the id, the host
 * program broadcasts the 128 bytes with whatever it has (MPI_Bcast in Nek5000), every rank calls
 * nsb_comm_init (with the CUDA device it will use) BEFORE nsb_init.  Single-GPU runs skip both calls. */
int nsb_comm_unique_id(char id_out[128]);
int nsb_comm_init(int rank, int nranks, const char id[128], int device);

/* ------------------------------------------------------------------ setup
 * nsb_init: what Nek5000 holds in SIZE/GEOM/SOLN/PARALLEL commons after nek_init, for the local
 * elements of this rank: lx1,lxd,lx2 (SIZE), xm1/ym1/zm1 (GEOM), v1mask..v3mask (SOLN), glo_num =
 * the global GLL node numbers Nek's setvert2d/3d produce for gs_setup [UPSTREAM navier8.f], nelgv.
 * zm1/v3mask may be NULL when ldim == 2.  Builds metrics, mass/stiffness factors, mesh-2 and dealiasing
 * metrics, Jacobi diagonals and the gather-scatter maps on the device. */
int nsb_init(int ldim, int lx1, int lxd, int lx2, int nelv, long long nelgv,
             const double* xm1, const double* ym1, const double* zm1,
             const double* v1mask, const double* v2mask, const double* v3mask,
             const long long* glo_num, int device);
int nsb_finalize(void);
const char* nsb_last_error(void);

/* param(2) (viscosity = 1/Re), param(1) (density), param(22)/param(21) (tolhv / tolps, absolute),
 * iteration caps (Nek: nmxh / nmxp). */
int nsb_set_params(double viscosity, double density, double tol_v, double tol_p, int maxit_v, int maxit_p);
/* bm1s: the energy weight with the sponge zeroed (core/usr_extra.f:102,116-118).  NULL -> bm1. */
int nsb_set_weights(const double* bm1s);
/* ubase, vbase, wbase (core/NEKSTAB: /nStab_bflows/; loaded at core/eigensolvers.f:180-186). */
int nsb_set_baseflow(const double* ubase, const double* vbase, const double* wbase);
/* Floquet / UPO analysis (uparam(1) = 3.11 direct, 3.21 adjoint; core/matvec.f:187-236, 277-320): with enable != 0 the base flow is
 * advanced with the FULL Navier-Stokes stepper next to the perturbation during the first nsb_matvec (Nek5000's `ifbase`), its orbit --
 * the reference's uor, vor, wor(lv, nsteps), core/krylov_subspace.f:18 -- is stored in device memory and replayed by every later
 * matvec (`ifstorebase = .true.`, core/usr_extra.f:24).  pbase (mesh 2; NULL = 0) is the pressure the base flow starts from (the P
 * field of the UPO file).  nsb_set_baseflow / nsb_set_timestep with a different nsteps discard the stored orbit.
 * nsb_get_orbit returns the stored snapshot U^{istep}, 1 <= istep <= nsteps (components may be NULL). */
int nsb_set_floquet(int enable, const double* pbase);
/* Sponge forcing of the FULL Navier-Stokes stepper (nsb_nonlinear_forward_map, the co-evolving Floquet base flow): the jp = 0 branch of
 * nekStab_forcing (core/utils.f:166-171), f += spng_str * spng_fun * (spng_vr - u) with spng_vr = the field held at nekStab_init
 * (core/utils.f:240).  ur/vr/wr NULL: the current base flow is the reference.  spng_str = 0 (the default) switches it off. */
int nsb_set_dns_sponge(double spng_str, const double* ur, const double* vr, const double* wr);
int nsb_get_orbit(int istep, double* u, double* v, double* w);
/* spng_fun (core/NEKSTAB: /nStab_sponge/): perturbation forcing -spng_fun*u' (core/utils.f:172-177).
 * NULL switches the sponge off (spng_str == 0). */
int nsb_set_sponge(const double* spng_fun);
/* prepare_linearized_solver (core/matvec.f:1-52): ctarg = compute_cfl(base flow, dt=1);
 * dt = cfl_target/ctarg; nsteps = ceiling(end_time/dt); dt = end_time/nsteps. Stores dt, nsteps. */
int nsb_prepare_linearized_solver(double end_time, double cfl_target, double* NSB_SCALAR dt, int* NSB_SCALAR nsteps, double* NSB_SCALAR ctarg);
int nsb_set_timestep(double dt, int nsteps);
/* Nek5000's `ifvcor` (no outflow-type boundary => E = D B^-1 D^T has the constant null vector => `ortho` on the
 * pressure right-hand side and solution) for the direct and the adjoint mask set: 1 / 0, or -1 to decide
 * numerically from ||E 1|| (the default after nsb_init). */
int nsb_set_ifvcor(int direct, int adjoint);
/* Pressure residual projection, Nek5000's `[PRESSURE] residualProj = yes` (on in every shipped .par, e.g. 1cyl.par:30)
 * with `mxprev` previous solutions (SIZE: mxprev=20) [UPSTREAM navier4.f setrhsp/gensolnp]: the right-hand side of the
 * pressure solve is first projected onto the E-orthonormal span of earlier solutions, which cuts the CG iteration
 * count; the converged answer is unchanged.  The basis persists across steps and matvecs, as in the reference.
 * 0 (the default) switches it off, which makes every matvec a pure function of its input (used by the parity tests). */
int nsb_set_projection(int mxprev);
/* Preconditioner of the pressure CG on E = D (mask B^-1 QQ^T) D^T.  kind 0: Jacobi (the default; the north-star's solver).
 * kind 1: three-level additive operator of the class the reference runs (`[PRESSURE] preconditioner = semg_xxt`,
 * 1cyl.par:28; [UPSTREAM] hsmg.f): element blocks by fast diagonalisation + Jacobi on the Q1 space of the element-vertex
 * mesh + an exactly solved problem on `nagg` element aggregates (total over all ranks; 0 = automatic:
 * nelv/32 per rank, at most 512 per rank and 4096 in total -- the aggregate size, hence the iteration count, then stays the same under
 * weak scaling).  Changes the
 * iteration count (measured 2 787 -> 206 on the cylinder mesh), not the converged pressure.  Rebuilt automatically when
 * nsb_set_adjoint_masks changes the adjoint operator.  Environment NSB_PRECOND=1 selects kind 1 at nsb_init. */
int nsb_set_pressure_preconditioner(int kind, int nagg);

/* ------------------------------------------------------------------ krylov_vector algebra
 * One slot = one `type(krylov_vector)` (core/krylov_subspace.f:10-15): vx,vy,vz(n), pr(n2); the
 * scalar `time` member stays on the host side.  theta (ldimt scalars) is not carried (ifheat=F). */
int nsb_vec_alloc(int nslots);                                   /* (re)allocate the slab Q(1:nslots) */
int nsb_vec_upload(int slot, const double* vx, const double* vy, const double* vz, const double* pr);
int nsb_vec_download(int slot, double* vx, double* vy, double* vz, double* pr);
int nsb_vec_copy(int dst, int src);                              /* krylov_copy  :190 */
int nsb_vec_zero(int slot);                                      /* krylov_zero  :166 */
int nsb_vec_cmult(int slot, double alpha);                       /* krylov_cmult :90  */
int nsb_vec_add2(int p, int q);                                  /* krylov_add2  :116 */
int nsb_vec_sub2(int p, int q);                                  /* krylov_sub2  :142 */
int nsb_vec_inner_product(int p, int q, double* NSB_SCALAR alpha);          /* krylov_inner_product :24 (NaN -> error) */
int nsb_vec_norm(int p, double* NSB_SCALAR alpha);                          /* krylov_norm :58 */
int nsb_vec_normalize(int p, double* NSB_SCALAR alpha);                     /* krylov_normalize :71 */
/* krylov_matmul (core/krylov_subspace.f:214-258): slot_out = sum_i y(i) * Q(first+i), i<k */
int nsb_basis_gemv(int k, int first_slot, const double* y, int slot_out);
/* basis rotation of schur_condensation (core/eigensolvers.f:466-474): Q(:,1:k) <- Q(:,1:k) * S(k,k),
 * S column-major with leading dimension lds. */
int nsb_basis_rotate(int k, int first_slot, const double* S, int lds);
/* update_hessenberg_matrix (core/krylov_decomposition.f:116-202): orthogonalise slot_f against
 * Q(first..first+k-1) twice, accumulate coefficients, normalise; hcol[0..k-1] = H(1:k,k), hcol[k] = H(k+1,k).
 * Classical Gram-Schmidt applied twice (DGKS) as two tall-skinny GEMV pairs. */
int nsb_orthonormalize(int k, int first_slot, int slot_f, double* hcol);
/* complex mode reconstruction of outpost_ks (core/eigensolvers.f:607-615): re/im slots =
 * sum_i (yre(i), yim(i)) * Q(first+i) */
int nsb_basis_gemv_complex(int k, int first_slot, const double* yre, const double* yim, int slot_re, int slot_im);

/* ------------------------------------------------------------------ direct / adjoint mode post-processing (BASELINE config 3)
 * biorthogonalize (core/sensitivity.f:428-504): direct mode (slots dre, dim) scaled to unit norm, adjoint mode (are, aim)
 * rotated / scaled so that <a, d> = 1 in the bm1s inner product; in place on the device-resident modes.
 * wave_maker (core/sensitivity.f:7-81): bi-orthonormalise, then wavemaker(x) = |u_direct(x)| |u_adjoint(x)| (n values, host). */
int nsb_biorthogonalize(int slot_dre, int slot_dim, int slot_are, int slot_aim);
int nsb_wave_maker(int slot_dre, int slot_dim, int slot_are, int slot_aim, double* wavemaker);

/* ------------------------------------------------------------------ matvec (core/matvec.f:64-159) */
enum {
  NSB_DIRECT = 1,          /* forward_linearized_map   uparam(1) in [3.0,3.2)  core/matvec.f:163 */
  NSB_ADJOINT = 2,         /* adjoint_linearized_map   uparam(1) in [3.2,3.3)  core/matvec.f:249 */
  NSB_DIRECT_ADJOINT = 3,  /* transient_growth_map     uparam(1) in [3.3,3.4)  core/matvec.f:332 */
  NSB_NEWTON = 4,          /* newton_linearized_map    floor(uparam(1)) == 2   core/matvec.f:381 (fixed points) */
  NSB_FORCE_SENS = 5       /* ts_force_sensitivity_map floor(uparam(1)) == 4   core/matvec.f:357 */
};
int nsb_matvec(int mode, int slot_in, int slot_out);
/* nonlinear_forward_map (core/newton_krylov.f:336-378): slot_f = phi_T(slot_q) - slot_q with the FULL Navier-Stokes
 * stepper (same kernels, advection term C(u)u; the Dirichlet data are the boundary values of slot_q), then ubase <- slot_q. */
/* Per-step host hook.  The reference calls the user's `nekstab_usrchk()` on the host before EVERY `nek_advance` of a matvec
 * (core/matvec.f:221,304; core/newton_krylov.f:358).  nsb_matvec keeps all `nsteps` steps on the device; a registered callback is
 * invoked on the calling host thread before step `istep` (1-based, restarted by every matvec like Nek's istep, core/matvec.f:216)
 * is enqueued, with the physical time at the start of the step.  The hook may call the nsb_set_* entry points (e.g. a time-dependent
 * base flow or sponge); it must not call nsb_matvec.  Device work of the previous step may still be in flight: use nsb_vec_download
 * to synchronise.  NULL (the default) removes the hook; the shipped cases use nekstab_usrchk only at istep == 0 (1cyl.usr:11-31). */
typedef void (*nsb_step_callback)(int istep, double time, void* user);
int nsb_set_step_callback(nsb_step_callback cb, void* user);
int nsb_nonlinear_forward_map(int slot_q, int slot_f);
/* Newton-GMRES for unstable periodic orbits (uparam(1) = 2.1).  With enable != 0:
 *  - every slot carries the `time` member of type krylov_vector (core/krylov_subspace.f:8-15; nsb_vec_set_time / nsb_vec_get_time); it
 *    follows copy / zero / cmult / add2 / sub2 / normalize / basis_gemv / basis_rotate / orthonormalize and enters
 *    krylov_inner_product as p%time * q%time (core/krylov_subspace.f:47-50);
 *  - nsb_nonlinear_forward_map stores the orbit uor, vor, wor(lv, nsteps) in device memory (core/newton_krylov.f:77-86, 364-368) and the two
 *    border vectors compute_bvec(fc_nwt), compute_bvec(ic_nwt) = one first-order Navier-Stokes step, (q1 - q0)/dt (core/matvec.f:435-475);
 *    the reference recomputes them in every matvec, here they are computed once per Newton iterate and stay resident;
 *  - nsb_matvec(NSB_NEWTON) replays the orbit as the base flow of the linearised steps (core/matvec.f:187-199, 228-231) and returns
 *    f = (exp(TL) - I) q + bvec(fc) * q%time, f%time = <bvec(ic), q> (core/matvec.f:397-419);
 *  - nsb_newton_krylov treats end_time as the first guess of the period and updates it with the Newton correction
 *    (core/newton_krylov.f:63-67, 122); read the period found with nsb_vec_get_time(q_slot). */
int nsb_set_upo(int enable);
/* Scalar transport (`ifheat`, one scalar: ldimt = 1).  After nsb_set_scalar(1, ..) a Krylov vector is [vx|vy|(vz)|theta|pr]
 * (type krylov_vector, core/krylov_subspace.f:8-15): theta enters krylov_inner_product with the weight bm1s (:41-45), follows all the
 * vector algebra, and is advanced by nsb_matvec(NSB_DIRECT / NSB_NEWTON) and nsb_nonlinear_forward_map next to the velocity
 * [UPSTREAM perturb.f heatp / cdscalp / convabp: rhocp (d/dt + U.grad) theta' + rhocp u'.grad Theta = conductivity lap theta'
 * - spng_fun theta' (nekStab_forcing_temp, core/utils.f:182-203); full equation: heat / cdscal / convab] with the velocity's BDF3/EXT3
 * scheme and Jacobi-PCG Helmholtz solver (tolerance tol_v).  conductivity = param(8), rhocp = param(7); tmask = Dirichlet mask of the
 * scalar (tmask of Nek5000); ri: the momentum equation feels f_gdir += ri * theta (ffy = temp * uparam(6) in the shipped Boussinesq
 * cases, e.g. examples/thersyphon/baseflow/tsyphon.usr), gdir 0-based.  Not built: the adjoint scalar equation (nsb_matvec(NSB_ADJOINT)
 * returns an error while the scalar is on), ldimt > 1, the scalar in Floquet / UPO orbit storage (tor).  The layout change discards
 * the slots: call nsb_vec_alloc afterwards.  nsb_set_scalar_base = tbase (core/eigensolvers.f:195-199; nsb_nonlinear_forward_map
 * sets it to slot_q's theta, core/newton_krylov.f:375). */
int nsb_set_scalar(int enable, double conductivity, double rhocp, const double* tmask, double ri, int gdir);
int nsb_set_scalar_base(const double* tbase);
int nsb_vec_upload_scalar(int slot, const double* theta);
int nsb_vec_download_scalar(int slot, double* theta);
int nsb_vec_set_time(int slot, double time);
int nsb_vec_get_time(int slot, double* NSB_SCALAR time);
/* prepare_linearized_solver evaluated on the velocity in `slot` (newton_krylov re-prepares on every iterate, :69). */
int nsb_prepare_solver_from_slot(int slot, double end_time, double cfl_target, double* NSB_SCALAR dt, int* NSB_SCALAR nsteps, double* NSB_SCALAR ctarg);
/* Adjoint problems may use different Dirichlet masks (outflow 'O' -> 'v', 1cyl.usr:126-132). NULL = same. */
int nsb_set_adjoint_masks(const double* v1mask, const double* v2mask, const double* v3mask);

/* ------------------------------------------------------------------ statistics of the last matvec / since reset */
typedef struct nsb_stats {
  long long steps;              /* linearised time steps taken */
  long long helm_iters;         /* sum over components of Helmholtz CG iterations */
  long long pres_iters;         /* pressure CG iterations */
  long long kernel_launches;    /* CUDA kernels launched by this library */
  double    step_ms;            /* device time in stepper (CUDA events) */
} nsb_stats;
int nsb_get_stats(nsb_stats* out, int reset);
/* Sampling kernel profiler (CUDA events on the launching stream around single launches, one sample set per host
 * poll of a CG loop).  enable: 1 start (clears), 0 stop (clears), -1 just read.  Arrays of 12: accumulated ms and
 * sample count per kind: 0 pressure-CG gradt, 1 dssum (ldim fields), 2 pressure-CG div, 3 pressure-CG vector update,
 * 4 Helmholtz-CG axhelm, 5 Helmholtz-CG vector update, 6 advection, 7 Helmholtz dssum, 8-10 pressure preconditioner
 * (8 restriction to the element corners, 9 vertex / aggregate levels, 10 element blocks + prolongation + z.r), 11 unused. */
int nsb_profile(int enable, double* ms_sum, long long* count);
/* measurement aid: FP64 FMA throughput of the device in TFLOP/s (register-resident DFMA loop, best of 4 after a warm-up launch) */
int nsb_fp64_peak(double* NSB_SCALAR tflops);

/* ------------------------------------------------------------------ operator-level entry points
 * Mirrors of the Nek5000 routines on the path, taking HOST arrays (copied in and out) so that every
 * kernel can be parity-tested against the oracle exactly like the reference routine would be.
 * [UPSTREAM] hmholtz.f axhelm; dssum.f dssum; math.f glsc3; navier1.f opgradt, opdiv, cdabdtp;
 * perturb.f advabp/advabp_adjoint; hmholtz.f hmholtz (Jacobi-PCG); navier1.f esolver (here Jacobi-PCG). */
int nsb_op_axhelm(const double* u, double h1, double h2, double* w);
int nsb_op_conv_scalar(const double* ax, const double* ay, const double* az, const double* phi, double* out); /* B (a.grad) phi, dealiased [UPSTREAM convect.f convop] */
int nsb_op_dssum(double* u);
int nsb_op_glsc3(const double* a, const double* b, const double* c, double* NSB_SCALAR out);
int nsb_op_opgradt(const double* p, double* wx, double* wy, double* wz);
int nsb_op_opdiv(const double* ux, const double* uy, const double* uz, double* q);
int nsb_op_cdabdtp(const double* p, double* ep);
int nsb_op_advab(int adjoint, const double* upx, const double* upy, const double* upz,
                 double* fx, double* fy, double* fz);            /* mass-weighted, un-assembled */
int nsb_op_hmholtz(double* ux, double* uy, double* uz, const double* rx, const double* ry, const double* rz,
                   double h1, double h2, int* NSB_SCALAR iters);            /* rhs un-assembled; returns du */
int nsb_op_esolver(const double* g, double* phi, int* NSB_SCALAR iters);
/* z = M^-1 r, the preconditioner selected with nsb_set_pressure_preconditioner(1, ..), on mesh-2 arrays */
int nsb_op_pc_apply(int adjoint, const double* r, double* z);
/* set-up data of that preconditioner (direct mask set) as doubles: which 0: local aggregate of every element [nelv];
 * 1: diag(P^T E P) per local vertex (ascending global corner id); 2: (Pa^T E Pa)^-1 [nagg*nagg];
 * 3: {local vertices, aggregates (all ranks), colours used for probing, local aggregates, first global aggregate id} */
int nsb_pc_get(int which, double* out, long long* NSB_SCALAR count);
int nsb_op_cfl(const double* ux, const double* uy, const double* uz, double dt, double* NSB_SCALAR cfl);
/* named geometry arrays for parity checks: "bm1","binvm1","jacm1","g1".."g6","bm2","ediag","hdiagA","vmult" */
int nsb_get_field(const char* name, double* out, long long* NSB_SCALAR count);
/* Host-only views of the gather-scatter plan (no CUDA/NCCL needed): the multi-rank map construction of gs_setup
 * (replaces gslib gs_setup's discovery of shared nodes) exposed for CPU tests with any transport.
 * 1) nsb_gs_host_candidates: this rank's element-surface global ids (ascending) -- what ranks all-gather;
 * 2) nsb_gs_host_plan: given every rank's list (counts[r], concatenated ids) build segments + halo lists;
 *    sizes_out = {nseg, len(seg_idx), nneighbours, nshared, len(rseg_pos), 0,0,0};
 * 3) nsb_gs_host_get(which): 0 seg_off 1 seg_idx 2 nbr_rank 3 nbr_off 4 send_seg 5 send_base 6 send_cnt 7 rseg_off
 *    8 rseg_pos 9 rseg_cnt 10 rseg_nbefore. */
int nsb_gs_host_candidates(int ldim, int lx1, int nelv, const long long* glo_num, long long* ids_out, long long* NSB_SCALAR count);
int nsb_gs_host_plan(int rank, int nranks, const long long* counts, const long long* ids, int sizes_out[8]);
int nsb_gs_host_get(int which, int* out);
/* Host-only pieces of the pressure-preconditioner set-up (no CUDA needed; CPU tests):
 * aggregates by recursive coordinate bisection of element centroids cent[nel][ldim] -> agg_out[nel] in [0, nagg);
 * greedy distance-2 colouring of the vertex graph from the corner ids vglo[nel][nk] (nk = 4 or 8) -> colour of every
 *   (element, corner) entry and the number of colours;
 * the 1-D FDM factors for given end weights: S[lx2*lx2] (row = node, column = mode, S^T M S = I), lam[lx2];
 * dense SPD inverse in place (row-major n x n). */
int nsb_pm_host_aggregates(int ldim, int nel, const double* cent, int nagg, int* agg_out);
int nsb_pm_host_colouring(int nel, int nk, const long long* vglo, int* colour_out, int* NSB_SCALAR ncolours);
int nsb_pm_host_fdm_1d(int lx1, double w_first, double w_last, double* S, double* lam);
int nsb_pm_host_spd_inverse(int n, double* A);
long long nsb_n(void);   /* nelv*lx1^ldim */
long long nsb_n2(void);  /* nelv*lx2^ldim */

/* ------------------------------------------------------------------ host-side Krylov drivers (C++),
 * mirroring the reference's Fortran drivers one to one; small dense work on host LAPACK
 * (dgeev/dgees/dtrsen/dgels, core/lapack_wrapper.f:7-339).  H is column-major (ldh >= ksize+1). */
int nsb_arnoldi_factorization(int mode, int first_slot, double* H, int ldh, int mstart, int mend, int ksize); /* core/krylov_decomposition.f:7 */
/* krylov_schur (core/eigensolvers.f:141-388): returns Ritz values (sorted by decreasing magnitude),
 * residuals, eigenvectors of H (column-major complex interleaved, k x k) and the count of converged. */
int nsb_krylov_schur(int mode, int k_dim, int schur_tgt, double eigen_tol, double schur_del, int seed_slot,
                     double* vals_re, double* vals_im, double* residual, double* vecs_reim, int* NSB_SCALAR n_converged,
                     int* NSB_SCALAR schur_cnt, int max_restarts);
int nsb_schur_condensation(int* NSB_SCALAR mstart, double* H, int ldh, int first_slot, int ksize, int schur_tgt, double schur_del); /* :395 */
int nsb_select_eigenvalues(int* selected, int* NSB_SCALAR cnt, const double* vals_re, const double* vals_im, double delta, int nev, int n); /* :729 */
/* ts_gmres (core/newton_krylov.f:175-297): solves matvec(mode) * sol = rhs; slots first..first+ksize hold the basis */
int nsb_ts_gmres(int mode, int rhs_slot, int sol_slot, int first_slot, int work_slot, int maxiter, int ksize, double tol,
                 int* NSB_SCALAR calls, double* NSB_SCALAR final_res);
/* newton_krylov (core/newton_krylov.f:5-168), fixed-point branch (uparam(1) = 2): Newton iterations on phi_T(q) - q = 0
 * with ts_gmres on newton_linearized_map; squared residual norms against tol as in the reference.  Returns 0 when
 * converged, 3 when maxiter_newton was reached. */
int nsb_newton_krylov(int q_slot, int f_slot, int dq_slot, int work_slot, int first_slot, int k_dim, double end_time,
                      double cfl_target, double tol, int maxiter_newton, int maxiter_gmres, int* NSB_SCALAR newton_iters,
                      double* NSB_SCALAR residual_out, double* hist, long long* NSB_SCALAR calls_out);
/* LAPACK wrappers exactly as core/lapack_wrapper.f (schur:7, ordschur:70, eig:129, lstsq:287). */
int nsb_lapack_eig(const double* A, int n, double* vals_re, double* vals_im, double* vecs_reim);
int nsb_lapack_schur(double* A, int n, double* vecs, double* vals_re, double* vals_im);
int nsb_lapack_ordschur(double* T, double* Q, const int* selected, int n);
int nsb_lapack_lstsq(const double* A, const double* b, double* x, int m, int n);
int nsb_lapack_load(const char* path);   /* optional: library exporting LP64 dgeev_/dgees_/dtrsen_/dgels_ (or scipy_ prefixed) */

#ifdef __cplusplus
}
#endif
#endif
