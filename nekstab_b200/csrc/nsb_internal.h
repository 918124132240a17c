// nsb_internal.h -- process-wide context and launcher declarations of the nekstab_b200 CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nekstab_b200.h"

#define NSB_MAXD 3
#define NSB_MAX_RED 16          // scalars per reduction
#define NSB_MAX_BLOCKS 4096     // upper bound on reduction grid sizes

void nsb_set_error(const char* fmt, ...);

#define NSB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      nsb_set_error("%s:%d CUDA error: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(e_), #call); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)
#define NSB_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r_ = (call);                                                            \
    if (r_ != ncclSuccess) {                                                             \
      nsb_set_error("%s:%d NCCL error: %s (%s)", __FILE__, __LINE__, ncclGetErrorString(r_), #call); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)
#define NSB_TRY(call)          \
  do {                         \
    int r__ = (call);          \
    if (r__) return r__;       \
  } while (0)

// Small dense SEM matrices, packed row-major with their true sizes (template sizes in kernels).
struct ConstMats {
  double D[144], Dt[144];                               // lx1 x lx1 derivative and transpose
  double J12[120], J12t[120], D12[120], D12t[120];      // lx2 x lx1 (t: lx1 x lx2)
  double Jd[216], Jdt[216], Dd[216], Ddt[216];          // lxd x lx1 (t: lx1 x lxd)
  double w1[12], w2[12], wd[18];
  double z1[12];
  double hat1[12];                                       // Q1 hat function (1+z)/2 at the GL(lx2) points (pressure preconditioner)
};

// Device-resident state of one conjugate-gradient recurrence (one per solved component).
struct CGState {
  double rtz1, rtz2, rho, alpha, beta, rnorm;
  double tol, vol;
  int iter, done, maxit, pad;
};

// Gather-scatter map (dssum): CSR segments of local copies of each shared node + halo lists.
struct GSMap {
  int nseg = 0;              // segments written back (local multiplicity > 1 or shared with another rank)
  int nseg_int = 0;          // the first nseg_int segments have no copy on another rank (summed while the halo is in flight)
  int* seg_off = nullptr;    // [nseg+1] into seg_idx
  int* seg_idx = nullptr;    // local dof indices
  // halo (multi-rank)
  int nnbr = 0;
  std::vector<int> nbr_rank, nbr_off;   // neighbour ranks, offsets into send/recv buffers (in shared nodes)
  int nshared = 0;           // total entries of the send (= recv) buffer per field
  int* send_seg = nullptr;   // [nshared] segment whose local sum goes to this send entry
  int* rseg_off = nullptr;   // [nseg+1] into rseg_pos: recv-buffer positions contributing to a segment
  int* rseg_pos = nullptr;
  int* rseg_nbefore = nullptr;  // [nseg] how many of those come from lower ranks (summed before the local part)
  int* send_base = nullptr;  // [nshared] position of every send entry inside the (NCCL) send buffer
  int* send_cnt = nullptr;   // [nshared] entries of its neighbour block (field stride)
  int* rseg_cnt = nullptr;   // parallel to rseg_pos
  double* sendbuf = nullptr; // [3*nshared]
  double* recvbuf = nullptr;
};

// NVLink peer-memory collectives (p2p.cu)
struct P2P {
  bool on = false;
  double* arena = nullptr;
  long long words = 0, halo_stride = 0;
  double* peer[16] = {nullptr};
  double** d_peer_dst = nullptr;                 // [2 parities][nnbr]
  unsigned long long** d_peer_flag = nullptr;    // [2 parities][nnbr]
  double** d_peer_base = nullptr;                // [16]
  int* d_cnt = nullptr;
  int* d_nbr_rank = nullptr;
  int* d_send_nbr = nullptr;
  int* d_send_j = nullptr;
  int* d_err = nullptr;
  unsigned long long* d_epoch = nullptr;         // device-resident exchange counters: [0] halo, [1] all-reduce
};

// Three-level additive pressure preconditioner (pmg.cu), one instance per mask set (direct / adjoint)
struct PMG {
  bool ready = false;
  int nv = 0;                       // local vertices of the element-vertex (Q1) mesh
  int nagg = 0, nagg_loc = 0, agg_first = 0, ncolours = 0;
  double* S = nullptr;              // [nel][ldim][lx2*lx2] FDM eigenvectors (row = nodal index, column = mode)
  double* lam = nullptr;            // [nel][ldim][lx2]     FDM eigenvalues times the box-geometry factor
  int* vid = nullptr;               // [nel][2^ldim] local vertex of every element corner
  int *voff = nullptr, *vent = nullptr;   // CSR vertex -> (element, corner) entries
  double* d1inv = nullptr;          // [nv] 1/diag(P^T E P)
  int* agg = nullptr;               // [nel] aggregate of every element (global aggregate id)
  int *aoff = nullptr, *aent = nullptr;   // CSR local aggregate -> elements
  double* A2inv = nullptr;          // [nagg][nagg] (Pa^T E Pa)^-1
  double *rc = nullptr, *xv = nullptr, *ra = nullptr, *x2 = nullptr;   // work: corner sums, vertex values, aggregate sums (+3 CG scalars)/values
  double* xc = nullptr;             // [nel][9] corner values of the vertex level + aggregate value per element (fused CG tail)
  double* hat = nullptr;            // [lx2] hat function (1+z)/2 at the GL points
  double* rc0 = nullptr;            // un-assembled copy of rc (multi-rank: rc is summed across ranks in place)
  std::vector<int> h_agg;
  std::vector<double> h_d1, h_A2inv;
};

struct Ctx {
  int ldim = 0, lx1 = 0, lxd = 0, lx2 = 0, nel = 0;
  long long nelg = 0;
  int np1 = 0, np2 = 0, npd = 0;
  long long n = 0, n2 = 0, nd = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  ConstMats cm;

  // geometry
  double* xyz[3] = {nullptr, nullptr, nullptr};
  double* R = nullptr;      // [d*d][n]    J*dr_i/dx_c at (i*d+c)
  double* jac = nullptr;
  double* bm1 = nullptr;
  double* binv = nullptr;   // 1/dssum(bm1)
  double* mult = nullptr;   // 1/multiplicity
  double* bm1s = nullptr;   // inner-product weight
  double* G = nullptr;      // [ng][n]  3D: 11,22,33,12,13,23 ; 2D: 11,22,12
  double* mask[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};     // [adjoint?][comp]
  double* mbinv[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};    // mask*binv
  bool has_adj_masks = false;
  bool mask_same[2] = {false, false};   // all components share one mask (one mask*binv array is streamed instead of ldim)
  double* RW2 = nullptr;    // [d*d][n2]  w2 * R interpolated to mesh 2
  double* bm2inv = nullptr; // 1/bm2
  double* Rd = nullptr;     // [d*d][nd]  wd * R interpolated to the dealiasing mesh
  double* hdiagA = nullptr; // assembled diag of stiffness A
  double* hdiagB = nullptr; // assembled bm1
  double* dinvH = nullptr;  // 1/(h1*hdiagA + h2*hdiagB) for the current h2
  double dinvH_h1 = -1, dinvH_h2 = -1;
  double* dinvE[2] = {nullptr, nullptr};
  // pressure preconditioner: 0 = Jacobi (north-star), 1 = three-level additive Schwarz/multilevel (pmg.cu)
  int pc_kind = 0, pc_nagg = 0;
  PMG pmg[2];
  double* pz = nullptr;       // [n2] z = M^-1 r
  double* ones2 = nullptr;    // [n2] all ones (lets the fused direction kernels read z in place of dinvE*r)
  std::vector<long long> vglo;   // [nel][2^ldim] global ids of the element corners (from glo_num)
  double vol = 0, vol2 = 0;
  long long n2_glob = 0;
  bool ifvcor[2] = {false, false};
  GSMap gs;
  P2P p2p;
  GSMap gsp;                 // the velocity-mesh map again, for vectors stored in the surface-first element layout (elem_common.cuh)
  P2P p2pp;
  bool gsp_ready = false;
  bool perm_h = false;       // surface-first layout of w in the Helmholtz CG loop (3-D, lx1 = 8; NSB_PERM=0 disables)
  bool ax_persistent = true;   // NSB_AX_PERSISTENT=0 selects the one-element-per-CTA k_axhelm3 (the only form for lx1 != 8)
  GSMap gsv;                 // gather-scatter over the element-vertex mesh (pressure preconditioner, multi-rank)
  P2P p2pv;                  // its own peer-memory halo channel
  bool gsv_ready = false;
  std::vector<int> pc_col;   // distance-2 colouring of the local vertices (consistent across ranks), reused by both mask sets
  int pc_ncol = 0;

  // parameters
  double visc = 1.0, rho = 1.0, tol_v = 1e-9, tol_p = 1e-7;
  int maxit_v = 1000, maxit_p = 20000;
  double dt = 0;
  int nsteps = 0;
  int check_every_v = 4, check_every_p = 32;
  bool persistent_pcg = true; // 3-D: persistent, TMA-pipelined div kernels in the pressure-CG loop (NSB_PERSISTENT=0 disables)
  bool pcg_fused = true;      // 3-D, pc_kind 1: fused CG tail (pm_pcg_tail) + coarse levels added in the direction kernel (NSB_PCG_FUSED=0 disables)

  double* adv_scratch = nullptr;   // per-CTA fine-mesh work arrays of the advection kernel (L2-resident)
  size_t adv_scratch_words = 0;

  // base flow, sponge
  double* ub = nullptr;     // [d][n]  the base flow the explicit terms see (Floquet: re-pointed into the stored orbit every step)
  double* ub0 = nullptr;    // [d][n]  the base flow given by nsb_set_baseflow (owner of the allocation)
  double spng_str_dns = 0.0; // DNS sponge strength (jp = 0 branch of nekStab_forcing, core/utils.f:166-171); 0 = off
  double* spng_ref = nullptr; // [d][n] its reference field spng_vr (core/utils.f:240)
  double* pb0 = nullptr;    // [n2]    its pressure (optional; initial pressure of the co-evolving base flow)
  // Floquet / UPO: ifbase co-evolution of the base flow with the full Navier-Stokes stepper and orbit storage (core/matvec.f:187-236)
  bool floquet = false, orbit_ready = false;
  double* orbit = nullptr;  // [orbit_steps][d][n]  uor, vor, wor (core/krylov_subspace.f:18)
  int orbit_steps = 0;
  // Scalar transport (ifheat, ldimt = 1; core/krylov_subspace.f:13,41-45): theta travels in every Krylov vector between the velocity
  // and the pressure, [vx|vy|(vz)|theta|pr]; advanced next to the velocity by the same BDF/EXT scheme (csrc/scalar.cu, stepper.cu)
  struct Scalar {
    bool on = false;
    double cond = 0.0, rhocp = 1.0;      // param(8), param(7): Helmholtz h1 and the factor of the time derivative / convection
    double ri = 0.0;                     // buoyancy f_g += ri * theta (uparam(6) of the shipped .usr files)
    int gdir = 1;                        // component that feels the buoyancy (y)
    double* tmask = nullptr;             // [n] Dirichlet mask of the scalar
    double* tb = nullptr;                // [n] base scalar field (tbase)
    double* th = nullptr;                // [n] current
    double* thlag[2] = {nullptr, nullptr};
    double* q[3] = {nullptr, nullptr, nullptr};   // explicit terms ring
    double* wk[2] = {nullptr, nullptr};  // [n] work
    double* mats = nullptr;              // Jd, Dd, Jdt (lxd x lx1 each) for the convection kernel
  } scal;
  long long poff = 0;                    // offset of the pressure inside a Krylov vector: n * (ldim + nscal)
  // UPO Newton (uparam(1) = 2.1): Krylov vectors carry a time component (the period unknown, core/krylov_subspace.f:8-15, :47-50); the
  // border vectors compute_bvec(fc_nwt), compute_bvec(ic_nwt) (core/matvec.f:407-419, 435-475) stay resident next to the orbit
  bool upo = false, bvec_ready = false;
  std::vector<double> slot_time;   // [nslots] host: q%time of every slot
  double* bvec = nullptr;          // [2][vlen]: bvec(fc_nwt), bvec(ic_nwt)
  struct StepState { double* u = nullptr; double* ulag[2] = {nullptr, nullptr}; double* f[3] = {nullptr, nullptr, nullptr}; double* pr = nullptr; double* prlag = nullptr; };
  StepState base_state;     // time-stepper state of the co-evolving base flow
  double* spng = nullptr;   // [n] or null

  // stepper state (all device)
  double* u = nullptr;      // [d][n]
  double* ulag[2] = {nullptr, nullptr};
  double* f[3] = {nullptr, nullptr, nullptr};   // explicit terms ring: f[0] current
  double* pr = nullptr;     // [n2]
  double* prlag = nullptr;
  double* pt = nullptr;     // extrapolated pressure
  double* wk[4] = {nullptr, nullptr, nullptr, nullptr};  // [d][n] work: rhs/residual, cg p, cg w, cg x
  double* rk = nullptr;     // [d][n] cg r
  double* pk[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [n2] work: g/r, x, p, Ep, spare
  CGState* cgs = nullptr;   // [4] device: 3 Helmholtz comps + pressure
  CGState* cgs_host = nullptr;  // pinned

  // reductions
  double* red_part = nullptr;   // [NSB_MAX_BLOCKS*NSB_MAX_RED]
  double* red_out = nullptr;    // [NSB_MAX_RED*4] device results
  double* red_host = nullptr;   // pinned host copy
  unsigned* red_count = nullptr;

  // krylov slab
  double* slab = nullptr;
  int nslots = 0;
  long long vlen = 0;       // d*n + n2
  double* hbuf = nullptr;   // device coefficient buffer (>= 4096 doubles)
  double* hpart = nullptr;  // partial multi-dot sums
  long long hpart_cap = 0;

  // pressure residual projection (Nek5000 `residualProj = yes`: setrhsp / gensolnp [UPSTREAM navier4.f]; every shipped
  // .par enables it, e.g. 1cyl.par:30): E-orthonormal basis of previous solutions and their images under E
  int proj_max = 0, proj_m = 0, proj_adj = -1;
  double* projX = nullptr;    // [proj_max][n2]
  double* projEX = nullptr;   // [proj_max][n2]

  // CUDA graphs of the CG iteration batches (single rank): one captured batch = `check_every` iterations; the launch-bound
  // 2-D / small cases replay it instead of issuing ~130 kernel launches per batch from the host
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int launches = 0; int adj = -1; double h1 = 0, h2 = 0; };
  bool use_graphs = true;
  GraphEntry graph_p[2];
  GraphEntry graph_h[8];
  bool graph_warm = false;

  // per-kernel sampling profiler (CUDA events on the launching stream; one sample set per host poll)
  int prof_on = 0;
  cudaEvent_t prof_ev[24] = {nullptr};
  double prof_ms[16] = {0};
  long long prof_cnt[16] = {0};

  // per-step host hook (the reference's nekstab_usrchk before every nek_advance, core/matvec.f:221)
  nsb_step_callback step_cb = nullptr;
  void* step_cb_user = nullptr;

  // stats
  nsb_stats stats = {0, 0, 0, 0, 0.0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

extern Ctx* g_ctx;
bool upo_active();          // nsb_set_upo(1) in effect (api.cu)
inline double* slot_ptr(Ctx* c, int s) { return c->slab + (long long)s * c->vlen; }
// The loop vector w = H p of the Helmholtz CG travels k_axhelm3p -> dssum -> k_hcg_update in the surface-first element layout
// (elem_common.cuh).  (r2, measured: the same layout in the pressure loop made the dssum 0.022 ms faster but k_gradt3's strided stores
// and k_div3q's shared-memory reads 0.021 ms slower -- removed there.)
inline bool perm_h_active(const Ctx* c) { return c->perm_h && c->ldim == 3 && c->lx1 == 8 && c->ax_persistent; }

// ---- host SEM (sem_host.cpp)
void sem_zwgll(int n, double* z, double* w);
void sem_zwgl(int n, double* z, double* w);
void sem_deriv(int n, const double* x, double* D);                                // row-major D[i*n+l]
void sem_interp(int nto, const double* xto, int nfrom, const double* xfrom, double* J);  // J[i*nfrom+l]
void sem_build_constmats(int lx1, int lx2, int lxd, ConstMats* cm);

// ---- gather-scatter (gs.cu)
int gs_setup(Ctx* c, const long long* glo_num);
int gs_free(Ctx* c);
int gs_dssum(Ctx* c, double* u, int nfields, long long stride, const CGState* skip_if_done = nullptr);
int gs_build(Ctx* c, GSMap& m, P2P& p2p, long long n, int N1, int np_e, const long long* glo, int surf_first = 0);
int gs_dssum_w(Ctx* c, double* w, int nfields, bool permuted, const CGState* skip);   // dssum of a loop vector in natural / surface-first layout
int gs_dssum_map(Ctx* c, GSMap& m, P2P& p2p, double* u, int nfields, long long stride, const CGState* skip_if_done);
int gs_free_map(Ctx* c, GSMap& m, P2P& p);

// ---- NVLink peer-memory collectives (p2p.cu)
int p2p_setup(Ctx* c, P2P& p, const GSMap& m, const std::vector<int>& send_nbr, const std::vector<int>& send_j);
int p2p_free(Ctx* c, P2P& p);
int p2p_dssum(Ctx* c, P2P& p, GSMap& m, double* u, int nfields, long long stride, const CGState* skip);
int p2p_allreduce(Ctx* c, double* dev, int count, int op, CGState* cgs, int ncomp, int kind);
int p2p_check_error(Ctx* c);

// ---- element kernels (elem_kernels.cu)
int ek_upload_constants(const ConstMats& cm);
int ek_geometry(Ctx* c);                      // metrics, G, bm1, mesh-2 and fine metrics, diag(A)
int ek_axhelm(Ctx* c, const double* u, double* w, int nfields, double h1, double h2);
// r[c] = b[c] + r[c] - (h1 A + h2 B) u[c]   (residual assembly of cresvipp)
int ek_axhelm_resid(Ctx* c, const double* u, const double* b, double* r, int nfields, double h1, double h2);
int ek_gradt(Ctx* c, const double* p, double* w);                       // w[d][n] = D^T p
int ek_div(Ctx* c, const double* u, const double* scale /*[d][n] or null*/, double* q, double sign);
int ek_advab(Ctx* c, int adjoint, const double* up, const double* ub, const double* spng, double* f);  // f = -adv - bm1*spng*up
int ek_ediag(Ctx* c, int adj);
int ek_cfl(Ctx* c, const double* u, double* cfl_dev);                  // max reduction into device scalar
// CG building blocks
int ek_hcg_dir_ax(Ctx* c, int ncomp, double h1, double h2);     // p = dinv r + beta p; w = H p; rho partial = sum p*w
int ek_pcg_dir_gradt(Ctx* c, int adj);                           // p = dinvE r + beta p ; w = gradt(p)
int ek_pcg_div(Ctx* c, int adj);                                 // Ep = div(mbinv w); rho = sum p Ep
int ek_div_mbinv(Ctx* c, const double* u, int adj, double* q, double sign);   // q = sign * D (mask*binv .* u)

// ---- second-generation 3-D pressure-operator kernels (pcg_kernels.cu)
int pk_upload_constants(const ConstMats& cm);
int pk_gradt(Ctx* c, const double* p, double* w);
int pk_pcg_dir_gradt(Ctx* c, int adj);
int pk_div(Ctx* c, const double* u, const double* s0, const double* s1, const double* s2, double* q, double sign);
int pk_pcg_div(Ctx* c, int adj, int fused);
int pk_axhelm(Ctx* c, int mode, const double* u, double* w, const double* b, int nfields, double h1, double h2);

// ---- pointwise / reduction kernels (vec_kernels.cu)
int vk_fill(Ctx* c, double* a, double v, long long n);
int vk_copy(Ctx* c, double* dst, const double* src, long long n);
int vk_scale(Ctx* c, double* a, double s, long long n);
int vk_axpy(Ctx* c, double* y, double a, const double* x, long long n);        // y += a x
int vk_mul(Ctx* c, double* a, const double* b, long long n);                     // a *= b
int vk_inv(Ctx* c, double* a, long long n);                                      // a = 1/a
int vk_lin2(Ctx* c, double* out, double a, const double* x, double b, const double* y, long long n);  // out = a x + b y
int vk_sum(Ctx* c, const double* a, long long n, double* out_dev);               // deterministic sum -> device scalar
int vk_dot3(Ctx* c, const double* a, const double* b, const double* w, long long n, double* out_dev);  // sum a*b*w (w may be null)
int vk_allreduce_sum(Ctx* c, double* dev, int count);
int vk_allreduce_max(Ctx* c, double* dev, int count);
int vk_add_scalar_from_dev(Ctx* c, double* a, const double* s_dev, double factor, long long n);  // a += factor * (*s_dev)
// stepper pointwise
int vk_make_rhs(Ctx* c, double* b, int k, const double* ab, const double* bd);   // EXT/BDF assembly
int vk_mask_fields(Ctx* c, double* r, int adj);                                   // r[c] *= mask[c]
int vk_press_extrap(Ctx* c, int k);
int vk_final_update(Ctx* c, int adj, double h2);   // u = u + du + mbinv*w ; p = pt + h2*phi
// CG pointwise pieces
int vk_hcg_init(Ctx* c, int ncomp);
int vk_hcg_update(Ctx* c, int ncomp, int adj);   // reads w in the surface-first layout when perm_h_active(c)
int vk_pcg_init(Ctx* c, int adj);
int vk_pcg_update(Ctx* c, int adj);
int vk_dinvH(Ctx* c, double h1, double h2);
// krylov
int vk_multidot(Ctx* c, int k, int first_slot, int slot_f, double* h_dev);      // h = Q^T (W f)
int vk_multiaxpy(Ctx* c, int k, int first_slot, int slot_f, const double* h_dev, double sign);  // f += sign * Q h
int vk_gemv_out(Ctx* c, int k, int first_slot, const double* y_dev, int slot_out);
int vk_multidot_raw(Ctx* c, int k, const double* Q, long long vlen, const double* f, const double* W, long long n, long long nw,
                    double* h_dev);
int vk_multiaxpy_raw(Ctx* c, int k, const double* Q, long long vlen, const double* f, double a, const double* h_dev, double sign,
                     double* out);
int vk_rotate(Ctx* c, int k, int first_slot, const double* S_dev, int lds);
int vk_fp64_peak(Ctx* c, double* tflops);
int vk_sponge_dns(Ctx* c, double* f, const double* u);       // f += bm1 * spng_str * spng_fun * (spng_vr - u)
int vk_rotate_pair(Ctx* c, double* are, double* aim, double g, double d, long long n);
int vk_wavemaker(Ctx* c, const double* dre, const double* dim, const double* are, const double* aim, double* wm);

// ---- pressure preconditioner (pmg.cu)
int pm_setup(Ctx* c, int set, int nagg_req);
int pm_apply(Ctx* c, int set, const double* r, double* z, int mode, int prof_slot = 0);
int pm_pcg_tail(Ctx* c, int set, int init, int prof_slot);
int pm_pcg_xfix(Ctx* c);                                     // x += alpha p of the last iteration (fused path)   // fused CG update + restriction + element blocks + coarse levels + scalars (3-D)
void pm_free(PMG& m);
int vk_cg_finalize_multi(Ctx* c, CGState* s, int ncomp, int kind);

// ---- solvers / stepper (stepper.cu)
int st_alloc(Ctx* c);
int st_helmholtz(Ctx* c, int adj, double h1, double h2, int* iters, int ncomp = 0, int graph_key = -1);
// scalar.cu: out = J^T [ (Rd J a).grad(J phi) + (Rd J b).grad(J psi) ]  (dealiased convection of a scalar; b, psi may be null)
int sk_setup(Ctx* c);
int sk_conv(Ctx* c, const double* a, const double* phi, const double* b, const double* psi, double* out);   // solves for wk[3] (x) from rk (r, assembled+masked)
int st_pressure(Ctx* c, int adj, int* iters);                          // solves E pk[1] = pk[0]
int st_linearized_map(Ctx* c, int adjoint, const double* vin, double* vout);   // vin/vout device krylov vectors

// ---- host krylov (host_krylov.cpp)
int nsb_lapack_spd_inverse(int n, double* A);   // in place, dpotrf + dpotri; non-zero when not positive definite / LAPACK missing
void nsb_count_launch(int n = 1);
