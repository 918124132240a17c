// host_krylov.cpp -- host-side Krylov drivers, mirroring nekStab's Fortran one to one above the C ABI.
//
// In a drop-in build these stay Fortran (core/krylov_decomposition.f, core/eigensolvers.f, core/newton_krylov.f call the
// ISO_C_BINDING shim); no Fortran compiler exists in this image, so the same control flow is written here in C++
// (task rule: host side in C++ where the reference is compiled code).  Small dense work goes to host LAPACK
// exactly as core/lapack_wrapper.f does (dgeev, dgees, dtrsen, dgels), resolved at run time with dlopen from any
// library exporting the LP64 Fortran symbols (`dgeev_` or scipy-openblas' `scipy_dgeev_`).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <numeric>

#include "nsb_internal.h"

typedef int lint;       // LP64 LAPACK integer
typedef int llogical;   // Fortran LOGICAL
typedef llogical (*select_fn)(const double*, const double*);
typedef void (*dgeev_t)(const char*, const char*, const lint*, double*, const lint*, double*, double*, double*, const lint*,
                        double*, const lint*, double*, const lint*, lint*, size_t, size_t);
typedef void (*dgees_t)(const char*, const char*, select_fn, const lint*, double*, const lint*, lint*, double*, double*,
                        double*, const lint*, double*, const lint*, llogical*, lint*, size_t, size_t);
typedef void (*dtrsen_t)(const char*, const char*, const llogical*, const lint*, double*, const lint*, double*, const lint*,
                         double*, double*, lint*, double*, double*, double*, const lint*, lint*, const lint*, lint*, size_t,
                         size_t);
typedef void (*dgels_t)(const char*, const lint*, const lint*, const lint*, double*, const lint*, double*, const lint*,
                        double*, const lint*, lint*, size_t);

typedef void (*dpotrf_t)(const char*, const lint*, double*, const lint*, lint*, size_t);
typedef void (*dpotri_t)(const char*, const lint*, double*, const lint*, lint*, size_t);
static dpotrf_t p_dpotrf = nullptr;
static dpotri_t p_dpotri = nullptr;

static void* g_lapack = nullptr;
static dgeev_t p_dgeev = nullptr;
static dgees_t p_dgees = nullptr;
static dtrsen_t p_dtrsen = nullptr;
static dgels_t p_dgels = nullptr;

static void* sym2(void* h, const char* a, const char* b) {
  void* s = dlsym(h, a);
  return s ? s : dlsym(h, b);
}

extern "C" int nsb_lapack_load(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { nsb_set_error("nsb_lapack_load: %s", dlerror()); return 1; }
  dgeev_t a = (dgeev_t)sym2(h, "dgeev_", "scipy_dgeev_");
  dgees_t b = (dgees_t)sym2(h, "dgees_", "scipy_dgees_");
  dtrsen_t t = (dtrsen_t)sym2(h, "dtrsen_", "scipy_dtrsen_");
  dgels_t l = (dgels_t)sym2(h, "dgels_", "scipy_dgels_");
  if (!a || !b || !t || !l) { nsb_set_error("nsb_lapack_load: %s lacks dgeev_/dgees_/dtrsen_/dgels_", path); dlclose(h); return 1; }
  g_lapack = h; p_dgeev = a; p_dgees = b; p_dtrsen = t; p_dgels = l;
  p_dpotrf = (dpotrf_t)sym2(h, "dpotrf_", "scipy_dpotrf_");      // optional: large aggregate operators of the pressure preconditioner
  p_dpotri = (dpotri_t)sym2(h, "dpotri_", "scipy_dpotri_");
  return 0;
}

static int need_lapack() {
  if (p_dgeev) return 0;
  const char* env = getenv("NSB_LAPACK_LIB");
  if (env && nsb_lapack_load(env) == 0) return 0;
  const char* cands[] = {"liblapack.so.3", "libopenblas.so.0", "liblapack.so", "libopenblas.so"};
  for (const char* cnd : cands)
    if (nsb_lapack_load(cnd) == 0) return 0;
  nsb_set_error("no LAPACK found: set NSB_LAPACK_LIB or call nsb_lapack_load(path) (e.g. scipy.libs/libscipy_openblas-*.so)");
  return 1;
}

// ------------------------------------------------------------------ core/lapack_wrapper.f
static llogical select_eigvals(const double* wr, const double* wi) {   // :258-270
  return std::sqrt((*wr) * (*wr) + (*wi) * (*wi)) > 0.9;
}

// eig (:129-202): dgeev, complex eigenvectors from the real pairs, then sort by decreasing magnitude (:204-256)
extern "C" int nsb_lapack_eig(const double* A, int n, double* vals_re, double* vals_im, double* vecs_reim) {
  NSB_TRY(need_lapack());
  std::vector<double> At(A, A + (size_t)n * n), wr(n), wi(n), vr((size_t)n * n), vl(n), work(4 * (size_t)n);
  lint nn = n, ldvl = 1, lwork = 4 * n, info = 0;
  p_dgeev("N", "V", &nn, At.data(), &nn, wr.data(), wi.data(), vl.data(), &ldvl, vr.data(), &nn, work.data(), &lwork, &info, 1, 1);
  if (info != 0) { nsb_set_error("dgeev info=%d", info); return 1; }
  typedef std::complex<double> cd;
  std::vector<cd> vals(n), vecs((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    vals[i] = cd(wr[i], wi[i]);
    for (int r = 0; r < n; ++r) vecs[(size_t)i * n + r] = cd(vr[(size_t)i * n + r], 0.0);
  }
  for (int i = 0; i < n - 1; ++i) {
    if (wi[i] > 0) {
      for (int r = 0; r < n; ++r) {
        vecs[(size_t)i * n + r] = cd(vr[(size_t)i * n + r], vr[(size_t)(i + 1) * n + r]);
        vecs[(size_t)(i + 1) * n + r] = cd(vr[(size_t)i * n + r], -vr[(size_t)(i + 1) * n + r]);
      }
    } else if (wi[i] == 0) {
      for (int r = 0; r < n; ++r) vecs[(size_t)i * n + r] = cd(vr[(size_t)i * n + r], 0.0);
    }
  }
  // sort_eigendecomp: the reference's O(n^2) exchange sort, same comparisons => same order for ties
  std::vector<double> nrm(n);
  for (int i = 0; i < n; ++i) nrm[i] = std::sqrt(vals[i].real() * vals[i].real() + vals[i].imag() * vals[i].imag());
  for (int k = 0; k < n - 1; ++k)
    for (int l = k + 1; l < n; ++l)
      if (nrm[k] < nrm[l]) {
        std::swap(nrm[k], nrm[l]);
        std::swap(vals[k], vals[l]);
        for (int r = 0; r < n; ++r) std::swap(vecs[(size_t)k * n + r], vecs[(size_t)l * n + r]);
      }
  for (int i = 0; i < n; ++i) {
    vals_re[i] = vals[i].real(); vals_im[i] = vals[i].imag();
    for (int r = 0; r < n; ++r) {
      vecs_reim[2 * ((size_t)i * n + r)] = vecs[(size_t)i * n + r].real();
      vecs_reim[2 * ((size_t)i * n + r) + 1] = vecs[(size_t)i * n + r].imag();
    }
  }
  return 0;
}

// schur (:7-63): dgees with sort='S' and select_eigvals (|lambda| > 0.9 first)
extern "C" int nsb_lapack_schur(double* A, int n, double* vecs, double* vals_re, double* vals_im) {
  NSB_TRY(need_lapack());
  lint nn = n, sdim = 0, lwork = std::max(1, 3 * n), info = 0;
  std::vector<double> work(lwork);
  std::vector<llogical> bwork(n);
  p_dgees("V", "S", select_eigvals, &nn, A, &nn, &sdim, vals_re, vals_im, vecs, &nn, work.data(), &lwork, bwork.data(), &info, 1, 1);
  if (info != 0 && info <= n) { nsb_set_error("dgees info=%d", info); return 1; }
  return 0;
}

// ordschur (:70-123): dtrsen job='N', compq='V'
extern "C" int nsb_lapack_ordschur(double* T, double* Q, const int* selected, int n) {
  NSB_TRY(need_lapack());
  lint nn = n, m = 0, lwork = std::max(1, n), liwork = 1, info = 0, iwork[1] = {0};
  double s = 0, sep = 0;
  std::vector<double> wr(n), wi(n), work(lwork);
  std::vector<llogical> sel(selected, selected + n);
  p_dtrsen("N", "V", sel.data(), &nn, T, &nn, Q, &nn, wr.data(), wi.data(), &m, &s, &sep, work.data(), &lwork, iwork, &liwork, &info, 1, 1);
  if (info != 0) { nsb_set_error("dtrsen info=%d", info); return 1; }
  return 0;
}

// lstsq (:287-339): dgels, min ||A x - b||, A (m x n) column-major
extern "C" int nsb_lapack_lstsq(const double* A, const double* b, double* x, int m, int n) {
  NSB_TRY(need_lapack());
  std::vector<double> At(A, A + (size_t)m * n), bt(b, b + m), work(2 * (size_t)m * n + 64);
  lint mm = m, nn = n, nrhs = 1, lwork = (lint)work.size(), info = 0;
  p_dgels("N", &mm, &nn, &nrhs, At.data(), &mm, bt.data(), &mm, work.data(), &lwork, &info, 1);
  if (info != 0) { nsb_set_error("dgels info=%d (Least-Squares solver UNsuccessful)", info); return 1; }
  for (int i = 0; i < n; ++i) x[i] = bt[i];
  return 0;
}

// dense SPD inverse in place (column-major = row-major for a symmetric matrix): dpotrf + dpotri, lower triangle mirrored
int nsb_lapack_spd_inverse(int n, double* A) {
  NSB_TRY(need_lapack());
  if (!p_dpotrf || !p_dpotri) { nsb_set_error("LAPACK library lacks dpotrf_/dpotri_"); return 1; }
  lint nn = n, info = 0;
  p_dpotrf("L", &nn, A, &nn, &info, 1);
  if (info != 0) { nsb_set_error("dpotrf info=%d (matrix not positive definite)", info); return 2; }
  p_dpotri("L", &nn, A, &nn, &info, 1);
  if (info != 0) { nsb_set_error("dpotri info=%d", info); return 2; }
  for (int j = 0; j < n; ++j)                  // column-major lower triangle: A(i,j), i >= j, at A[j*n+i]
    for (int i = j + 1; i < n; ++i) A[(size_t)i * n + j] = A[(size_t)j * n + i];
  return 0;
}

// ------------------------------------------------------------------ core/krylov_decomposition.f:7-104
// Q(i) = slot first_slot + i - 1 (1-based i as in the reference); H column-major with leading dimension ldh.
extern "C" int nsb_arnoldi_factorization(int mode, int first_slot, double* H, int ldh, int mstart, int mend, int ksize) {
  if (ksize == 0) { nsb_set_error("Krylov base dimension == 0! Increase it.. STOP"); return 1; }   // :64-67
  std::vector<double> hcol(ksize + 2);
  for (int mstep = mstart; mstep <= mend; ++mstep) {
    const int q = first_slot + mstep - 1, f = first_slot + mstep;      // f is built directly in Q(mstep+1)
    NSB_TRY(nsb_matvec(mode, q, f));                                    // :82
    NSB_TRY(nsb_orthonormalize(mstep, first_slot, f, hcol.data()));     // :85 update_hessenberg_matrix
    for (int i = 0; i <= mstep; ++i) H[(size_t)(mstep - 1) * ldh + i] = hcol[i];
  }
  return 0;
}

// ------------------------------------------------------------------ core/eigensolvers.f:729-795
extern "C" int nsb_select_eigenvalues(int* selected, int* cnt, const double* vre, const double* vim, double delta, int nev, int n) {
  std::vector<double> mag(n);
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) { mag[i] = std::hypot(vre[i], vim[i]); idx[i] = i; }
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return mag[a] < mag[b]; });   // ascending, like quicksort2
  for (int i = 0; i < n; ++i) selected[i] = mag[i] >= (1.0 - delta);
  for (int j = std::max(0, n - (nev + 4)); j < n; ++j) selected[idx[j]] = 1;                  // idx(n-(nev+3):n)
  if (n - (nev + 5) >= 0) {
    int a = idx[n - (nev + 4)], b = idx[n - (nev + 5)];
    if (vim[a] == -vim[b]) selected[b] = 1;                                                   // keep conjugate pairs together
  }
  int c = 0;
  for (int i = 0; i < n; ++i) c += selected[i] ? 1 : 0;
  *cnt = c;
  return 0;
}

// ------------------------------------------------------------------ core/eigensolvers.f:395-499
extern "C" int nsb_schur_condensation(int* mstart, double* H, int ldh, int first_slot, int ksize, int schur_tgt, double schur_del) {
  const int k = ksize;
  std::vector<double> T((size_t)k * k), vecs((size_t)k * k, 0.0), wr(k), wi(k), b_vec(k, 0.0);
  b_vec[k - 1] = H[(size_t)(k - 1) * ldh + k];                           // H(ksize+1, ksize)
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) T[(size_t)j * k + i] = H[(size_t)j * ldh + i];
  NSB_TRY(nsb_lapack_schur(T.data(), k, vecs.data(), wr.data(), wi.data()));
  std::vector<int> selected(k);
  int ms = 0;
  NSB_TRY(nsb_select_eigenvalues(selected.data(), &ms, wr.data(), wi.data(), schur_del, schur_tgt, k));
  NSB_TRY(nsb_lapack_ordschur(T.data(), vecs.data(), selected.data(), k));
  // zero the unwanted blocks (:449-450) and write back
  for (int j = 0; j < k; ++j)
    for (int i = 0; i <= k; ++i) {
      double v = (i < k) ? T[(size_t)j * k + i] : 0.0;
      if (i < ms && j >= ms) v = 0.0;
      if (i >= ms) v = 0.0;
      H[(size_t)j * ldh + i] = v;
    }
  // Q(:,1:k) <- Q(:,1:k) * vecs   (:466-474)
  NSB_TRY(nsb_basis_rotate(k, first_slot, vecs.data(), k));
  // b_vec = b_vec * vecs ; H(mstart+1,:) = b_vec   (:478-479)
  for (int j = 0; j < k; ++j) {
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += b_vec[i] * vecs[(size_t)j * k + i];
    H[(size_t)j * ldh + ms] = s;
  }
  ms += 1;                                                                // :482
  NSB_TRY(nsb_vec_copy(first_slot + ms - 1, first_slot + k));             // Q(mstart) = Q(ksize+1)   :484-485
  *mstart = ms;
  return 0;
}

// ------------------------------------------------------------------ core/eigensolvers.f:141-388 (Krylov-Schur loop, :335-373)
// Q(1) must already hold the normalised seed in slot seed_slot = first slot of a (k_dim+1)-slot range.
extern "C" int nsb_krylov_schur(int mode, int k_dim, int schur_tgt, double eigen_tol, double schur_del, int seed_slot,
                                double* vals_re, double* vals_im, double* residual, double* vecs_reim, int* n_converged,
                                int* schur_cnt, int max_restarts) {
  const int ldh = k_dim + 1;
  std::vector<double> H((size_t)ldh * k_dim, 0.0), Hk((size_t)k_dim * k_dim);
  int mstart = 1, cnt = 0, scnt = 0;
  bool converged = false;
  while (!converged) {
    NSB_TRY(nsb_arnoldi_factorization(mode, seed_slot, H.data(), ldh, mstart, k_dim, k_dim));
    for (int j = 0; j < k_dim; ++j)
      for (int i = 0; i < k_dim; ++i) Hk[(size_t)j * k_dim + i] = H[(size_t)j * ldh + i];
    NSB_TRY(nsb_lapack_eig(Hk.data(), k_dim, vals_re, vals_im, vecs_reim));
    const double hlast = H[(size_t)(k_dim - 1) * ldh + k_dim];
    cnt = 0;
    for (int i = 0; i < k_dim; ++i) {                                     // residual = |H(k+1,k) * vecs(k,:)|  :349
      double re = vecs_reim[2 * ((size_t)i * k_dim + k_dim - 1)], im = vecs_reim[2 * ((size_t)i * k_dim + k_dim - 1) + 1];
      residual[i] = std::fabs(hlast) * std::hypot(re, im);
      if (residual[i] < eigen_tol) ++cnt;
    }
    if (schur_tgt <= 0) converged = true;                                  // :357-359
    else if (cnt >= schur_tgt) converged = true;                           // :363-364
    else {
      if (max_restarts >= 0 && scnt >= max_restarts) break;
      ++scnt;
      NSB_TRY(nsb_schur_condensation(&mstart, H.data(), ldh, seed_slot, k_dim, schur_tgt, schur_del));
    }
  }
  if (n_converged) *n_converged = cnt;
  if (schur_cnt) *schur_cnt = scnt;
  return converged ? 0 : 3;
}

// ------------------------------------------------------------------ core/newton_krylov.f:175-297 (+ :305-328)
// Solves A sol = rhs with A = matvec(mode); Q(1..ksize+1) = slots first_slot.. ; work_slot = scratch vector.
extern "C" int nsb_ts_gmres(int mode, int rhs_slot, int sol_slot, int first_slot, int work_slot, int maxiter, int ksize,
                            double tol, int* calls, double* final_res) {
  const int ldh = ksize + 1;
  std::vector<double> H((size_t)ldh * ksize), yvec(ksize), evec(ksize + 1);
  double beta = 0.0;
  NSB_TRY(nsb_vec_zero(sol_slot));
  NSB_TRY(nsb_vec_copy(first_slot, rhs_slot));
  NSB_TRY(nsb_vec_normalize(first_slot, &beta));                          // :236-237
  int ncalls = 0;
  for (int it = 1; it <= maxiter; ++it) {
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(yvec.begin(), yvec.end(), 0.0);
    std::fill(evec.begin(), evec.end(), 0.0);
    evec[0] = beta;
    int k = 1, kused = ksize;
    for (k = 1; k <= ksize; ++k) {
      NSB_TRY(nsb_arnoldi_factorization(mode, first_slot, H.data(), ldh, k, k, ksize));     // :255
      std::vector<double> Hs((size_t)(k + 1) * k);
      for (int j = 0; j < k; ++j)
        for (int i = 0; i <= k; ++i) Hs[(size_t)j * (k + 1) + i] = H[(size_t)j * ldh + i];
      NSB_TRY(nsb_lapack_lstsq(Hs.data(), evec.data(), yvec.data(), k + 1, k));             // :258
      double r2 = 0.0;
      for (int i = 0; i <= k; ++i) {
        double s = evec[i];
        for (int j = 0; j < k; ++j) s -= Hs[(size_t)j * (k + 1) + i] * yvec[j];
        r2 += s * s;
      }
      beta = std::sqrt(r2);                                                                 // :261
      if (beta * beta < tol) { ncalls += k; kused = k; break; }                             // :268-271 (squared norm vs tol)
    }
    if (k > ksize) kused = ksize;
    NSB_TRY(nsb_basis_gemv(kused, first_slot, yvec.data(), work_slot));                     // krylov_matmul :275
    NSB_TRY(nsb_vec_add2(sol_slot, work_slot));                                              // :276
    // initialize_gmres_vector (:305-328): q = (rhs - A sol)/beta
    NSB_TRY(nsb_vec_copy(first_slot, sol_slot));
    NSB_TRY(nsb_matvec(mode, first_slot, work_slot));
    NSB_TRY(nsb_vec_sub2(work_slot, rhs_slot));
    NSB_TRY(nsb_vec_cmult(work_slot, -1.0));
    NSB_TRY(nsb_vec_normalize(work_slot, &beta));
    NSB_TRY(nsb_vec_copy(first_slot, work_slot));
    if (beta * beta < tol) break;                                                            // :288
  }
  if (calls) *calls = ncalls;
  if (final_res) *final_res = beta * beta;
  return 0;
}

// ------------------------------------------------------------------ core/newton_krylov.f:5-168 (fixed points, uparam(1) = 2; UPOs, 2.1, after nsb_set_upo(1))
// q_slot: current estimate (in/out).  Slots used: f_slot, dq_slot, work_slot, and first_slot..first_slot+k_dim for the GMRES
// basis.  Residuals are SQUARED norms compared with tol, as in the reference (:99,:109).  hist (optional, maxiter_newton
// entries) receives the residual of every Newton iteration (residu_newton.dat).
extern "C" int nsb_newton_krylov(int q_slot, int f_slot, int dq_slot, int work_slot, int first_slot, int k_dim, double end_time,
                                 double cfl_target, double tol, int maxiter_newton, int maxiter_gmres, int* newton_iters,
                                 double* residual_out, double* hist, long long* calls_out) {
  double residual = 0.0;
  long long calls_counter = 0;
  int it = 0;
  const bool upo = upo_active();                           // uparam(1) = 2.1: q%time is the period, an unknown of the iteration
  for (it = 1; it <= maxiter_newton; ++it) {
    double dt = 0, ct = 0;
    int nsteps = 0;
    if (upo) {                                             // :63-67 first guess = endTime, later guesses travel in q%time
      if (it == 1) NSB_TRY(nsb_vec_set_time(q_slot, end_time));
      else NSB_TRY(nsb_vec_get_time(q_slot, &end_time));
      if (!(end_time > 0)) { nsb_set_error("UPO Newton: the period became %g", end_time); return 2; }
    }
    // :69.  The reference's prepare_linearized_solver takes the CFL of whatever Nek holds in vx,vy,vz: the initial q at iteration 1, from
    // iteration 2 on the base flow of the LAST matvec (= the previous iterate, core/matvec.f:103).  Here it is the CURRENT iterate; the
    // two differ by the Newton update, i.e. they can give a different integer nsteps only far from convergence.
    NSB_TRY(nsb_prepare_solver_from_slot(q_slot, end_time, cfl_target, &dt, &nsteps, &ct));
    NSB_TRY(nsb_nonlinear_forward_map(q_slot, f_slot));                                        // :90
    calls_counter += nsteps;
    NSB_TRY(nsb_vec_norm(f_slot, &residual));
    residual *= residual;                                                                      // :99
    if (hist) hist[it - 1] = residual;
    if (residual < tol) break;                                                                 // :109
    int calls = 0;
    double gres = 0;
    NSB_TRY(nsb_ts_gmres(NSB_NEWTON, f_slot, dq_slot, first_slot, work_slot, maxiter_gmres, k_dim, tol, &calls, &gres));   // :117
    calls_counter += (long long)calls * nsteps;
    NSB_TRY(nsb_vec_sub2(q_slot, dq_slot));                                                    // :122
  }
  if (newton_iters) *newton_iters = it;
  if (residual_out) *residual_out = residual;
  if (calls_out) *calls_out = calls_counter;
  return (residual < tol) ? 0 : 3;
}
