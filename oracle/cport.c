/* ORACLE -- test / measurement infrastructure, NOT product code.  Only tests/, bench.py's cpu_baseline / --impl reference legs
 * and __graft_entry__ may build or call this file (see oracle/ops.py header).
 *
 * C / OpenMP restatement of the 3-D linearised Navier-Stokes step of oracle/stepper.py (PCG mode): the same operators as
 * oracle/ops.py (axhelm, opgradt, opdiv, dssum, dealiased advection) and the same solver loops (Jacobi-PCG Helmholtz,
 * pressure PCG with the Jacobi or the three-level preconditioner of oracle/pmg.py), threaded over elements.  It exists so
 * that the CPU baseline next to the GPU numbers runs on ALL host cores at compiled-code speed; tests/test_oracle_cport.py
 * pins every routine to the numpy oracle, which in turn is pinned to the reference's shipped fixtures.
 * Reference routines restated ([UPSTREAM] Nek5000, reached from core/matvec.f:222): hmholtz.f axhelm/cggo, navier1.f
 * opgradt/opdiv/cdabdtp, dssum.f, convect.f convect_new, perturb.f advabp.
 *
 * Layout: element-major, a(i,j,k,e) with i fastest (Nek order), float64; N = lx1, L = lx2, M = lxd.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* thread control (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline wants every core) */
void c_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int c_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define MAXP 12
#define MAXP3 (MAXP * MAXP * MAXP)

/* One 1-D operator along `axis` of an (n0,n1,n2) array (n0 fastest): out(.., o, ..) = sum_l Aop(o,l) in(.., l, ..), o < no.
 * tr == 0: A is (no x ni) row-major, Aop(o,l) = A[o*ni+l]; tr != 0: A is (ni x no) row-major and applied transposed,
 * Aop(o,l) = A[l*no+o].  Loop orders keep the innermost loop contiguous so that gcc vectorises them. */
static void apply_axis(const double* restrict in, double* restrict out, const double* restrict A, int tr, int axis, int n0, int n1,
                       int n2, int no) {
  if (axis == 0) {
    const int nr = n1 * n2;
    for (int r = 0; r < nr; ++r) {
      const double* pi = in + r * n0;
      double* po = out + r * no;
      if (tr) {
        for (int o = 0; o < no; ++o) po[o] = 0.0;
        for (int l = 0; l < n0; ++l) {
          const double v = pi[l];
          const double* a = A + l * no;
          for (int o = 0; o < no; ++o) po[o] += a[o] * v;
        }
      } else {
        for (int o = 0; o < no; ++o) {
          const double* a = A + o * n0;
          double s = 0.0;
          for (int l = 0; l < n0; ++l) s += a[l] * pi[l];
          po[o] = s;
        }
      }
    }
  } else {
    /* axis 1: n2 slabs of (ni = n1) x (m = n0); axis 2: one slab of (ni = n2) x (m = n0*n1) */
    const int ni = axis == 1 ? n1 : n2, m = axis == 1 ? n0 : n0 * n1, nslab = axis == 1 ? n2 : 1;
    for (int b = 0; b < nslab; ++b) {
      const double* pi = in + b * ni * m;
      double* po = out + b * no * m;
      for (int o = 0; o < no; ++o) {
        double* row = po + o * m;
        for (int q = 0; q < m; ++q) row[q] = 0.0;
        for (int l = 0; l < ni; ++l) {
          const double a = tr ? A[l * no + o] : A[o * ni + l];
          const double* src = pi + l * m;
          for (int q = 0; q < m; ++q) row[q] += a * src[q];
        }
      }
    }
  }
}

/* out = (A2 x A1 x A0) in, all matrices (no x ni) (tr = 0) or their transposes (tr = 1: matrices are (ni x no), out = A^T in) */
static void tens3(const double* in, double* out, const double* A0, const double* A1, const double* A2, int tr, int ni, int no) {
  double t1[MAXP3], t2[MAXP3];
  apply_axis(in, t1, A0, tr, 0, ni, ni, ni, no);
  apply_axis(t1, t2, A1, tr, 1, no, ni, ni, no);
  apply_axis(t2, out, A2, tr, 2, no, no, ni, no);
}

/* ------------------------------------------------------------------------------------------------ axhelm
 * w_f = (h1 A + h2 B) u_f, nf fields (stride n).  G6: 11,22,33,12,13,23 (each n).  mode 1: w = b + w - H u. */
void c_axhelm(int nel, int N, const double* D, const double* G6, const double* bm1, double h1, double h2, int nf,
              const double* u, double* w, const double* b, int mode, long long n) {
  const int np = N * N * N;
  double Dt[MAXP * MAXP];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) Dt[i * N + j] = D[j * N + i];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nel; ++e) {
    const long long e0 = (long long)e * np;
    double ur[MAXP3], us[MAXP3], ut[MAXP3], t[MAXP3], acc[MAXP3];
    for (int f = 0; f < nf; ++f) {
      const double* uf = u + f * n + e0;
      apply_axis(uf, ur, D, 0, 0, N, N, N, N);
      apply_axis(uf, us, D, 0, 1, N, N, N, N);
      apply_axis(uf, ut, D, 0, 2, N, N, N, N);
      for (int p = 0; p < np; ++p) {
        const double g0 = G6[e0 + p], g1 = G6[n + e0 + p], g2 = G6[2 * n + e0 + p], g3 = G6[3 * n + e0 + p],
                     g4 = G6[4 * n + e0 + p], g5 = G6[5 * n + e0 + p];
        const double a = ur[p], bb = us[p], c = ut[p];
        ur[p] = g0 * a + g3 * bb + g4 * c;
        us[p] = g3 * a + g1 * bb + g5 * c;
        ut[p] = g4 * a + g5 * bb + g2 * c;
      }
      apply_axis(ur, acc, Dt, 0, 0, N, N, N, N);
      apply_axis(us, t, Dt, 0, 1, N, N, N, N);
      for (int p = 0; p < np; ++p) acc[p] += t[p];
      apply_axis(ut, t, Dt, 0, 2, N, N, N, N);
      double* wf = w + f * n + e0;
      for (int p = 0; p < np; ++p) {
        const double hv = h1 * (acc[p] + t[p]) + h2 * bm1[e0 + p] * uf[p];
        wf[p] = mode ? b[f * n + e0 + p] + wf[p] - hv : hv;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------ opgradt / opdiv
 * RW2[(i*3+c)*n2 + .] = w2 * J dr_i/dx_c on mesh 2.  w_c = sum_i T_i^T (RW2[i][c] p);  q = sum_c sum_i RW2[i][c] (T_i u_c),
 * T_i = D12 along direction i, J12 along the others.  scale (3 arrays of n, may be NULL) multiplies u_c before the divergence. */
void c_gradt(int nel, int N, int L, const double* J12, const double* D12, const double* RW2, const double* p, double* w,
             long long n, long long n2) {
  const int np = N * N * N, np2 = L * L * L;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nel; ++e) {
    const long long e1 = (long long)e * np, e2 = (long long)e * np2;
    double q[MAXP3], t[MAXP3];
    for (int c = 0; c < 3; ++c) {
      double* wc = w + c * n + e1;
      for (int i = 0; i < 3; ++i) {
        const double* rw = RW2 + (long long)(i * 3 + c) * n2 + e2;
        for (int k = 0; k < np2; ++k) q[k] = rw[k] * p[e2 + k];
        tens3(q, t, i == 0 ? D12 : J12, i == 1 ? D12 : J12, i == 2 ? D12 : J12, 1, L, N);
        if (i == 0) for (int k = 0; k < np; ++k) wc[k] = t[k];
        else for (int k = 0; k < np; ++k) wc[k] += t[k];
      }
    }
  }
}

void c_div(int nel, int N, int L, const double* J12, const double* D12, const double* RW2, const double* u, const double* scale,
           double* q, double sign, long long n, long long n2) {
  const int np = N * N * N, np2 = L * L * L;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nel; ++e) {
    const long long e1 = (long long)e * np, e2 = (long long)e * np2;
    double us[MAXP3], t[MAXP3], acc[MAXP3];
    for (int k = 0; k < np2; ++k) acc[k] = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double* uc = u + c * n + e1;
      if (scale) for (int k = 0; k < np; ++k) us[k] = uc[k] * scale[c * n + e1 + k];
      else for (int k = 0; k < np; ++k) us[k] = uc[k];
      for (int i = 0; i < 3; ++i) {
        const double* rw = RW2 + (long long)(i * 3 + c) * n2 + e2;
        tens3(us, t, i == 0 ? D12 : J12, i == 1 ? D12 : J12, i == 2 ? D12 : J12, 0, N, L);
        for (int k = 0; k < np2; ++k) acc[k] += rw[k] * t[k];
      }
    }
    for (int k = 0; k < np2; ++k) q[e2 + k] = sign * acc[k];
  }
}

/* ------------------------------------------------------------------------------------------------ dssum over CSR segments */
void c_dssum(int nseg, const int* seg_off, const int* seg_idx, double* u, int nf, long long stride) {
#pragma omp parallel for schedule(static)
  for (int s = 0; s < nseg; ++s) {
    for (int f = 0; f < nf; ++f) {
      double acc = 0.0;
      for (int j = seg_off[s]; j < seg_off[s + 1]; ++j) acc += u[f * stride + seg_idx[j]];
      for (int j = seg_off[s]; j < seg_off[s + 1]; ++j) u[f * stride + seg_idx[j]] = acc;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ pressure preconditioner
 * (oracle/pmg.py): z = FDM(r) + P diag^-1 P^T r + Pa A2^-1 Pa^T r */
typedef struct {
  int nv, nagg;
  const double* S;       /* [nel][3][L*L] row = node, column = mode */
  const double* deninv;  /* [nel][L^3] */
  const double* phi;     /* [8][L^3] */
  const int* vid;        /* [nel][8] */
  const int* voff;       /* [nv+1] */
  const int* vent;       /* [nel*8] */
  const double* d1inv;   /* [nv] */
  const int* agg;        /* [nel] */
  const double* A2inv;   /* [nagg][nagg] */
  double *rc, *xv, *ra, *x2;   /* work: [nel*8], [nv], [nagg], [nagg] */
} cpmg_t;

static void pmg_apply(int nel, int L, const cpmg_t* m, const double* r, double* z) {
  const int np2 = L * L * L;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nel; ++e) {
    const double* re = r + (long long)e * np2;
    for (int k = 0; k < 8; ++k) {
      double s = 0.0;
      for (int p = 0; p < np2; ++p) s += m->phi[k * np2 + p] * re[p];
      m->rc[e * 8 + k] = s;
    }
  }
#pragma omp parallel for schedule(static)
  for (int v = 0; v < m->nv; ++v) {
    double s = 0.0;
    for (int j = m->voff[v]; j < m->voff[v + 1]; ++j) s += m->rc[m->vent[j]];
    m->xv[v] = m->d1inv[v] * s;
  }
  for (int a = 0; a < m->nagg; ++a) m->ra[a] = 0.0;
  for (int e = 0; e < nel; ++e) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += m->rc[e * 8 + k];
    m->ra[m->agg[e]] += s;
  }
#pragma omp parallel for schedule(static)
  for (int a = 0; a < m->nagg; ++a) {
    double s = 0.0;
    for (int b = 0; b < m->nagg; ++b) s += m->A2inv[(long long)a * m->nagg + b] * m->ra[b];
    m->x2[a] = s;
  }
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nel; ++e) {
    const double* re = r + (long long)e * np2;
    const double* S = m->S + (long long)e * 3 * L * L;
    double t[MAXP3], o[MAXP3];
    tens3(re, t, S, S + L * L, S + 2 * L * L, 1, L, L);                 /* S^T along every direction */
    for (int p = 0; p < np2; ++p) t[p] *= m->deninv[(long long)e * np2 + p];
    tens3(t, o, S, S + L * L, S + 2 * L * L, 0, L, L);
    double xk[8];
    for (int k = 0; k < 8; ++k) xk[k] = m->xv[m->vid[e * 8 + k]];
    const double c2 = m->x2[m->agg[e]];
    for (int p = 0; p < np2; ++p) {
      double v = o[p] + c2;
      for (int k = 0; k < 8; ++k) v += m->phi[k * np2 + p] * xk[k];
      z[(long long)e * np2 + p] = v;
    }
  }
}

void c_pmg_apply(int nel, int L, const cpmg_t* m, const double* r, double* z) { pmg_apply(nel, L, m, r, z); }

/* ------------------------------------------------------------------------------------------------ pressure PCG
 * E = D (mbinv QQ^T) D^T.  g: right-hand side (destroyed: becomes the residual), x: solution.  Stops on
 * sqrt(sum r^2/bm2 / vol2) <= tol or after maxit iterations; returns the iteration count.  pm == NULL: Jacobi (dinvE). */
typedef struct {
  int nel, N, L, nseg;
  const double *J12, *D12, *RW2, *mbinv /* [3][n] */, *bm2inv, *dinvE;
  const int *seg_off, *seg_idx;
  double vol2;
} cpres_t;

static void apply_E(const cpres_t* c, const double* p, double* w /* [3][n] work */, double* ep, long long n, long long n2) {
  c_gradt(c->nel, c->N, c->L, c->J12, c->D12, c->RW2, p, w, n, n2);
  c_dssum(c->nseg, c->seg_off, c->seg_idx, w, 3, n);
  c_div(c->nel, c->N, c->L, c->J12, c->D12, c->RW2, w, c->mbinv, ep, 1.0, n, n2);
}
void c_apply_E(const cpres_t* c, const double* p, double* w, double* ep) {
  const long long n = (long long)c->nel * c->N * c->N * c->N, n2 = (long long)c->nel * c->L * c->L * c->L;
  apply_E(c, p, w, ep, n, n2);
}

int c_pressure_pcg(const cpres_t* c, const cpmg_t* pm, double* g, double* x, double* pd, double* z, double* ep, double* w,
                   double tol, int maxit) {
  const long long n = (long long)c->nel * c->N * c->N * c->N, n2 = (long long)c->nel * c->L * c->L * c->L;
  double rtz1 = 1.0;
  int it = 0;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n2; ++i) { x[i] = 0.0; pd[i] = 0.0; }
  for (;;) {
    if (pm) pmg_apply(c->nel, c->L, pm, g, z);
    double rtz = 0.0, rn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rtz, rn)
    for (long long i = 0; i < n2; ++i) {
      if (!pm) z[i] = c->dinvE[i] * g[i];
      rtz += z[i] * g[i];
      rn += g[i] * g[i] * c->bm2inv[i];
    }
    const double rtz2 = rtz1;
    rtz1 = rtz;
    if (sqrt(rn / c->vol2) <= tol || it >= maxit) break;
    const double beta = it == 0 ? 0.0 : rtz1 / rtz2;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n2; ++i) pd[i] = z[i] + beta * pd[i];
    apply_E(c, pd, w, ep, n, n2);
    double rho = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rho)
    for (long long i = 0; i < n2; ++i) rho += ep[i] * pd[i];
    const double alpha = rtz1 / rho;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n2; ++i) { x[i] += alpha * pd[i]; g[i] -= alpha * ep[i]; }
    ++it;
  }
  return it;
}

/* ------------------------------------------------------------------------------------------------ Helmholtz Jacobi-PCG
 * one component: H x = r (r assembled and masked, destroyed).  Stops on sqrt(sum r^2 mult binv / vol) <= tol or maxit. */
typedef struct {
  int nel, N, nseg;
  const double *D, *G6, *bm1, *dinv, *mult, *binv;
  const int *seg_off, *seg_idx;
  double vol;
} chelm_t;

int c_helmholtz_pcg(const chelm_t* c, const double* mask, double h1, double h2, double* r, double* x, double* pd, double* w,
                    double tol, int maxit) {
  const long long n = (long long)c->nel * c->N * c->N * c->N;
  double rtz1 = 1.0;
  int it = 0;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) { x[i] = 0.0; pd[i] = 0.0; }
  for (;;) {
    double rtz = 0.0, rb = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rtz, rb)
    for (long long i = 0; i < n; ++i) {
      const double zi = c->dinv[i] * r[i] * mask[i];
      rtz += zi * r[i] * c->mult[i];
      rb += r[i] * r[i] * c->mult[i] * c->binv[i];
    }
    const double rtz2 = rtz1;
    rtz1 = rtz;
    if (sqrt(fmax(rb, 0.0) / c->vol) <= tol || it >= maxit) break;
    const double beta = it == 0 ? 0.0 : rtz1 / rtz2;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) pd[i] = c->dinv[i] * r[i] * mask[i] + beta * pd[i];
    c_axhelm(c->nel, c->N, c->D, c->G6, c->bm1, h1, h2, 1, pd, w, NULL, 0, n);
    c_dssum(c->nseg, c->seg_off, c->seg_idx, w, 1, n);
    double rho = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rho)
    for (long long i = 0; i < n; ++i) { w[i] *= mask[i]; rho += w[i] * pd[i] * c->mult[i]; }
    const double alpha = rtz1 / rho;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) { x[i] += alpha * pd[i]; r[i] -= alpha * w[i]; }
    ++it;
  }
  return it;
}

/* ------------------------------------------------------------------------------------------------ dealiased advection (direct)
 * f_k = -( B[(u'.grad)U_k + (U.grad)u'_k] ) - bm1*spng*u'_k.   Rd[(i*3+c)*nd + .] = wd * J dr_i/dx_c on GL(lxd). */
void c_advab_direct(int nel, int N, int M, const double* Jd, const double* Dd, const double* Rd, const double* bm1,
                    const double* spng, const double* up, const double* ub, double* f, long long n, long long nd) {
  const int np = N * N * N, npd = M * M * M;
#pragma omp parallel
  {
    double* buf = (double*)malloc(sizeof(double) * (size_t)npd * 16);
    double *Fp = buf, *Fb = buf + 3 * npd, *crp = buf + 6 * npd, *crb = buf + 9 * npd, *g0 = buf + 12 * npd, *g1 = buf + 13 * npd,
           *g2 = buf + 14 * npd, *acc = buf + 15 * npd;
#pragma omp for schedule(static)
    for (int e = 0; e < nel; ++e) {
      const long long e1 = (long long)e * np, ed = (long long)e * npd;
      for (int c = 0; c < 3; ++c) {
        tens3(up + c * n + e1, Fp + c * npd, Jd, Jd, Jd, 0, N, M);
        tens3(ub + c * n + e1, Fb + c * npd, Jd, Jd, Jd, 0, N, M);
      }
      for (int i = 0; i < 3; ++i)
        for (int q = 0; q < npd; ++q) {
          double sp = 0.0, sb = 0.0;
          for (int c = 0; c < 3; ++c) {
            const double r = Rd[(long long)(i * 3 + c) * nd + ed + q];
            sp += r * Fp[c * npd + q];
            sb += r * Fb[c * npd + q];
          }
          crp[i * npd + q] = sp;
          crb[i * npd + q] = sb;
        }
      for (int k = 0; k < 3; ++k) {
        for (int q = 0; q < npd; ++q) acc[q] = 0.0;
        for (int pass = 0; pass < 2; ++pass) {                 /* pass 0: (u'.grad) U_k ; pass 1: (U.grad) u'_k */
          const double* src = (pass == 0 ? ub : up) + k * n + e1;
          const double* cr = pass == 0 ? crp : crb;
          tens3(src, g0, Dd, Jd, Jd, 0, N, M);
          tens3(src, g1, Jd, Dd, Jd, 0, N, M);
          tens3(src, g2, Jd, Jd, Dd, 0, N, M);
          for (int q = 0; q < npd; ++q) acc[q] += cr[q] * g0[q] + cr[npd + q] * g1[q] + cr[2 * npd + q] * g2[q];
        }
        double o[MAXP3];
        tens3(acc, o, Jd, Jd, Jd, 1, M, N);
        for (int p = 0; p < np; ++p) {
          double v = -o[p];
          if (spng) v -= bm1[e1 + p] * spng[e1 + p] * up[k * n + e1 + p];
          f[k * n + e1 + p] = v;
        }
      }
    }
    free(buf);
  }
}

/* ------------------------------------------------------------------------------------------------ pointwise pieces of the step */
/* b_c = sum_j ab_j f_j,c + (bm1/dt) sum_j bd_(j+1) u_j,c   (k terms; pointers to the history rings) */
void c_make_rhs(long long n, int k, const double* ab, const double* bd, double dt, const double* bm1, const double* f0,
                const double* f1, const double* f2, const double* u0, const double* u1, const double* u2, double* b) {
  const double* fs[3] = {f0, f1, f2};
  const double* us[3] = {u0, u1, u2};
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < 3 * n; ++i) {
    double s = 0.0, h = 0.0;
    for (int j = 0; j < k; ++j) { s += ab[j] * fs[j][i]; h += bd[j + 1] * us[j][i]; }
    b[i] = s + bm1[i % n] * h / dt;
  }
}
/* r_c = mask_c * r_c (after dssum) */
void c_mask3(long long n, const double* mask, double* r) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < 3 * n; ++i) r[i] *= mask[i];
}
/* unew = u + du + mbinv * w ; pnew = pt + h2 * phi */
void c_final_update(long long n, long long n2, const double* u, const double* du, const double* mbinv, const double* w, double* unew,
                    const double* pt, const double* phi, double h2, double* pnew) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < 3 * n; ++i) unew[i] = u[i] + du[i] + mbinv[i] * w[i];
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n2; ++i) pnew[i] = pt[i] + h2 * phi[i];
}
void c_add3(long long n, const double* a, const double* b, double* out) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) out[i] = a[i] + b[i];
}
