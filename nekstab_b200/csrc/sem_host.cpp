// sem_host.cpp -- Gauss-Lobatto-Legendre / Gauss-Legendre quadrature and the small dense matrices of the
// spectral-element method, computed once on the host in long double and rounded to double.
// Stands in for [UPSTREAM Nek5000 speclib.f: zwgll, zwgl, dgll, igllm] that `nek_init` runs before nekStab's
// first matvec (core/matvec.f:64).  Cross-checked against numpy.polynomial in tests/test_sem.py.
#include <cmath>
#include <vector>

#include "nsb_internal.h"

typedef long double ld;

static void legendre(int n, ld x, ld* p, ld* dp) {
  ld p0 = 1, p1 = x, d0 = 0, d1 = 1;
  if (n == 0) { *p = 1; *dp = 0; return; }
  for (int k = 1; k < n; ++k) {
    ld p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1);
    ld d2 = d0 + (2 * k + 1) * p1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  *p = p1; *dp = d1;
}

void sem_zwgll(int n, double* z, double* w) {
  const int N = n - 1;
  std::vector<ld> x(n);
  const ld pi = acosl(-1.0L);
  for (int i = 0; i < n; ++i) x[i] = -cosl(pi * i / N);
  for (int i = 1; i < N; ++i) {
    for (int it = 0; it < 100; ++it) {
      ld p, dp;
      legendre(N, x[i], &p, &dp);
      ld ddp = (2 * x[i] * dp - (ld)N * (N + 1) * p) / (1 - x[i] * x[i]);
      ld dx = dp / ddp;
      x[i] -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
  }
  x[0] = -1; x[N] = 1;
  for (int i = 0; i < n; ++i) {
    ld xs = 0.5L * (x[i] - x[N - i]);
    ld p, dp;
    legendre(N, xs, &p, &dp);
    z[i] = (double)xs;
    w[i] = (double)(2.0L / ((ld)N * (N + 1) * p * p));
  }
}

void sem_zwgl(int n, double* z, double* w) {
  std::vector<ld> x(n);
  const ld pi = acosl(-1.0L);
  for (int i = 0; i < n; ++i) {
    x[i] = -cosl(pi * (i + 0.75L) / (n + 0.5L));
    for (int it = 0; it < 100; ++it) {
      ld p, dp;
      legendre(n, x[i], &p, &dp);
      ld dx = p / dp;
      x[i] -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
  }
  for (int i = 0; i < n; ++i) {
    ld xs = 0.5L * (x[i] - x[n - 1 - i]);
    ld p, dp;
    legendre(n, xs, &p, &dp);
    z[i] = (double)xs;
    w[i] = (double)(2.0L / ((1 - xs * xs) * dp * dp));
  }
}

static void bary_weights(int n, const double* x, std::vector<ld>& bw) {
  bw.assign(n, 1);
  for (int l = 0; l < n; ++l) {
    ld p = 1;
    for (int m = 0; m < n; ++m)
      if (m != l) p *= ((ld)x[l] - (ld)x[m]);
    bw[l] = 1 / p;
  }
}

void sem_deriv(int n, const double* x, double* D) {
  std::vector<ld> bw;
  bary_weights(n, x, bw);
  for (int i = 0; i < n; ++i) {
    ld s = 0;
    for (int l = 0; l < n; ++l) {
      if (l == i) continue;
      ld v = (bw[l] / bw[i]) / ((ld)x[i] - (ld)x[l]);
      D[i * n + l] = (double)v;
      s += v;
    }
    D[i * n + i] = (double)(-s);
  }
}

void sem_interp(int nto, const double* xto, int nfrom, const double* xfrom, double* J) {
  std::vector<ld> bw;
  bary_weights(nfrom, xfrom, bw);
  for (int i = 0; i < nto; ++i) {
    int hit = -1;
    for (int l = 0; l < nfrom; ++l)
      if (fabs(xto[i] - xfrom[l]) < 1e-15) hit = l;
    if (hit >= 0) {
      for (int l = 0; l < nfrom; ++l) J[i * nfrom + l] = (l == hit) ? 1.0 : 0.0;
      continue;
    }
    ld s = 0;
    std::vector<ld> t(nfrom);
    for (int l = 0; l < nfrom; ++l) {
      t[l] = bw[l] / ((ld)xto[i] - (ld)xfrom[l]);
      s += t[l];
    }
    for (int l = 0; l < nfrom; ++l) J[i * nfrom + l] = (double)(t[l] / s);
  }
}

static void matmul(int m, int k, int n, const double* A, const double* B, double* C) {  // C(m,n) = A(m,k) B(k,n)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      ld s = 0;
      for (int l = 0; l < k; ++l) s += (ld)A[i * k + l] * (ld)B[l * n + j];
      C[i * n + j] = (double)s;
    }
}

static void transpose(int m, int n, const double* A, double* At) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

void sem_build_constmats(int lx1, int lx2, int lxd, ConstMats* cm) {
  memset(cm, 0, sizeof(*cm));
  double z2[12], zd[18];
  sem_zwgll(lx1, cm->z1, cm->w1);
  sem_zwgl(lx2, z2, cm->w2);
  for (int i = 0; i < lx2; ++i) cm->hat1[i] = 0.5 * (1.0 + z2[i]);
  sem_zwgl(lxd, zd, cm->wd);
  sem_deriv(lx1, cm->z1, cm->D);
  transpose(lx1, lx1, cm->D, cm->Dt);
  sem_interp(lx2, z2, lx1, cm->z1, cm->J12);
  matmul(lx2, lx1, lx1, cm->J12, cm->D, cm->D12);
  transpose(lx2, lx1, cm->J12, cm->J12t);
  transpose(lx2, lx1, cm->D12, cm->D12t);
  sem_interp(lxd, zd, lx1, cm->z1, cm->Jd);
  matmul(lxd, lx1, lx1, cm->Jd, cm->D, cm->Dd);
  transpose(lxd, lx1, cm->Jd, cm->Jdt);
  transpose(lxd, lx1, cm->Dd, cm->Ddt);
}
