// scalar.cu -- dealiased convection of a scalar field (the `ifheat` path: temperature / passive scalar travelling with the Krylov
// vector, core/krylov_subspace.f:13,41-45).
//
//   out = J^T [ (Rd J a) . grad_rst(J phi)  +  (Rd J b) . grad_rst(J psi) ]                 [UPSTREAM convect.f convop / convect_new,
//                                                                                            perturb.f convabp: U.grad(theta') + u'.grad(Theta)]
// with J the GLL(lx1) -> GL(lxd) interpolation, Rd the contravariant metrics times the quadrature weights on the fine mesh (the array the
// velocity advection kernels use) -- i.e. exactly one "component" of the velocity operator of elem_kernels.cu, for a field that is not a
// velocity component.  First generation: one CTA per element, everything in shared memory, run-time polynomial orders; it runs once
// per time step next to ~70 CG iterations, so it is written for clarity.  The scalar's Helmholtz solve, dssum and pointwise updates
// reuse the velocity kernels (stepper.cu scalar_explicit / scalar_solve).
#include "nsb_internal.h"

namespace {

// out (no^D) = (M_{D-1} x .. x M_0) in (ni^D); M_ax is no x ni row-major and acts on axis ax (0 = fastest index).  t1, t2: scratch.
__device__ void tensor_apply(const double* in, double* out, double* t1, double* t2, const double* M0, const double* M1,
                             const double* M2, int D, int ni, int no, int tid, int nthr) {
  if (D == 2) {
    for (int p = tid; p < ni * no; p += nthr) {             // t1[j][o]
      const int o = p % no, j = p / no;
      double s = 0.0;
      for (int l = 0; l < ni; ++l) s = fma(M0[o * ni + l], in[j * ni + l], s);
      t1[p] = s;
    }
    __syncthreads();
    for (int p = tid; p < no * no; p += nthr) {             // out[q][o]
      const int o = p % no, q = p / no;
      double s = 0.0;
      for (int l = 0; l < ni; ++l) s = fma(M1[q * ni + l], t1[l * no + o], s);
      out[p] = s;
    }
    __syncthreads();
    return;
  }
  for (int p = tid; p < ni * ni * no; p += nthr) {          // t1[k][j][o]
    const int o = p % no, r = p / no;                       // r = k * ni + j
    double s = 0.0;
    for (int l = 0; l < ni; ++l) s = fma(M0[o * ni + l], in[r * ni + l], s);
    t1[p] = s;
  }
  __syncthreads();
  for (int p = tid; p < ni * no * no; p += nthr) {          // t2[k][q][o]
    const int o = p % no, q = (p / no) % no, k = p / (no * no);
    double s = 0.0;
    for (int l = 0; l < ni; ++l) s = fma(M1[q * ni + l], t1[(k * ni + l) * no + o], s);
    t2[p] = s;
  }
  __syncthreads();
  for (int p = tid; p < no * no * no; p += nthr) {          // out[m][q][o]
    const int qo = p % (no * no), m = p / (no * no);
    double s = 0.0;
    for (int l = 0; l < ni; ++l) s = fma(M2[m * ni + l], t2[l * no * no + qo], s);
    out[p] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_conv_scalar(const double* __restrict__ a, const double* __restrict__ phi, const double* __restrict__ b,
              const double* __restrict__ psi, const double* __restrict__ Rd, const double* __restrict__ mats,
              double* __restrict__ out, long long n, long long nd, int D, int N, int ND) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  int NP1 = N * N, NPD = ND * ND;
  if (D == 3) { NP1 *= N; NPD *= ND; }
  double* sJ = sm;                       // Jd  [ND][N]
  double* sD = sJ + ND * N;              // Dd  [ND][N]
  double* sJt = sD + ND * N;             // Jdt [N][ND]
  double* co = sJt + ND * N;             // coarse field [NP1]
  double* F = co + NP1;                  // fine field [NPD]
  double* t1 = F + NPD;
  double* t2 = t1 + NPD;
  double* cr = t2 + NPD;                 // contravariant advecting field [D][NPD]
  double* acc = cr + 3 * NPD;            // [NPD]
  for (int q = tid; q < 3 * ND * N; q += nthr) sm[q] = mats[q];
  for (int q = tid; q < NPD; q += nthr) acc[q] = 0.0;
  const long long e1 = (long long)blockIdx.x * NP1, ed = (long long)blockIdx.x * NPD;
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    const double* vel = pass ? b : a;
    const double* sc = pass ? psi : phi;
    if (!vel || !sc) continue;
    for (int q = tid; q < D * NPD; q += nthr) cr[q] = 0.0;
    __syncthreads();
    for (int c = 0; c < D; ++c) {
      for (int q = tid; q < NP1; q += nthr) co[q] = vel[(long long)c * n + e1 + q];
      __syncthreads();
      tensor_apply(co, F, t1, t2, sJ, sJ, sJ, D, N, ND, tid, nthr);
      for (int q = tid; q < NPD; q += nthr) {
        const double v = F[q];
        for (int i = 0; i < D; ++i) cr[i * NPD + q] = fma(Rd[(long long)(i * D + c) * nd + ed + q], v, cr[i * NPD + q]);
      }
      __syncthreads();
    }
    for (int q = tid; q < NP1; q += nthr) co[q] = sc[e1 + q];
    __syncthreads();
    for (int i = 0; i < D; ++i) {
      tensor_apply(co, F, t1, t2, i == 0 ? sD : sJ, i == 1 ? sD : sJ, i == 2 ? sD : sJ, D, N, ND, tid, nthr);
      for (int q = tid; q < NPD; q += nthr) acc[q] = fma(cr[i * NPD + q], F[q], acc[q]);
      __syncthreads();
    }
  }
  tensor_apply(acc, co, t1, t2, sJt, sJt, sJt, D, ND, N, tid, nthr);
  for (int q = tid; q < NP1; q += nthr) out[e1 + q] = co[q];
}

}  // namespace

int sk_setup(Ctx* c) {
  Ctx::Scalar& z = c->scal;
  const int N = c->lx1, ND = c->lxd;
  std::vector<double> m(3 * (size_t)ND * N);
  for (int i = 0; i < ND * N; ++i) { m[i] = c->cm.Jd[i]; m[ND * N + i] = c->cm.Dd[i]; m[2 * ND * N + i] = c->cm.Jdt[i]; }
  if (!z.mats) NSB_CUDA(cudaMalloc((void**)&z.mats, m.size() * sizeof(double)));
  NSB_CUDA(cudaMemcpy(z.mats, m.data(), m.size() * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

int sk_conv(Ctx* c, const double* a, const double* phi, const double* b, const double* psi, double* out) {
  const int D = c->ldim, N = c->lx1, ND = c->lxd;
  long long np1 = (long long)N * N, npd = (long long)ND * ND;
  if (D == 3) { np1 *= N; npd *= ND; }
  const size_t smem = (size_t)(3 * ND * N + np1 + 7 * npd) * sizeof(double);
  if (smem > 220 * 1024) { nsb_set_error("scalar transport: lx1 = %d / lxd = %d needs %zu KB of shared memory (first-generation kernel)", N, ND, smem / 1024); return 1; }
  static size_t attr = 0;
  if (smem > attr) {
    NSB_CUDA(cudaFuncSetAttribute(k_conv_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  k_conv_scalar<<<c->nel, 256, smem, c->stream>>>(a, phi, b, psi, c->Rd, c->scal.mats, out, c->n, c->nd, D, N, ND);
  nsb_count_launch();
  NSB_CUDA(cudaGetLastError());
  return 0;
}
