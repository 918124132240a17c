// elem_common.cuh -- building blocks of the element-local tensor-product kernels.
//
// Layout: one CTA works on one spectral element.  Element data sit in shared memory as (NK,NJ,NI) arrays
// whose rows are padded to an ODD number of doubles so that 64-bit column accesses along any axis are
// bank-conflict free (2 wavefronts per warp, the minimum for 8-byte words).  A 1-D operator (lx x lx
// derivative / interpolation matrix) is applied along one axis by giving each thread whole *columns*:
// the thread loads the NL inputs of a column into registers and produces the NO outputs with NO*NL DFMAs
// whose matrix operand comes straight from the constant bank (compile-time index => `DFMA R, R, c[..]`),
// so shared memory sees NL+NO accesses per column instead of 2*NO*NL.
#pragma once
#include "nsb_internal.h"

static __constant__ ConstMats cm;

template <int N>
struct OddPitch {
  static constexpr int v = (N % 2 == 0) ? N + 1 : N;
};

// shape helper for a pitched (NK,NJ,NI) shared-memory array
template <int NK, int NJ, int NI>
struct Shp {
  static constexpr int PI = OddPitch<NI>::v;
  static constexpr int size = NK * NJ * PI;
  static constexpr int npts = NK * NJ * NI;
  __device__ __forceinline__ static int at(int k, int j, int i) { return (k * NJ + j) * PI + i; }
  // linear (unpitched, i fastest) point index -> pitched offset
  __device__ __forceinline__ static int lin(int p) {
    int i = p % NI;
    int r = p / NI;
    return r * PI + i;
  }
};

// out(.., o, ..) (+)= sum_l M[o*NL+l] * in(.., l, ..) along axis AX (0 = i fastest, 1 = j, 2 = k).
// (NK,NJ,NI) are the INPUT dims; the output has NO along AX.  Columns are dealt to threads with the
// fastest remaining axis varying fastest across lanes.  Callers must __syncthreads() between dependent stages;
// two calls with identical (AX, input dims, NO) map columns to threads identically, so ACC=true after
// ACC=false on the same `out` needs no barrier in between.
template <int AX, int NO, int NL, int NK, int NJ, int NI, bool ACC>
__device__ __forceinline__ void contract(const double* __restrict__ in, double* __restrict__ out,
                                         const double* __restrict__ M, int tid, int nthr) {
  constexpr int PIi = OddPitch<NI>::v;
  if constexpr (AX == 0) {
    static_assert(NI == NL, "axis length");
    constexpr int PIo = OddPitch<NO>::v;
    constexpr int ncol = NK * NJ;
    for (int c = tid; c < ncol; c += nthr) {
      const double* pi = in + c * PIi;
      double* po = out + c * PIo;
      double v[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) v[l] = pi[l];
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        double s = ACC ? po[o] : 0.0;
#pragma unroll
        for (int l = 0; l < NL; ++l) s = fma(M[o * NL + l], v[l], s);
        po[o] = s;
      }
    }
  } else if constexpr (AX == 1) {
    static_assert(NJ == NL, "axis length");
    constexpr int ncol = NK * NI;
    for (int c = tid; c < ncol; c += nthr) {
      const int i = c % NI, k = c / NI;
      const double* pi = in + k * NJ * PIi + i;
      double* po = out + k * NO * PIi + i;
      double v[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) v[l] = pi[l * PIi];
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        double s = ACC ? po[o * PIi] : 0.0;
#pragma unroll
        for (int l = 0; l < NL; ++l) s = fma(M[o * NL + l], v[l], s);
        po[o * PIi] = s;
      }
    }
  } else {
    static_assert(NK == NL, "axis length");
    constexpr int ncol = NJ * NI;
    constexpr int str = NJ * PIi;
    for (int c = tid; c < ncol; c += nthr) {
      const int i = c % NI, j = c / NI;
      const double* pi = in + j * PIi + i;
      double* po = out + j * PIi + i;
      double v[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) v[l] = pi[l * str];
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        double s = ACC ? po[o * str] : 0.0;
#pragma unroll
        for (int l = 0; l < NL; ++l) s = fma(M[o * NL + l], v[l], s);
        po[o * str] = s;
      }
    }
  }
}

// "Surface-first" layout of one element's N^3 points (3-D): the 6 N^2 - 12 N + 8 surface points first -- plane k = 0, plane k = N-1,
// then for every interior plane its ring (row j = 0, row j = N-1, then the (i = 0, i = N-1) pairs of the interior rows) -- followed by
// the (N-2)^3 interior points.  The direct-stiffness sum only touches surface points: in the natural layout the i = 0 / i = N-1 points
// of interior rows sit alone in their 32-byte sectors (r1d ncu: k_gs_sum moved 1.53 x its algorithmic bytes); here the surface block is
// contiguous.  Used for the vector that travels  producer kernel -> dssum -> consumer kernel  inside the two CG loops.
template <int N>
struct SurfFirst {
  static constexpr int NS = N * N * N - (N - 2) * (N - 2) * (N - 2);
  static constexpr int B = 4 * N - 4;
  __host__ __device__ static constexpr int ring(int j, int i) {
    return j == 0 ? i : (j == N - 1 ? N + i : 2 * N + 2 * (j - 1) + (i == N - 1 ? 1 : 0));
  }
  __host__ __device__ static constexpr bool surf_col(int j, int i) { return j == 0 || j == N - 1 || i == 0 || i == N - 1; }
  // position of point (k, j, i)
  __host__ __device__ static constexpr int pos(int k, int j, int i) {
    return k == 0 ? j * N + i
                  : (k == N - 1 ? N * N + j * N + i
                                : (surf_col(j, i) ? 2 * N * N + (k - 1) * B + ring(j, i)
                                                  : NS + ((k - 1) * (N - 2) + (j - 1)) * (N - 2) + (i - 1)));
  }
  __host__ __device__ static constexpr int pos_lin(int q) { return pos(q / (N * N), (q / N) % N, q % N); }
  // a k-column (fixed j, i): first interior-plane position and the stride between interior planes
  __host__ __device__ static constexpr int mid0(int j, int i) {
    return surf_col(j, i) ? 2 * N * N + ring(j, i) : NS + (j - 1) * (N - 2) + (i - 1);
  }
  __host__ __device__ static constexpr int mids(int j, int i) { return surf_col(j, i) ? B : (N - 2) * (N - 2); }
};
// run-time N (host set-up, streaming kernels)
__host__ __device__ inline int surf_first_pos(int N, int q) {
  const int i = q % N, j = (q / N) % N, k = q / (N * N);
  const int NS = N * N * N - (N - 2) * (N - 2) * (N - 2), B = 4 * N - 4;
  if (k == 0) return j * N + i;
  if (k == N - 1) return N * N + j * N + i;
  const bool sc = j == 0 || j == N - 1 || i == 0 || i == N - 1;
  if (sc) return 2 * N * N + (k - 1) * B + (j == 0 ? i : (j == N - 1 ? N + i : 2 * N + 2 * (j - 1) + (i == N - 1 ? 1 : 0)));
  return NS + ((k - 1) * (N - 2) + (j - 1)) * (N - 2) + (i - 1);
}

// deterministic block reduction of NV values (sum); result in thread 0 (all threads must call)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sred /* >= NV*32 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sred[q * 32 + wid] = x;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double x = (lane < nw) ? sred[q * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      v[q] = x;
    }
  }
  __syncthreads();
}

// Write this block's partial sums and, if it is the last block to arrive, reduce all partials in block
// order (deterministic) into out[0..NV).  Returns true in ALL threads of the last block.
template <int NV>
__device__ __forceinline__ bool grid_sum_finish(double (&v)[NV], double* part, unsigned* counter, double* out,
                                                double* sred) {
  __shared__ int s_last;
  block_sum<NV>(v, sred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) part[(size_t)blockIdx.x * NV + q] = v[q];
    __threadfence();
    unsigned t = atomicInc(counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0.0;
  // fixed assignment of partials to threads, fixed combination order => deterministic
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] += __ldcg(&part[(size_t)b * NV + q]);
  }
  block_sum<NV>(acc, sred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) out[q] = acc[q];
    __threadfence();
  }
  __syncthreads();
  return true;
}
