// ransac_common.cuh -- device helpers shared by the homography (ransac.cu) and fundamental-matrix
// (ransac_f.cu) LO-RANSAC kernels: counter-based RNG, MSAC gain, warp reductions, inlier compaction,
// the stopping rule.  fp64, compiled with --fmad=false.
#pragma once
#include <cmath>

namespace {

constexpr int RS_MAX_B = 4096;
constexpr int LO_REPS = 10;       // RAN_REP, rtools.h:8
constexpr int ILSQ_ITERS = 4;     // rtools.h:9
constexpr double TC = 4.0;        // rtools.h:10
constexpr double MWM = 2.0;       // rtools.h:33: (9/4) in integer arithmetic
constexpr int ITER_SAM = 50;

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned rs_rand(unsigned long long seed, unsigned long long stream, unsigned draw, unsigned range) {
  return (unsigned)(mix64(seed ^ mix64(stream * 0x100000001B3ull + draw)) % range);
}

__device__ __forceinline__ double truncQuad(double eps, double thr) {
  if (thr == 0) return 0;
  if (eps >= thr * 9 / 4) return 0;
  return 1 - (eps / (thr * 9 / 4));
}

__device__ __forceinline__ double det3(const double* A) {
  double r = (A[0] * A[4] * A[8] + A[2] * A[3] * A[7] + A[1] * A[5] * A[6]);
  r -= (A[2] * A[4] * A[6] + A[0] * A[5] * A[7] + A[1] * A[3] * A[8]);
  return r;
}

__device__ __forceinline__ bool inv3(const double* A, double* R) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0 || !isfinite(det)) return false;
  const double id = 1.0 / det;
  R[0] = c0 * id; R[1] = (A[2] * A[7] - A[1] * A[8]) * id; R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c1 * id; R[4] = (A[0] * A[8] - A[2] * A[6]) * id; R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c2 * id; R[7] = (A[1] * A[6] - A[0] * A[7]) * id; R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cross3(double* o, const double* a, const double* b) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// warp-wide: indices j (ascending) with d[j] <= th into idx[]; returns the count  (rtools.c inlidxs)
__device__ int compact_inliers(const double* d, int T, double th, int* idx, int lane) {
  int n = 0;
  for (int base = 0; base < T; base += 32) {
    const int j = base + lane;
    const bool in = j < T && d[j] <= th;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) idx[n + __popc(m & ((1u << lane) - 1))] = j;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// rtools.c:196-224
__device__ int nsamples(int ninl, int ptNum, int samsiz, double conf) {
  double a = 1, b = 1;
  for (int i = 0; i < samsiz; i++) { a *= ninl - i; b *= ptNum - i; }
  a = a / b;
  if (a < 2.2204e-16) return 1000000;
  a = 1 - a;
  if (a < 2.2204e-16) return 1;
  b = log(1 - conf) / log(a);
  if (b > 1000000) return 1000000;
  return (int)ceil(b);
}

// K distinct indices out of T: partial Fisher-Yates over a virtual pool (rtools.c sample()), counter-based draws
template <int K>
__device__ __forceinline__ void draw_sample(unsigned long long seed, unsigned long long stream, int T, int* idx) {
  int pos[K], val[K];
  for (int i = 0; i < K; i++) {
    const int s = (int)rs_rand(seed, stream, i, (unsigned)(T - i)), last = T - i - 1;
    int vs = s, vl = last;
    for (int k = 0; k < i; k++) { if (pos[k] == s) vs = val[k]; if (pos[k] == last) vl = val[k]; }
    idx[i] = vs;
    pos[i] = s; val[i] = vl;
  }
}

}  // namespace
