// ransac_f.cu -- batched LO-RANSAC for a fundamental matrix on the device (seam S4).
//
// Replaces exp_ransacFcustom (degensac/exp_ranF.c:805-1201) as LORANSACFiltering calls it
// (matching.cpp:722: do_lo = 1, Sampson error FDs / exFDs, MSAC score, symmetric check).  Same organisation as
// ransac.cu:
//   k_rf_hyp     one WARP per 7-point sample: counter-based RNG -> 7x9 system (lin_fm, Ftools.c:15-37) ->
//                2-D null space by pivoted Gauss-Jordan (utools.c:97-167) -> cubic det(x f1 + (1-x) f2) = 0
//                (slcm Ftools.c:39-81, rroots3 :200-247) -> up to 3 models -> oriented epipolar constraint
//                (all_ori_valid :430-445) -> Sampson error of all T correspondences (FDs :83-101), lanes
//                striding over T -> MSAC score; the sample's best model is kept
//   k_rf_select / k_rf_lo / k_rf_accept  per batch: best model of the batch, symmetric epipolar check (exp_ranF.c:935-950:
//                more than 0.6*I correspondences within 16*th), local optimisation = LSQ on the 8*th band
//                (__LSQ_BEFORE_LO__) + 10 inner samples of <= 14 inliers x 4 shrinking-threshold weighted LSQ
//                steps (exp_inFranicustom :759-803, exp_iterFcustom :601-757), one warp per inner sample;
//                adaptive stopping nsamples(I+1, T, 7, conf)
// LSQ = normalised 8-point (u2f / u2fw, Ftools.c:299-410): normu -> 9x9 normal matrix -> smallest
// eigenvector (inverse iteration) -> rank-2 projection (singulF :279-297) -> denormF.
// fp64, --fmad=false, fixed-order reductions: reproducible from (u, params.seed).
//
// Deviations from the reference, on purpose:
//  (1)-(3) as in ransac.cu (batched samples, inverse iteration instead of LAPACK dsyev_/ccmath svduv, no
//      inlier-set hashing);
//  (4) matching.cpp:722 passes inlLimit = 0, which with __D3__ (exp_ranF.h:25) makes every LO least-squares
//      step of the reference use a RANDOM SUBSET OF 8 inliers; this implementation uses all inliers of the
//      band (the estimator LO-RANSAC describes) -- models are at least as well supported;
//  (5) the DEGENSAC branch (checksample -> innerH -> rFtH plane-and-parallax, exp_ranF.c:963-1016, DegUtils.c) is
//      implemented with the same batching: plane-and-parallax hypotheses come 2048 at a time and only the best of
//      a batch is refined by innerFH (the reference refines every hypothesis that improves on the previous one);
//      the epipole of Hdetect comes from row cross products instead of ccmath svduv.
#include "common.cuh"
#include "ransac_common.cuh"
#include "ransac_h.cuh"
#include <cmath>
#include <algorithm>

namespace {

constexpr double F_CHECK_COEF = 16.0;   // exp_ranF.c:20
constexpr double F_SYMM_COEF = 0.6;     // exp_ranF.c:21
constexpr double XEPS = 1.9984e-15;     // Ftools.c:410

struct RfState {
  double F[9];            // best model (maxS)
  double J; int I;
  double Fs[9];           // best sample model so far (maxSs / FBest)
  double Js; int Is;
  int max_sam, no_sam, lo_runs, sym_rejects, done, have_sample;
  // DEGENSAC (exp_ranF.c:963-1016): a new best sample whose 7 points contain >= 5 on one plane
  int degen_pending, degen_cnt, Ihmax, pad;
  double Hdeg[9];         // homography of that plane (column-major, image 2 -> image 1)
  double Fdeg[9];         // the sample's model (kept for the I / J bookkeeping after plane-and-parallax)
};
struct FHyp { double F[9]; double J; int I; int flag; int idx[7]; int pad; };   // flag 0 ok, 1 all models OC-rejected, 2 degenerate sample

// Ftools.c:83-101 FDs for one correspondence; den = the Sampson denominator (exFDs weight^-2)
__device__ __forceinline__ double fds(const double* F, const double* u, double* den) {
  const double rxc = F[0] * u[3] + F[3] * u[4] + F[6];
  const double ryc = F[1] * u[3] + F[4] * u[4] + F[7];
  const double rwc = F[2] * u[3] + F[5] * u[4] + F[8];
  const double r = (u[0] * rxc + u[1] * ryc + rwc);
  const double rx = F[0] * u[0] + F[1] * u[1] + F[2];
  const double ry = F[3] * u[0] + F[4] * u[1] + F[5];
  const double w = rxc * rxc + ryc * ryc + rx * rx + ry * ry;
  if (den) *den = w;
  return r * r / w;
}
// Ftools.c:103-124 FDsSym
__device__ __forceinline__ double fds_sym(const double* F, const double* u) {
  const double rxc = F[0] * u[3] + F[3] * u[4] + F[6];
  const double ryc = F[1] * u[3] + F[4] * u[4] + F[7];
  const double rwc = F[2] * u[3] + F[5] * u[4] + F[8];
  const double r = (u[0] * rxc + u[1] * ryc + rwc);
  const double rx = F[0] * u[0] + F[1] * u[1] + F[2];
  const double ry = F[3] * u[0] + F[4] * u[1] + F[5];
  const double a = rxc * rxc + ryc * ryc, b = rx * rx + ry * ry;
  return r * r * (a + b) / (a * b);
}

// utools.c:97-167 nullspace() on a 9x9 row-major system; returns the nullity, null vectors (<= 2) in sol[k*9+..]
__device__ int nullspace9x2(double* m, double* sol) {
  const int n = 9;
  int nopivot[9], pivotc[9], nnp = 0, npv = 0;
  const double tol = 1e-12;
  int i = 0;
  for (int j = 0; j < n; j++) {
    double pivot = i < n ? fabs(m[n * i + j]) : 0.0;
    int mx = i;
    for (int k = i + 1; k < n; k++) { double t = fabs(m[n * k + j]); if (pivot < t) { pivot = t; mx = k; } }
    if (pivot < tol) {
      nopivot[nnp++] = j;
      for (int k = i; k < n; k++) m[n * k + j] = 0;
    } else {
      pivotc[npv++] = j;
      for (int k = j; k < n; k++) { double t = m[i * n + k]; m[i * n + k] = m[mx * n + k]; m[mx * n + k] = t; }
      pivot = m[i * n + j];
      for (int k = j; k < n; k++) m[i * n + k] /= pivot;
      for (int k = 0; k < i; k++) { double p = -m[k * n + j]; for (int l = j; l < n; l++) m[k * n + l] += p * m[i * n + l]; }
      for (int k = i + 1; k < n; k++) { double p = m[k * n + j]; for (int l = j; l < n; l++) m[k * n + l] -= p * m[i * n + l]; }
      i++;
    }
  }
  if (nnp <= 2) {
    for (int k = 0; k < nnp; k++) {
      const int j = nopivot[k];
      for (int l = 0; l < n - nnp; l++) sol[k * n + pivotc[l]] = -m[l * n + j];
      for (int l = 0; l < nnp; l++) sol[k * n + nopivot[l]] = (j == nopivot[l]) ? 1.0 : 0.0;
    }
  }
  return nnp;
}

// Ftools.c:39-81 slcm: p[4] with det(x A + (1-x) B) = 0; REPLACES B by A - B like the reference
__device__ void slcm(const double* A, double* B, double* p) {
#define a11 A[0]
#define a12 A[1]
#define a13 A[2]
#define a21 A[3]
#define a22 A[4]
#define a23 A[5]
#define a31 A[6]
#define a32 A[7]
#define a33 A[8]
#define b11 B[0]
#define b12 B[1]
#define b13 B[2]
#define b21 B[3]
#define b22 B[4]
#define b23 B[5]
#define b31 B[6]
#define b32 B[7]
#define b33 B[8]
  p[0] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
  p[1] = -(a33 * b12 * b21) + a32 * b13 * b21 + a33 * b11 * b22 - a31 * b13 * b22 - a32 * b11 * b23 + a31 * b12 * b23 +
         a23 * b12 * b31 - a22 * b13 * b31 - a13 * b22 * b31 + 3 * b13 * b22 * b31 + a12 * b23 * b31 - 3 * b12 * b23 * b31 -
         a23 * b11 * b32 + a21 * b13 * b32 + a13 * b21 * b32 - 3 * b13 * b21 * b32 - a11 * b23 * b32 + 3 * b11 * b23 * b32 +
         (a22 * b11 - a21 * b12 - a12 * b21 + 3 * b12 * b21 + a11 * b22 - 3 * b11 * b22) * b33;
  p[2] = -(a21 * a33 * b12) + a21 * a32 * b13 + a13 * a32 * b21 - a12 * a33 * b21 + 2 * a33 * b12 * b21 - 2 * a32 * b13 * b21 -
         a13 * a31 * b22 + a11 * a33 * b22 - 2 * a33 * b11 * b22 + 2 * a31 * b13 * b22 + a12 * a31 * b23 - a11 * a32 * b23 +
         2 * a32 * b11 * b23 - 2 * a31 * b12 * b23 + 2 * a13 * b22 * b31 - 3 * b13 * b22 * b31 - 2 * a12 * b23 * b31 +
         3 * b12 * b23 * b31 + a13 * a21 * b32 - 2 * a21 * b13 * b32 - 2 * a13 * b21 * b32 + 3 * b13 * b21 * b32 +
         2 * a11 * b23 * b32 - 3 * b11 * b23 * b32 +
         a23 * (-(a32 * b11) + a31 * b12 + a12 * b31 - 2 * b12 * b31 - a11 * b32 + 2 * b11 * b32) +
         (-(a12 * a21) + 2 * a21 * b12 + 2 * a12 * b21 - 3 * b12 * b21 - 2 * a11 * b22 + 3 * b11 * b22) * b33 +
         a22 * (a33 * b11 - a31 * b13 - a13 * b31 + 2 * b13 * b31 + a11 * b33 - 2 * b11 * b33);
  for (int i = 0; i < 9; i++) B[i] = A[i] - B[i];
  p[3] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
#undef a11
#undef a12
#undef a13
#undef a21
#undef a22
#undef a23
#undef a31
#undef a32
#undef a33
#undef b11
#undef b12
#undef b13
#undef b21
#undef b22
#undef b23
#undef b31
#undef b32
#undef b33
}

// Ftools.c:200-247 rroots3: real roots of po[0] x^3 + po[1] x^2 + po[2] x + po[3]
__device__ int rroots3(const double* po, double* r) {
  const double b = po[1] / po[0], c = po[2] / po[0];
  const double b2 = b * b, bt = b / 3;
  const double p = (3 * c - b2) / 9;
  const double q = ((2 * b2 * b) / 27 - b * c / 3 + po[3] / po[0]) / 2;
  const double D = q * q + p * p * p;
  if (D > 0) {
    const double A = sqrt(D) - q;
    if (A > 0) { const double v = pow(A, 1.0 / 3); r[0] = v - p / v - bt; }
    else { const double v = pow(-A, 1.0 / 3); r[0] = p / v - v - bt; }
    return 1;
  }
  const double e = q > 0 ? 1.0 : -1.0;
  const double R = e * sqrt(-p), _2R = R * 2;
  double cosphi = q / (R * R * R);
  if (cosphi > 1) cosphi = 1; else if (cosphi < -1) cosphi = -1;
  const double phit = acos(cosphi) / 3, pit = 3.14159265358979 / 3;
  r[0] = -_2R * cos(phit) - bt;
  r[1] = _2R * cos(pit - phit) - bt;
  r[2] = _2R * cos(pit + phit) - bt;
  return 3;
}

// Ftools.c:412-445 epipole + getorisig + all_ori_valid over the 7 sample correspondences
__device__ bool all_ori_valid(const double* F, const double* us, const int* idx, int N) {
  double ec[3];
  cross3(ec, F, F + 6);
  bool big = false;
  for (int i = 0; i < 3; i++) if (ec[i] > XEPS || ec[i] < -XEPS) big = true;
  if (!big) cross3(ec, F + 3, F + 6);
  double sig1 = 0;
  for (int i = 0; i < N; i++) {
    const double* u = us + 6 * idx[i];
    const double s1 = F[0] * u[3] + F[3] * u[4] + F[6] * u[5];
    const double s2 = ec[1] * u[2] - ec[2] * u[1];
    const double sig = s1 * s2;
    if (i == 0) sig1 = sig;
    else if (sig1 * sig < 0) return false;
  }
  return true;
}

// warp-wide: Sampson errors under F into d[] (optional), weights 1/sqrt(den) into w[] (optional), MSAC score
__device__ void f_score_all(const double* __restrict__ u, int T, const double* F, double th, double* d, double* w, int lane,
                            int* I, double* J) {
  int ci = 0;
  double cj = 0;
  for (int j = lane; j < T; j += 32) {
    double den;
    const double e = fds(F, u + 6 * j, &den);
    if (d) d[j] = e;
    if (w) w[j] = 1 / sqrt(den);
    if (e <= th) ci++;
    cj += truncQuad(e, th);
  }
  *I = warp_sum_i(ci);
  *J = warp_sum_d(cj);
}

// smallest eigenvector of a symmetric positive semi-definite n x n matrix given by its packed lower triangle
// (row i at i*(i+1)/2): 10 steps of inverse iteration on a ridge-regularised Cholesky factor.  Deterministic.
template <int N>
__device__ bool smallest_eigvec(const double* C, double* x) {
  double L[N * (N + 1) / 2];
  double maxd = 0;
  for (int i = 0; i < N; i++) { const double v = C[i * (i + 1) / 2 + i]; if (v > maxd) maxd = v; }
  const double ridge = 1e-13 * maxd, tiny = 1e-30 * maxd + 1e-300;
  for (int i = 0; i < N; i++)
    for (int j = 0; j <= i; j++) {
      double s = C[i * (i + 1) / 2 + j] + (i == j ? ridge : 0.0);
      for (int k = 0; k < j; k++) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      if (i == j) L[i * (i + 1) / 2 + i] = sqrt(s > tiny ? s : tiny);
      else L[i * (i + 1) / 2 + j] = s / L[j * (j + 1) / 2 + j];
    }
  for (int i = 0; i < N; i++) x[i] = 1.0 + 0.1 * i;
  for (int it = 0; it < 10; it++) {
    for (int i = 0; i < N; i++) {
      double s = x[i];
      for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    for (int i = N - 1; i >= 0; i--) {
      double s = x[i];
      for (int k = i + 1; k < N; k++) s -= L[k * (k + 1) / 2 + i] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    double nrm = 0;
    for (int i = 0; i < N; i++) nrm += x[i] * x[i];
    nrm = sqrt(nrm);
    if (!(nrm > 0) || !isfinite(nrm)) return false;
    for (int i = 0; i < N; i++) x[i] /= nrm;
  }
  return true;
}

// Ftools.c:279-297 singulF: closest rank-2 matrix = F - (F v3) v3^T, v3 = right singular vector of the smallest
// singular value (smallest eigenvector of F^T F)
__device__ void singulF(double* F) {
  double M[6];
  int t = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j <= i; j++, t++) M[t] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
  double v[3];
  if (!smallest_eigvec<3>(M, v)) return;
  for (int r = 0; r < 3; r++) {
    const double fv = F[3 * r] * v[0] + F[3 * r + 1] * v[1] + F[3 * r + 2] * v[2];
    for (int c = 0; c < 3; c++) F[3 * r + c] -= fv * v[c];
  }
}

// warp-wide normalised 8-point least squares (u2f / u2fw, Ftools.c:299-410).  idx[0..n) selects the
// correspondences, w (optional, indexed by correspondence) scales their rows.  n < 8 leaves F untouched.
__device__ void lsq_f(const double* __restrict__ u, const int* idx, int n, const double* w, double* F, int lane) {
  if (n < 8) return;
  // normu (utools.c:7-50)
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int k = lane; k < n; k += 32) { const double* p = u + 6 * idx[k]; s0 += p[0]; s1 += p[1]; s2 += p[3]; s3 += p[4]; }
  const double m1x = warp_sum_d(s0) / n, m1y = warp_sum_d(s1) / n, m2x = warp_sum_d(s2) / n, m2y = warp_sum_d(s3) / n;
  double q1 = 0, q2 = 0;
  for (int k = lane; k < n; k += 32) {
    const double* p = u + 6 * idx[k];
    double a = p[0] - m1x, b = p[1] - m1y;
    q1 += sqrt(a * a + b * b);
    a = p[3] - m2x; b = p[4] - m2y;
    q2 += sqrt(a * a + b * b);
  }
  q1 = warp_sum_d(q1); q2 = warp_sum_d(q2);
  double A1[3] = {q1, m1x, m1y}, A2[3] = {q2, m2x, m2y};
  if (A1[0] != 0) A1[0] = n * sqrt(2.0) / A1[0];
  if (A2[0] != 0) A2[0] = n * sqrt(2.0) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0]; A2[1] *= -A2[0]; A2[2] *= -A2[0];
  // normal matrix of the n x 9 design matrix (lin_fmN + scalmul + cov_mat), packed lower triangle
  double C[45];
#pragma unroll
  for (int i = 0; i < 45; i++) C[i] = 0;
  for (int k = lane; k < n; k += 32) {
    const int j = idx[k];
    const double* p = u + 6 * j;
    const double a[3] = {p[0] * A1[0] + A1[1], p[1] * A1[0] + A1[2], 1.0};
    const double b[3] = {p[3] * A2[0] + A2[1], p[4] * A2[0] + A2[2], 1.0};
    const double ws = w ? w[j] : 1.0;
    double r[9];
#pragma unroll
    for (int kk = 0; kk < 3; kk++)
#pragma unroll
      for (int l = 0; l < 3; l++) r[kk * 3 + l] = a[l] * b[kk] * ws;
    int t = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
      for (int jj = 0; jj <= i; jj++, t++) C[t] += r[i] * r[jj];
  }
#pragma unroll
  for (int i = 0; i < 45; i++) C[i] = warp_sum_d(C[i]);
  double x[9];
  if (!smallest_eigvec<9>(C, x)) return;
  singulF(x);
  // denormF (utools.c:53-70)
  double r = A2[0], xx = A2[1], yy = A2[2];
  x[6] += xx * x[0] + yy * x[3];
  x[7] += xx * x[1] + yy * x[4];
  x[8] += xx * x[2] + yy * x[5];
  x[0] *= r; x[1] *= r; x[2] *= r; x[3] *= r; x[4] *= r; x[5] *= r;
  r = A1[0]; xx = A1[1]; yy = A1[2];
  x[2] += xx * x[0] + yy * x[1];
  x[5] += xx * x[3] + yy * x[4];
  x[8] += xx * x[6] + yy * x[7];
  x[0] *= r; x[3] *= r; x[6] *= r;
  x[1] *= r; x[4] *= r; x[7] *= r;
  for (int i = 0; i < 9; i++) F[i] = x[i];
}

// ---- hypothesis generation + scoring: one warp per 7-point sample -----------------------------------------
__global__ void __launch_bounds__(256)
k_rf_hyp(const double* __restrict__ u, int T, double th, unsigned long long seed, int base, int nhyp, FHyp* __restrict__ out) {
  const int wv = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wv >= nhyp) return;
  int idx[7];
  draw_sample<7>(seed, 0x4600000000000000ull + (unsigned long long)(base + wv), T, idx);
  // 7 x 9 system: entry (k,l) = u2[k] * u1[l]  (lin_fm)
  double M[81], sol[18];
  for (int i = 0; i < 7; i++) {
    const double* s = u + 6 * idx[i];
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) M[9 * i + 3 * k + l] = s[k + 3] * s[l];
  }
  for (int i = 63; i < 81; i++) M[i] = 0.0;
  int flag = 0;
  double bestF[9], bestJ = -1;
  int bestI = 0;
  if (nullspace9x2(M, sol) != 2) flag = 2;
  else {
    double poly[4], roots[3];
    double* f1 = sol; double* f2 = sol + 9;
    slcm(f1, f2, poly);
    int nsol = 0;
    if (poly[0] != 0 && isfinite(poly[0])) nsol = rroots3(poly, roots);
    bool any = false;
    for (int i = 0; i < nsol; i++) {
      double f[9];
      for (int j = 0; j < 9; j++) f[j] = f1[j] * roots[i] + f2[j] * (1 - roots[i]);
      if (!isfinite(f[0]) || !all_ori_valid(f, u, idx, 7)) continue;
      any = true;
      int I; double J;
      f_score_all(u, T, f, th, nullptr, nullptr, lane, &I, &J);
      if (J > bestJ) { bestJ = J; bestI = I; for (int j = 0; j < 9; j++) bestF[j] = f[j]; }
    }
    if (!any) flag = nsol > 0 ? 1 : 2;
  }
  if (lane == 0) {
    FHyp o;
    for (int i = 0; i < 9; i++) o.F[i] = flag == 0 ? bestF[i] : 0.0;
    o.I = flag == 0 ? bestI : 0; o.J = flag == 0 ? bestJ : -1.0; o.flag = flag;
    for (int i = 0; i < 7; i++) o.idx[i] = idx[i];
    o.pad = 0;
    out[wv] = o;
  }
}

// exp_ranF.c:935-950: the model is bad when no more than floor(0.6*I) correspondences are within 16*th symmetric
__device__ bool f_sym_check_ok(const double* __restrict__ u, int T, const double* F, double th, int I, int lane) {
  int c = 0;
  for (int j = lane; j < T; j += 32) if (fds_sym(F, u + 6 * j) <= F_CHECK_COEF * th) c++;
  c = warp_sum_i(c);
  return c > (int)floor(F_SYMM_COEF * I);
}

// exp_iterFcustom (exp_ranF.c:601-757) for one inner sample, warp-wide.  d0 = errors of the start model f.
__device__ void f_lo_iterate(const double* __restrict__ u, int T, double th, double* f, const double* d0, double* d, double* w,
                             int* idx, int lane, int* bestI, double* bestJ, double* Fbest) {
  int mI = 0; double mJ = 0;
  for (int j = lane; j < T; j += 32) { if (d0[j] <= th) mI++; mJ += truncQuad(d0[j], th); }
  mI = warp_sum_i(mI); mJ = warp_sum_d(mJ);
  *bestI = 0; *bestJ = 0;
  if (mI < 8) return;
  for (int i = 0; i < 9; i++) Fbest[i] = f[i];
  int n = compact_inliers(d0, T, th * MWM, idx, lane);
  lsq_f(u, idx, n, nullptr, f, lane);
  double ths = TC * th;
  const double dth = (ths - th) / ILSQ_ITERS;
  for (int it = 0; it < ILSQ_ITERS; it++) {
    int sI; double sJ;
    f_score_all(u, T, f, th, d, w, lane, &sI, &sJ);
    __syncwarp();
    if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Fbest[i] = f[i]; }
    n = compact_inliers(d, T, ths * MWM, idx, lane);
    if (n < 8) { *bestI = mI; *bestJ = mJ; return; }
    lsq_f(u, idx, n, w, f, lane);
    ths -= dth;
  }
  int sI; double sJ;
  f_score_all(u, T, f, th, nullptr, nullptr, lane, &sI, &sJ);
  if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Fbest[i] = f[i]; }
  *bestI = mI; *bestJ = mJ;
}

// ---- DEGENSAC: is the 7-point sample dominated by one plane? (DegUtils.c:42-91 checksample, :93-162 Hdetect) ---------
// Hdetect: the homography compatible with F through 3 correspondences (Hartley & Zisserman, "scene planes and
// homographies"): H = A - e v^T with A = [e]x M^T, e the null vector of M (x2^T M x1 = 0, M[k][l] = F[3k+l]).
// The reference takes e from ccmath's svduv; for the rank-2 M of the 7-point solver that is the unit null vector,
// obtained here from the largest cross product of two rows (the sign of e does not change H up to scale).
__device__ void Hdetect(const double* F, const double* u, const int* idx7, const int* tri, double* H) {
  const double* r0 = F; const double* r1 = F + 3; const double* r2 = F + 6;
  double c01[3], c02[3], c12[3], ec[3];
  cross3(c01, r0, r1); cross3(c02, r0, r2); cross3(c12, r1, r2);
  const double n01 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
  const double n02 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
  const double n12 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
  const double* cb = (n01 >= n02 && n01 >= n12) ? c01 : (n02 >= n12 ? c02 : c12);
  const double nb = sqrt(cb[0] * cb[0] + cb[1] * cb[1] + cb[2] * cb[2]);
  for (int i = 0; i < 3; i++) ec[i] = cb[i] / nb;
  const double Ex[9] = {0, -ec[2], ec[1], ec[2], 0, -ec[0], -ec[1], ec[0], 0};
  double A[9];                                    // A = Ex * M^T
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[3 * i + j] = Ex[3 * i] * F[3 * j] + Ex[3 * i + 1] * F[3 * j + 1] + Ex[3 * i + 2] * F[3 * j + 2];
  double b[3], Mx[9];
  for (int i = 0; i < 3; i++) {
    const double* p = u + 6 * idx7[tri[i]];
    const double x1[3] = {p[0], p[1], p[2]}, x2[3] = {p[3], p[4], p[5]};
    double ax2[3], p1[3], p2[3];
    for (int r = 0; r < 3; r++) ax2[r] = A[3 * r] * x2[0] + A[3 * r + 1] * x2[1] + A[3 * r + 2] * x2[2];
    cross3(p1, x1, ax2);                          // x1 x (A x2)
    for (int r = 0; r < 3; r++) p2[r] = -(Ex[3 * r] * x1[0] + Ex[3 * r + 1] * x1[1] + Ex[3 * r + 2] * x1[2]);   // -[e]x x1
    b[i] = (p1[0] * p2[0] + p1[1] * p2[1] + p1[2] * p2[2]) / (p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
    Mx[3 * i] = x2[0]; Mx[3 * i + 1] = x2[1]; Mx[3 * i + 2] = x2[2];
  }
  double Mi[9];
  const bool ok = inv3(Mx, Mi);
  double v[3];
  for (int r = 0; r < 3; r++) v[r] = Mi[3 * r] * b[0] + Mi[3 * r + 1] * b[1] + Mi[3 * r + 2] * b[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) H[i + j * 3] = A[i * 3 + j] - ec[i] * v[j];   // column-major like the reference
  if (!ok || isnan(H[0]) || isinf(H[0])) { for (int i = 0; i < 9; i++) H[i] = (i % 4 == 0) ? 1.0 : 0.0; }
}

// warp-wide; true when some F-compatible homography explains >= 5 of the 7 sample correspondences (error < th)
__device__ bool checksample(const double* F, const double* __restrict__ u, const int* idx7, double th, double* H, int lane) {
  const int IDXS[5][3] = {{0, 1, 2}, {3, 4, 5}, {0, 1, 6}, {3, 4, 6}, {2, 5, 6}};
  for (int i = 0; i < 5; ++i) {
    Hdetect(F, u, idx7, IDXS[i], H);
    double Ds[7];
    int ord[7];
    for (int j = 0; j < 7; j++) { Ds[j] = sampson(H, u + 6 * idx7[j]); ord[j] = j; }
    for (int a = 0; a < 7; ++a)                  // sortDs (DegUtils.c:164-184)
      for (int b = a + 1; b < 7; ++b)
        if (Ds[b] < Ds[a]) { double t = Ds[b]; Ds[b] = Ds[a]; Ds[a] = t; int q = ord[b]; ord[b] = ord[a]; ord[a] = q; }
    int id5[5];
    for (int j = 0; j < 5; j++) id5[j] = idx7[ord[j]];
    lsq_h(u, id5, 5, H, lane);
    int inl = 0;
    for (int j = 0; j < 7; j++) if (sampson(H, u + 6 * idx7[j]) < th) ++inl;
    if (inl > 4) return true;
  }
  return false;
}

// ---- per-batch update: three launches (select / inner samples on LO_REPS SMs / accept), as in ransac.cu ------
// scratch: per inner sample w: dbuf[w][2T]; then dS[T]; then wbuf[w][T] doubles; ibuf[w][T] ints + inl0[T]
struct FLoShare {
  double f0[9];
  int n0, run_lo, lo_id, pad;
  double loJ[LO_REPS]; int loI[LO_REPS]; double loF[LO_REPS][9];
};
constexpr int RF_NW = 12;

__global__ void __launch_bounds__(384)
k_rf_select(const double* __restrict__ u, int T, double th, int do_sym, int do_degen, FHyp* __restrict__ hyp, int nhyp,
            int force_lo, RfState* st, FLoShare* sh, double* dscr, int* iscr) {
  __shared__ double sJ[12]; __shared__ int sIdx[12];
  __shared__ int again_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* dW = dscr;
  int* iW = iscr;
  double* dS = dscr + (size_t)RF_NW * 2 * T;
  int* inl0 = iscr + (size_t)RF_NW * T;

  // (a)+(b): best model of the batch; a model that would become the best-so-far must pass the symmetric check,
  // otherwise it is discarded and the next best is tried (exp_ranF.c:933-960 `continue`)
  bool new_best_sample = false;   // meaningful in warp 0 only
  for (int attempt = 0; attempt < 8; attempt++) {
    double bj = -1; int bi = -1;
    for (int k = threadIdx.x; k < nhyp; k += blockDim.x)
      if (hyp[k].flag == 0 && hyp[k].J > bj) { bj = hyp[k].J; bi = k; }
    for (int o = 16; o > 0; o >>= 1) {
      double oj = __shfl_xor_sync(0xffffffffu, bj, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (oj > bj || (oj == bj && (bi < 0 || oi < bi)))) { bj = oj; bi = oi; }
    }
    if (lane == 0) { sJ[warp] = bj; sIdx[warp] = bi; }
    __syncthreads();
    bj = -1; bi = -1;
    for (int k = 0; k < nw; k++) if (sIdx[k] >= 0 && (sJ[k] > bj || (sJ[k] == bj && sIdx[k] < bi))) { bj = sJ[k]; bi = sIdx[k]; }
    if (warp == 0) {
      int again = 0;
      if (bi >= 0) {
        const double curJ = st->J, curJs = st->Js;
        const int have = st->have_sample || st->Js > 0;
        __syncwarp();
        double f[9];
        for (int i = 0; i < 9; i++) f[i] = hyp[bi].F[i];
        const int I = hyp[bi].I;
        bool bad = false;
        if (curJ < bj) {
          if (do_sym && !f_sym_check_ok(u, T, f, th, I, lane)) bad = true;
          else if (lane == 0) { for (int i = 0; i < 9; i++) st->F[i] = f[i]; st->J = bj; st->I = I; }
        }
        if (bad) {
          if (lane == 0) { hyp[bi].flag = 3; st->sym_rejects++; }
          again = 1;
        } else if (!have || curJs < bj) {
          // maxSs (exp_ranF.c:961-962); a plane-dominated sample goes to the DEGENSAC branch instead of the LO
          int id7[7];
          for (int i = 0; i < 7; i++) id7[i] = hyp[bi].idx[i];
          double Hd[9];
          if (lane == 0) { st->Js = bj; st->Is = I; }
          if (do_degen && checksample(f, u, id7, 3 * th, Hd, lane)) {
            if (lane == 0) { for (int i = 0; i < 9; i++) { st->Hdeg[i] = Hd[i]; st->Fdeg[i] = f[i]; } st->degen_pending = 1; }
          } else {
            if (lane == 0) { for (int i = 0; i < 9; i++) st->Fs[i] = f[i]; st->have_sample = 1; }
            new_best_sample = true;
          }
        }
      }
      if (lane == 0) again_s = again;
    }
    __syncthreads();
    if (!again_s) break;
  }
  if (warp != 0) return;

  // (b') when to run the local optimisation (exp_ranF.c:1017-1040): a new best sample after ITER_SAM samples,
  // or once when ITER_SAM is reached
  const int no_sam = st->no_sam, lo_runs = st->lo_runs, have = st->have_sample;
  __syncwarp();
  if (lane == 0) sh->run_lo = 0;
  if (st->degen_pending) return;      // the host runs the DEGENSAC kernels for this batch; no LO from a degenerate sample
  bool run_lo = false;
  if (have) {
    if (lo_runs == 0 && no_sam + nhyp >= ITER_SAM) run_lo = true;
    if (new_best_sample && no_sam + nhyp > ITER_SAM) run_lo = true;
  }
  if (force_lo) run_lo = have && lo_runs == 0;
  __syncwarp();
  if (run_lo) {
    // __LSQ_BEFORE_LO__ (exp_ranF.c:1048-1054): LSQ on the TC*th*MWM band of the best sample, inliers at th
    double f[9];
    for (int i = 0; i < 9; i++) f[i] = st->Fs[i];
    int I; double J;
    f_score_all(u, T, f, th, dW, nullptr, lane, &I, &J);
    __syncwarp();
    int n = compact_inliers(dW, T, TC * th * MWM, iW, lane);
    lsq_f(u, iW, n, nullptr, f, lane);
    f_score_all(u, T, f, th, dS, nullptr, lane, &I, &J);
    __syncwarp();
    n = compact_inliers(dS, T, th, inl0, lane);
    if (lane == 0) { for (int i = 0; i < 9; i++) sh->f0[i] = f[i]; sh->n0 = n; st->lo_runs = lo_runs + 1; sh->lo_id = lo_runs + 1; }
  }
  if (lane == 0) sh->run_lo = run_lo ? 1 : 0;
}

// (c) inner RANSAC (exp_inFranicustom): one warp (= one CTA) per inner sample of <= 14 inliers
__global__ void __launch_bounds__(32)
k_rf_lo(const double* __restrict__ u, int T, double th, unsigned long long seed, FLoShare* sh, double* dscr, int* iscr) {
  if (!sh->run_lo) return;
  const int rep = blockIdx.x, lane = threadIdx.x;
  double* dW = dscr + (size_t)rep * 2 * T;
  double* wW = dscr + (size_t)(2 * RF_NW + 1) * T + (size_t)rep * T;
  int* iW = iscr + (size_t)rep * T;
  const int* inl0 = iscr + (size_t)RF_NW * T;
  const int n0 = sh->n0;
  int bI = 0; double bJ = 0; double Fb[9], f0[9];
  for (int i = 0; i < 9; i++) { f0[i] = sh->f0[i]; Fb[i] = f0[i]; }
  if (n0 >= 16) {
    int ssiz = n0 / 2; if (ssiz > 14) ssiz = 14;
    for (int k = lane; k < n0; k += 32) iW[k] = inl0[k];
    __syncwarp();
    if (lane == 0) {
      const unsigned long long stream = 0x464C000000000000ull + (unsigned long long)sh->lo_id * 64 + rep;
      for (int i = 0; i < ssiz; i++) {
        const int s = (int)rs_rand(seed, stream, i, (unsigned)(n0 - i)), j = n0 - i - 1;
        const int q = iW[s]; iW[s] = iW[j]; iW[j] = q;
      }
    }
    __syncwarp();
    double f[9];
    for (int i = 0; i < 9; i++) f[i] = f0[i];
    lsq_f(u, iW + n0 - ssiz, ssiz, nullptr, f, lane);
    int I; double J;
    f_score_all(u, T, f, th, dW, nullptr, lane, &I, &J);
    __syncwarp();
    f_lo_iterate(u, T, th, f, dW, dW + T, wW, iW, lane, &bI, &bJ, Fb);
  }
  if (lane == 0) { sh->loI[rep] = bI; sh->loJ[rep] = bJ; for (int i = 0; i < 9; i++) sh->loF[rep][i] = Fb[i]; }
}

// (d) accept the best inner sample (exp_ranF.c:1064-1074), update the stopping rule
__global__ void k_rf_accept(int T, double conf, int nhyp, RfState* st, const FLoShare* sh) {
  if (threadIdx.x != 0) return;
  if (sh->run_lo) {
    int best = -1; double bJ = 0; int bI = 0;
    for (int k = 0; k < LO_REPS; k++) if (bJ < sh->loJ[k]) { bJ = sh->loJ[k]; bI = sh->loI[k]; best = k; }
    if (best >= 0 && st->J < bJ) { for (int i = 0; i < 9; i++) st->F[i] = sh->loF[best][i]; st->J = bJ; st->I = bI; }
  }
  st->no_sam += nhyp;
  if (st->I > 0) { const int ns = nsamples(st->I + 1, T, 7, conf); if (ns < st->max_sam) st->max_sam = ns; }
  st->done = st->no_sam >= st->max_sam;
}

// ---- DEGENSAC branch (exp_ranF.c:963-1016; DegUtils.c rFtH :254-444, innerFH :488-594, u2Fit :635-691) -------------
// Runs only when a new best 7-point sample is plane-dominated (checksample), a handful of times per run, so it is
// orchestrated from the host with small kernels:
//   k_rfd_prep      consensus of the plane homography at 3*th; inlier list at 16*th for the inner H-RANSAC
//   k_rs_lo         (ransac_h.cuh) LO_REPS inner samples of the homography LO at threshold 16*th  = innerH
//   k_rfd_h_accept  refined H, its inliers (on-plane list) and the clearly off-plane correspondences (error > 100*th)
//   k_rfd_pp_hyp    plane-and-parallax hypotheses: epipole from two off-plane correspondences, F = ([e]x H)^T,
//                   support among the off-plane correspondences at 2*th (one warp per hypothesis)
//   k_rfd_pp_update best hypothesis of the batch -> innerFH (15 x {6 on-plane + 4 consistent off-plane -> 8-point LSQ
//                   -> iterated refit u2Fit}), stopping rule nsamples(., ., 2, 0.999)
//   k_rfd_finish    accept F if it has more inliers than the best model (exp_ranF.c:990-1013)
struct DegShare {
  double H[9], Fpp[9];
  int run_pp, nN, nH, max_i, m_i, max_sam, no_sam, have_F, done, pad;
};
struct PPHyp { double F[9]; int no_i, pad; };
constexpr int PP_B = 2048;

__device__ __forceinline__ int compact_lt(const double* d, int T, double th, int* idx, int lane) {
  int n = 0;
  for (int base = 0; base < T; base += 32) {
    const int j = base + lane;
    const bool in = j < T && d[j] < th;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) idx[n + __popc(m & ((1u << lane) - 1))] = j;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}
// Sampson errors under F into d[]; returns the number strictly below th  (the `Ds[i] < th` counts of DegUtils.c)
__device__ __forceinline__ int fds_all_count_lt(const double* __restrict__ u, int T, const double* F, double th, double* d, int lane) {
  int c = 0;
  for (int j = lane; j < T; j += 32) { const double e = fds(F, u + 6 * j, nullptr); d[j] = e; if (e < th) c++; }
  c = warp_sum_i(c);
  __syncwarp();
  return c;
}

__global__ void __launch_bounds__(32)
k_rfd_prep(const double* __restrict__ u, int T, double th, RfState* st, LoShare* hsh, double* dscr, int* iscr) {
  const int lane = threadIdx.x;
  hsh->run_lo = 0;
  if (!st->degen_pending) return;
  double H[9];
  for (int i = 0; i < 9; i++) H[i] = st->Hdeg[i];
  double* hd = dscr + (size_t)20 * T;
  int* inl0 = iscr + (size_t)RS_NW * T;
  int c3 = 0;
  for (int j = lane; j < T; j += 32) { const double e = sampson(H, u + 6 * j); hd[j] = e; if (e < 3 * th) c3++; }
  c3 = warp_sum_i(c3);
  __syncwarp();
  if (c3 < 8) { if (lane == 0) st->degen_pending = 0; return; }     // exp_ranF.c:972-974 `break`
  const int n = compact_inliers(hd, T, 16 * th, inl0, lane);
  if (lane == 0) {
    for (int i = 0; i < 9; i++) hsh->h0[i] = H[i];
    hsh->n0 = n; hsh->run_lo = 1; hsh->lo_id = 1000 + st->degen_cnt;
    for (int k = 0; k < LO_REPS; k++) { hsh->loJ[k] = 0; hsh->loI[k] = 0; }
  }
}

__global__ void __launch_bounds__(32)
k_rfd_h_accept(const double* __restrict__ u, int T, double th, RfState* st, const LoShare* hsh, DegShare* deg, double* dscr, int* iscr) {
  const int lane = threadIdx.x;
  if (lane == 0) { deg->run_pp = 0; deg->have_F = 0; deg->done = 1; deg->max_i = 0; }
  if (!st->degen_pending) return;
  double H[9];
  for (int i = 0; i < 9; i++) H[i] = st->Hdeg[i];
  int best = -1; double bJ = 0;
  for (int k = 0; k < LO_REPS; k++) if (bJ < hsh->loJ[k]) { bJ = hsh->loJ[k]; best = k; }
  if (best >= 0) for (int i = 0; i < 9; i++) H[i] = hsh->loH[best][i];
  double* hd = dscr + (size_t)20 * T;
  int* offp = iscr + (size_t)10 * T;
  int* onp = iscr + (size_t)11 * T;
  for (int j = lane; j < T; j += 32) hd[j] = sampson(H, u + 6 * j);
  __syncwarp();
  const int nH = compact_inliers(hd, T, 16 * th, onp, lane);        // innerH's inlier mask (DegUtils.c:715-723)
  if (lane == 0 && nH > st->Ihmax) st->Ihmax = nH;
  if (nH <= 6) { if (lane == 0) st->degen_pending = 0; return; }    // exp_ranF.c:984 `if (I > 6)`
  int nN = 0;                                                       // rFtH: nhinl = HDs > 100*th
  for (int base = 0; base < T; base += 32) {
    const int j = base + lane;
    const bool in = j < T && hd[j] > 100 * th;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) offp[nN + __popc(m & ((1u << lane) - 1))] = j;
    nN += __popc(m);
  }
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < 9; i++) deg->H[i] = H[i];
    deg->nN = nN; deg->nH = nH;
    deg->max_i = 3; deg->m_i = 4; deg->max_sam = 10000; deg->no_sam = 0; deg->have_F = 0;
    const bool ok = !(nN < 4 || nH < 6);                            // DegUtils.c:345
    if (!ok) deg->max_i = 0;
    deg->run_pp = ok ? 1 : 0; deg->done = ok ? 0 : 1;
  }
}

__global__ void __launch_bounds__(256)
k_rfd_pp_hyp(const double* __restrict__ u, int T, double th, unsigned long long seed, int base, int nhyp, const DegShare* deg,
             const int* __restrict__ iscr, PPHyp* __restrict__ out) {
  if (!deg->run_pp || deg->done) return;
  const int wv = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wv >= nhyp) return;
  const int nN = deg->nN;
  const int* offp = iscr + (size_t)10 * T;
  int pr[2];
  draw_sample<2>(seed, 0x5050000000000000ull + (unsigned long long)(base + wv), nN, pr);
  const double* H = deg->H;
  double c[2][3];
  for (int k = 0; k < 2; k++) {
    const double* p = u + 6 * offp[pr[k]];
    const double hx[3] = {H[0] * p[3] + H[3] * p[4] + H[6] * p[5], H[1] * p[3] + H[4] * p[4] + H[7] * p[5],
                          H[2] * p[3] + H[5] * p[4] + H[8] * p[5]};
    cross3(c[k], p, hx);
  }
  double ec[3];
  cross3(ec, c[0], c[1]);
  const double nrm = sqrt(ec[0] * ec[0] + ec[1] * ec[1] + ec[2] * ec[2]);
  for (int i = 0; i < 3; i++) ec[i] /= nrm;
  const double Ex[9] = {0, -ec[2], ec[1], ec[2], 0, -ec[0], -ec[1], ec[0], 0};
  double F[9];                                     // F = ([e]x * Hm)^T with Hm[i][j] = H[i + 3j]   (DegUtils.c:371-373)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) F[3 * j + i] = Ex[3 * i] * H[3 * j] + Ex[3 * i + 1] * H[3 * j + 1] + Ex[3 * i + 2] * H[3 * j + 2];
  int no_i = 0;
  if (isfinite(F[0]) && nrm > 0)
    for (int k = lane; k < nN; k += 32) if (fds(F, u + 6 * offp[k], nullptr) < th * 2) no_i++;
  no_i = warp_sum_i(no_i);
  if (lane == 0) {
    PPHyp o;
    for (int i = 0; i < 9; i++) o.F[i] = F[i];
    o.no_i = no_i; o.pad = 0;
    out[wv] = o;
  }
}

// u2Fit (DegUtils.c:635-691), warp-wide; F is refined in place; returns the inlier count at th
__device__ int u2Fit(const double* __restrict__ u, int T, double* F, double th, double ths, int iters, double* d, int* idx, int lane) {
  const double dth = (ths - th) / (iters - 1);
  for (int it = 0; it < iters; ++it) {
    fds_all_count_lt(u, T, F, ths, d, lane);
    const int n = compact_lt(d, T, ths, idx, lane);
    if (n < 8) return n;
    lsq_f(u, idx, n, nullptr, F, lane);
    ths -= dth;
  }
  return fds_all_count_lt(u, T, F, th, d, lane);
}

__global__ void __launch_bounds__(32)
k_rfd_pp_update(const double* __restrict__ u, int T, double th, unsigned long long seed, int nhyp, DegShare* deg,
                const PPHyp* __restrict__ hyp, double* dscr, int* iscr) {
  const int lane = threadIdx.x;
  if (!deg->run_pp || deg->done) return;
  const int nN = deg->nN, nH = deg->nH;
  const int* offp = iscr + (size_t)10 * T;
  const int* onp = iscr + (size_t)11 * T;
  const double* hd = dscr + (size_t)20 * T;
  double* d = dscr;                    // scratch slots of the (finished) inner H-RANSAC
  int* vlist = iscr;
  int* idx = iscr + T;
  // best hypothesis of the batch with more support than any before (first on ties)
  int bi = -1, bn = deg->m_i;
  for (int k = lane; k < nhyp; k += 32) if (hyp[k].no_i > bn) { bn = hyp[k].no_i; bi = k; }
  for (int o = 16; o > 0; o >>= 1) {
    const int on = __shfl_xor_sync(0xffffffffu, bn, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (on > bn || (on == bn && (bi < 0 || oi < bi)))) { bn = on; bi = oi; }
  }
  if (bi >= 0) {
    double Fp[9];
    for (int i = 0; i < 9; i++) Fp[i] = hyp[bi].F[i];
    // uV = off-plane correspondences consistent with the hypothesis
    int nV = 0;
    for (int base = 0; base < nN; base += 32) {
      const int k = base + lane;
      const bool in = k < nN && fds(Fp, u + 6 * offp[k], nullptr) < th * 2;
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (in) vlist[nV + __popc(m & ((1u << lane) - 1))] = offp[k];
      nV += __popc(m);
    }
    __syncwarp();
    // innerFH(uH, uV, u, th, 15, 6, 4)
    int max_i = 0, max_s = 0;
    double Fbest[9];
    for (int i = 0; i < 9; i++) Fbest[i] = 1.0;
    if (nV >= 4 && nH >= 6) {
      for (int rep = 0; rep < 15; ++rep) {
        int s6[6], s4[4], id10[10];
        const unsigned long long stream = 0x5046000000000000ull + (unsigned long long)deg->no_sam * 64 + rep;
        draw_sample<6>(seed, stream, nH, s6);
        draw_sample<4>(seed, stream + 32, nV, s4);
        for (int i = 0; i < 6; i++) id10[i] = onp[s6[i]];
        for (int i = 0; i < 4; i++) id10[6 + i] = vlist[s4[i]];
        double aF[9];
        for (int i = 0; i < 9; i++) aF[i] = 1.0;
        lsq_f(u, id10, 10, nullptr, aF, lane);
        int no_i = fds_all_count_lt(u, T, aF, th, d, lane);
        if (max_i < no_i) { max_i = no_i; for (int i = 0; i < 9; i++) Fbest[i] = aF[i]; }
        if (no_i > max_s) {
          max_s = no_i;
          no_i = u2Fit(u, T, aF, th, th * 3, 4, d, idx, lane);
          if (max_i < no_i) { max_i = no_i; for (int i = 0; i < 9; i++) Fbest[i] = aF[i]; }
        }
      }
    }
    if (lane == 0) deg->m_i = bn;
    if (max_i > deg->max_i) {
      // inliers that are clearly off the plane drive the stopping rule (DegUtils.c:416-424)
      int maxni = 0;
      for (int j = lane; j < T; j += 32) if (fds(Fbest, u + 6 * j, nullptr) < th && hd[j] > 100 * th) maxni++;
      maxni = warp_sum_i(maxni);
      if (lane == 0) {
        deg->max_i = max_i; deg->have_F = 1;
        for (int i = 0; i < 9; i++) deg->Fpp[i] = Fbest[i];
        const int ns = nsamples(maxni, nN, 2, 0.999);
        if (ns < deg->max_sam) deg->max_sam = ns;
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    deg->no_sam += nhyp;
    deg->done = deg->no_sam >= 2 * deg->max_sam;
  }
}

__global__ void __launch_bounds__(32)
k_rfd_finish(const double* __restrict__ u, int T, double th, double conf, RfState* st, const DegShare* deg) {
  const int lane = threadIdx.x;
  if (!st->degen_pending) return;
  if (deg->have_F && deg->max_i > st->I) {          // exp_ranF.c:990-996 `if (I > maxS.I)`
    double F[9];
    for (int i = 0; i < 9; i++) F[i] = deg->Fpp[i];
    int I; double J;
    f_score_all(u, T, F, th, nullptr, nullptr, lane, &I, &J);
    if (lane == 0) {
      for (int i = 0; i < 9; i++) st->F[i] = F[i];
      st->I = I; st->J = J;
      const int ns = nsamples(st->I + 1, T, 7, conf);
      if (ns < st->max_sam) st->max_sam = ns;
      st->done = st->no_sam >= st->max_sam;
    }
  }
  __syncwarp();
  if (lane == 0) { st->degen_cnt++; st->degen_pending = 0; }
}

__global__ void k_rf_final(const double* __restrict__ u, int T, double th, const RfState* st, unsigned char* inl) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T) return;
  inl[j] = (st->I > 0 && fds(st->F, u + 6 * j, nullptr) <= th) ? 1 : 0;
}

}  // namespace

int mg_ransac_F_run(modsgpu_ctx* ctx, const double* d_u, int T, const modsgpu_ransac_params* p,
                    double* F, unsigned char* inl, modsgpu_ransac_result* res) {
  const int NW = RF_NW;
  static_assert(sizeof(RfState) <= 512, "state slot");
  size_t off_sh = 512, off_hsh = off_sh + ((sizeof(FLoShare) + 255) & ~(size_t)255);
  size_t off_deg = off_hsh + ((sizeof(LoShare) + 255) & ~(size_t)255);
  size_t off_hyp = off_deg + ((sizeof(DegShare) + 255) & ~(size_t)255), off_d = off_hyp + sizeof(FHyp) * RS_MAX_B;
  size_t off_i = off_d + sizeof(double) * (size_t)(3 * NW + 1) * T;
  size_t off_inl = off_i + sizeof(int) * (size_t)(NW + 1) * T;
  size_t total = off_inl + T + 64;
  MG_CUDA(ctx, ctx->rs_buf.ensure(total));
  uint8_t* base = ctx->rs_buf.as<uint8_t>();
  RfState* st = reinterpret_cast<RfState*>(base);
  FLoShare* sh = reinterpret_cast<FLoShare*>(base + off_sh);
  LoShare* hsh = reinterpret_cast<LoShare*>(base + off_hsh);
  DegShare* deg = reinterpret_cast<DegShare*>(base + off_deg);
  const int do_degen = getenv("MODSGPU_NO_DEGENSAC") ? 0 : 1;
  FHyp* hyp = reinterpret_cast<FHyp*>(base + off_hyp);
  double* dscr = reinterpret_cast<double*>(base + off_d);
  int* iscr = reinterpret_cast<int*>(base + off_i);
  unsigned char* dinl = base + off_inl;
  MG_CUDA(ctx, ctx->h_stage.ensure(sizeof(RfState) + T + 64));
  RfState* hs = ctx->h_stage.as<RfState>();
  memset(hs, 0, sizeof(RfState));
  hs->max_sam = p->max_samples;
  MG_CUDA(ctx, cudaMemcpyAsync(st, hs, sizeof(RfState), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  int basei = 0, batch = 0;
  for (;;) {
    int B = batch == 0 ? 512 : (batch == 1 ? 1024 : RS_MAX_B);
    MG_PROF(ctx, "k_rf_hyp", 2, (double)B);
    k_rf_hyp<<<ceil_div(B, 8), 256, 0, ctx->stream>>>(d_u, T, p->th, p->seed, basei, B, hyp);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_rf_select", 2, (double)T);
    k_rf_select<<<1, 384, 0, ctx->stream>>>(d_u, T, p->th, p->do_sym_check, do_degen, hyp, B, 0, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_rf_lo", 2, (double)T);
    k_rf_lo<<<LO_REPS, 32, 0, ctx->stream>>>(d_u, T, p->th, p->seed, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rf_accept<<<1, 32, 0, ctx->stream>>>(T, p->conf, B, st, sh);
    MG_LAUNCHED(ctx);
    MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RfState), cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    if (hs->degen_pending) {
      // DEGENSAC: plane homography by an inner H-RANSAC at 16*th, then plane-and-parallax
      DegShare hd;
      k_rfd_prep<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, st, hsh, dscr, iscr);
      MG_LAUNCHED(ctx);
      k_rs_lo<<<LO_REPS, 32, 0, ctx->stream>>>(d_u, T, 16 * p->th, p->seed, hsh, dscr, iscr);
      MG_LAUNCHED(ctx);
      k_rfd_h_accept<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, st, hsh, deg, dscr, iscr);
      MG_LAUNCHED(ctx);
      int pbase = batch * 1000003;
      for (int it = 0; it < 64; it++) {
        MG_CUDA(ctx, cudaMemcpyAsync(&hd, deg, sizeof(DegShare), cudaMemcpyDeviceToHost, ctx->stream));
        MG_CUDA(ctx, mg_stream_sync(ctx));
        if (!hd.run_pp || hd.done) break;
        MG_PROF(ctx, "k_rfd_pp_hyp", 2, (double)PP_B);
        k_rfd_pp_hyp<<<ceil_div(PP_B, 8), 256, 0, ctx->stream>>>(d_u, T, p->th, p->seed, pbase, PP_B, deg, iscr,
                                                               reinterpret_cast<PPHyp*>(hyp));
        MG_LAUNCHED(ctx);
        MG_PROF(ctx, "k_rfd_pp_update", 2, (double)T);
        k_rfd_pp_update<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, p->seed, PP_B, deg, reinterpret_cast<const PPHyp*>(hyp), dscr, iscr);
        MG_LAUNCHED(ctx);
        pbase += PP_B;
      }
      k_rfd_finish<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, p->conf, st, deg);
      MG_LAUNCHED(ctx);
      MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RfState), cudaMemcpyDeviceToHost, ctx->stream));
      MG_CUDA(ctx, mg_stream_sync(ctx));
    }
    basei += B; batch++;
    if (hs->done) break;
  }
  if (hs->lo_runs == 0 && hs->degen_cnt == 0) {   // exp_ranF.c:1086: "If there were no LOs, do at least one NOW!"
    k_rf_select<<<1, 384, 0, ctx->stream>>>(d_u, T, p->th, p->do_sym_check, 0, hyp, 0, 1, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rf_lo<<<LO_REPS, 32, 0, ctx->stream>>>(d_u, T, p->th, p->seed, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rf_accept<<<1, 32, 0, ctx->stream>>>(T, p->conf, 0, st, sh);
    MG_LAUNCHED(ctx);
  }
  k_rf_final<<<ceil_div(T, 256), 256, 0, ctx->stream>>>(d_u, T, p->th, st, dinl);
  MG_LAUNCHED(ctx);
  unsigned char* hinl = reinterpret_cast<unsigned char*>(hs + 1);
  MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RfState), cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(hinl, dinl, T, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  const bool have = hs->J > 0;
  int ninl = 0;
  for (int i = 0; i < T; i++) { if (!have) hinl[i] = 0; ninl += hinl[i]; }
  for (int i = 0; i < 9; i++) F[i] = have ? hs->F[i] : 0.0;
  memcpy(inl, hinl, T);
  if (res) { res->n_inliers = ninl; res->J = hs->J; res->samples = hs->no_sam; res->lo_runs = hs->lo_runs; res->oc_rejects = hs->sym_rejects;
             res->degen_runs = hs->degen_cnt; res->h_inliers = hs->Ihmax; }
  return 0;
}

extern "C" int modsgpu_ransac_F(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                                double* F, unsigned char* inl, modsgpu_ransac_result* res) {
  if (!ctx || !p || !F || T < 0 || (T > 0 && (!u || !inl))) return MODSGPU_EINVAL;
  if (p->error_type != MODSGPU_ERR_SAMPSON)
    MG_FAIL(ctx, MODSGPU_EINVAL, "modsgpu_ransac_F implements the Sampson error (exFDs / FDs) only; the symmetric epipolar "
                                 "distance (exFDsSym / FDsSym, Ftools.c:103-124, :177-199) is not built");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (res) memset(res, 0, sizeof(*res));
  for (int i = 0; i < 9; i++) F[i] = 0;
  if (T < 8) {   // fewer than a minimal sample + 1: no model
    for (int i = 0; i < T; i++) inl[i] = 0;
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  }
  MG_CUDA(ctx, ctx->io_a.ensure((size_t)T * 48));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, u, (size_t)T * 48, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_ransac_F_run(ctx, ctx->io_a.as<double>(), T, p, F, inl, res);
  if (rc) return rc;
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}
