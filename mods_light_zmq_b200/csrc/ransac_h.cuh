// ransac_h.cuh -- warp-level homography estimation helpers shared by ransac.cu (LO-RANSAC for H) and ransac_f.cu
// (the DEGENSAC branch of the fundamental-matrix RANSAC estimates the dominant plane with them): Sampson error
// (Htools.c:138-198), symmetric transfer error (:201-242), 4-point solver, normalised-DLT least squares (u2h),
// the shrinking-threshold LO iteration (exp_iterHcustom).  fp64, --fmad=false.
#pragma once
#include "ransac_common.cuh"

namespace {

constexpr double CHECK_COEF = 9.0;
constexpr int MIN_GOOD_SYM_PTS = 5;

// exp_ranH.c:883-892 / :1021-1032: reject H close to singular
__device__ __forceinline__ bool det_ok(const double* h) {
  double v = det3(h), tol = h[8];
  if (tol == 0) {
    for (int i = 0; i < 9; ++i) tol += h[i] * h[i];
    tol = sqrt(tol);
    tol *= 0.001;
  }
  tol = tol * tol * tol;
  return !(fabs(v / tol) < 10e-2);
}

// Htools.c:138-158 pinvJ + :160-198 HDs for one correspondence (H column-major, maps image 2 -> image 1)
__device__ __forceinline__ double sampson(const double* H, const double* u) {
  const double x1 = u[0], y1 = u[1], x2 = u[3], y2 = u[4], w2 = u[5];
  double r1 = 0, r2 = 0;
  r1 += H[0] * x2; r1 += H[2] * (-x1 * x2); r1 += H[3] * y2; r1 += H[5] * (-x1 * y2); r1 += H[6] * w2; r1 += H[8] * (-x1 * w2);
  r2 += H[1] * x2; r2 += H[2] * (-y1 * x2); r2 += H[4] * y2; r2 += H[5] * (-y1 * y2); r2 += H[7] * w2; r2 += H[8] * (-y1 * w2);
  const double a = H[0] - H[2] * x1, b = H[3] - H[5] * x1, c = -H[8] - H[2] * x2 - H[5] * y2;
  const double d = H[1] - H[2] * y1, e = H[4] - H[5] * y1;
  const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
  const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
  const double Q = c * (c2pd2 + e2);
  double pJ[8];
  pJ[0] = -b * de + a * (c2 + e2);
  pJ[1] = b * c2pd2 - a * de;
  pJ[2] = Q;
  pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e;
  pJ[5] = -ab * d + e * (a2 + c2);
  pJ[6] = pJ[3];
  pJ[7] = c * (a2 + b2 + c2);
  const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
  double p = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double t = (pJ[j] / N) * r1 + (pJ[j + 4] / N) * r2;
    p += t * t;
  }
  return p;
}

// Htools.c:201-242 HDsSym for one correspondence; Hm = row-major 2->1 map, H1 = its inverse
__device__ __forceinline__ double sym_err(const double* Hm, const double* H1, const double* u) {
  const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8];
  const double b = Hm[6] * u[3] + Hm[7] * u[4] + Hm[8];
  double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a, ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
  double xd = u[3] - xa, yd = u[4] - ya;
  const double d1 = xd * xd + yd * yd;
  xa = (Hm[0] * u[3] + Hm[1] * u[4] + Hm[2]) / b; ya = (Hm[3] * u[3] + Hm[4] * u[4] + Hm[5]) / b;
  xd = u[0] - xa; yd = u[1] - ya;
  return d1 + (xd * xd + yd * yd);
}

// utools.c:97-167 nullspace() on the 9x9 (8 rows + zero row) system; returns the nullity and, when it is 1,
// the null vector in sol[9].
__device__ int nullspace9(double* m, double* sol) {
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
  if (nnp == 1) {
    const int j = nopivot[0];
    for (int l = 0; l < n - 1; l++) sol[pivotc[l]] = -m[l * n + j];
    sol[j] = 1;
  }
  return nnp;
}

// two DLT rows of one correspondence (lin_hg, Htools.c:19-57), h stored column-major
__device__ __forceinline__ void dlt_rows(const double* u, double* r1, double* r2) {
  const double x1 = u[0], y1 = u[1], x2 = u[3], y2 = u[4], w2 = u[5];
  r1[0] = x2; r1[1] = 0; r1[2] = -x1 * x2; r1[3] = y2; r1[4] = 0; r1[5] = -x1 * y2; r1[6] = w2; r1[7] = 0; r1[8] = -x1 * w2;
  r2[0] = 0; r2[1] = x2; r2[2] = -y1 * x2; r2[3] = 0; r2[4] = y2; r2[5] = -y1 * y2; r2[6] = 0; r2[7] = w2; r2[8] = -y1 * w2;
}

// Htools.c:526-551 all_Hori_valid
__device__ bool all_Hori_valid(const double* us, const int* idx) {
  const double *a = us + 6 * idx[0], *b = us + 6 * idx[1], *c = us + 6 * idx[2], *d = us + 6 * idx[3];
  double p[3], q[3];
  cross3(p, a, b); cross3(q, a + 3, b + 3);
  if ((p[0] * c[0] + p[1] * c[1] + p[2] * c[2]) * (q[0] * c[3] + q[1] * c[4] + q[2] * c[5]) < 0) return false;
  if ((p[0] * d[0] + p[1] * d[1] + p[2] * d[2]) * (q[0] * d[3] + q[1] * d[4] + q[2] * d[5]) < 0) return false;
  cross3(p, c, d); cross3(q, c + 3, d + 3);
  if ((p[0] * a[0] + p[1] * a[1] + p[2] * a[2]) * (q[0] * a[3] + q[1] * a[4] + q[2] * a[5]) < 0) return false;
  if ((p[0] * b[0] + p[1] * b[1] + p[2] * b[2]) * (q[0] * b[3] + q[1] * b[4] + q[2] * b[5]) < 0) return false;
  return true;
}

// minimal solver: H (column-major) from 4 correspondences; false when the null space is not 1-D
__device__ bool h_from_4(const double* u, const int* idx, double* h) {
  double M[81];
  for (int i = 0; i < 4; i++) dlt_rows(u + 6 * idx[i], M + 18 * i, M + 18 * i + 9);
  for (int i = 72; i < 81; i++) M[i] = 0.0;
  return nullspace9(M, h) == 1;
}

// warp-wide: Sampson errors of all T correspondences under H into d[] (optional) + MSAC score at th
__device__ void score_all(const double* __restrict__ u, int T, const double* H, double th, double* d, int lane, int* I, double* J) {
  int ci = 0;
  double cj = 0;
  for (int j = lane; j < T; j += 32) {
    const double e = sampson(H, u + 6 * j);
    if (d) d[j] = e;
    if (e <= th) ci++;
    cj += truncQuad(e, th);
  }
  *I = warp_sum_i(ci);
  *J = warp_sum_d(cj);
}

// warp-wide normalised-DLT least squares (u2h, Htools.c:100-132: normu + lin_hgN + cov_mat + smallest
// eigenvector + denormH).  idx[0..n) selects the correspondences.  n < 4 leaves H untouched.
__device__ void lsq_h(const double* __restrict__ u, const int* idx, int n, double* H, int lane) {
  if (n < 4) return;
  if (n == 4) {
    int id4[4] = {idx[0], idx[1], idx[2], idx[3]};
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = H[i];
    h_from_4(u, id4, h);   // like u2h's len == 4 branch; a degenerate sample keeps the previous H
    for (int i = 0; i < 9; i++) H[i] = h[i];
    return;
  }
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
  // normal matrix C = Z^T Z of the 2n x 9 design matrix (lin_hgN + cov_mat), lower triangle
  double C[45];
#pragma unroll
  for (int i = 0; i < 45; i++) C[i] = 0;
  for (int k = lane; k < n; k += 32) {
    const double* p = u + 6 * idx[k];
    const double a0 = p[0] * A1[0] + A1[1], a1 = p[1] * A1[0] + A1[2];
    const double b0 = p[3] * A2[0] + A2[1], b1 = p[4] * A2[0] + A2[2], b2 = 1;
    const double r1[9] = {b0, 0, -a0 * b0, b1, 0, -a0 * b1, b2, 0, -a0 * b2};
    const double r2[9] = {0, b0, -a1 * b0, 0, b1, -a1 * b1, 0, b2, -a1 * b2};
    int t = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
      for (int j = 0; j <= i; j++, t++) C[t] += r1[i] * r1[j] + r2[i] * r2[j];
  }
#pragma unroll
  for (int i = 0; i < 45; i++) C[i] = warp_sum_d(C[i]);
  // smallest eigenvector by inverse iteration on the Cholesky factor (every lane redundantly: deterministic)
  double L[45];
  double maxd = 0;
  { int t = 0; for (int i = 0; i < 9; i++) { t += i; if (C[t] > maxd) maxd = C[t]; t++; } }
  const double ridge = 1e-13 * maxd, tiny = 1e-30 * maxd + 1e-300;
  for (int i = 0; i < 9; i++) {
    for (int j = 0; j <= i; j++) {
      double s = C[i * (i + 1) / 2 + j] + (i == j ? ridge : 0.0);
      for (int k = 0; k < j; k++) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      if (i == j) L[i * (i + 1) / 2 + i] = sqrt(s > tiny ? s : tiny);
      else L[i * (i + 1) / 2 + j] = s / L[j * (j + 1) / 2 + j];
    }
  }
  double x[9];
  for (int i = 0; i < 9; i++) x[i] = 1.0 + 0.1 * i;
  for (int it = 0; it < 10; it++) {
    for (int i = 0; i < 9; i++) {           // L y = x
      double s = x[i];
      for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    for (int i = 8; i >= 0; i--) {          // L^T z = y
      double s = x[i];
      for (int k = i + 1; k < 9; k++) s -= L[k * (k + 1) / 2 + i] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    double nrm = 0;
    for (int i = 0; i < 9; i++) nrm += x[i] * x[i];
    nrm = sqrt(nrm);
    if (!(nrm > 0) || !isfinite(nrm)) return;   // keep the previous H
    for (int i = 0; i < 9; i++) x[i] /= nrm;
  }
  // denormH (utools.c:70-90)
  double* F = x;
  double r = A2[0], xx = A2[1], yy = A2[2];
  F[6] += xx * F[0] + yy * F[3];
  F[7] += xx * F[1] + yy * F[4];
  F[8] += xx * F[2] + yy * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r; F[3] *= r; F[4] *= r; F[5] *= r;
  r = 1 / A1[0]; xx = -A1[1] * r; yy = -A1[2] * r;
  for (int i = 0; i < 9; i += 3) {
    F[i] = r * F[i] + xx * F[i + 2];
    F[i + 1] = r * F[i + 1] + yy * F[i + 2];
  }
  for (int i = 0; i < 9; i++) H[i] = F[i];
}

__device__ bool sym_check_ok(const double* __restrict__ u, int T, const double* H, double th, int lane) {
  // exp_ranH.c:905-947: at least MIN_GOOD_SYM_PTS+1 correspondences within CHECK_COEF*th symmetric transfer error
  double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]}, H1[9];
  if (!inv3(Hm, H1)) return false;
  int c = 0;
  for (int j = lane; j < T; j += 32) if (sym_err(Hm, H1, u + 6 * j) <= CHECK_COEF * th) c++;
  return warp_sum_i(c) > MIN_GOOD_SYM_PTS;
}

// exp_iterHcustom (exp_ranH.c:617-737) for one inner sample, warp-wide.  d0 = errors of the start model h.
// Returns the best (I,J) seen and leaves the matching model in Hbest.
__device__ void lo_iterate(const double* __restrict__ u, int T, double th, double* h, const double* d0, double* d, int* idx,
                           int lane, int* bestI, double* bestJ, double* Hbest) {
  int mI = 0; double mJ = 0;
  for (int j = lane; j < T; j += 32) { if (d0[j] <= th) mI++; mJ += truncQuad(d0[j], th); }
  mI = warp_sum_i(mI); mJ = warp_sum_d(mJ);
  *bestI = 0; *bestJ = 0;
  if (mI < 4) return;
  for (int i = 0; i < 9; i++) Hbest[i] = h[i];
  int n = compact_inliers(d0, T, th * MWM, idx, lane);
  lsq_h(u, idx, n, h, lane);
  double ths = TC * th;
  const double dth = (ths - th) / ILSQ_ITERS;
  for (int it = 0; it < ILSQ_ITERS; it++) {
    int sI; double sJ;
    score_all(u, T, h, th, d, lane, &sI, &sJ);
    __syncwarp();
    n = compact_inliers(d, T, ths * MWM, idx, lane);
    if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
    if (n < 4) { *bestI = mI; *bestJ = mJ; return; }
    lsq_h(u, idx, n, h, lane);
    ths -= dth;
  }
  int sI; double sJ;
  score_all(u, T, h, th, nullptr, lane, &sI, &sJ);
  if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
  *bestI = mI; *bestJ = mJ;
}

// ---- local optimisation shared state + the inner-sample kernel (used by both RANSACs) ---------------------------
struct LoShare {
  double h0[9];
  int n0, run_lo, lo_id, pad;
  double loJ[LO_REPS]; int loI[LO_REPS]; double loH[LO_REPS][9];
};
constexpr int RS_NW = 12;    // scratch slots (>= LO_REPS)

// (c) inner RANSAC (exp_inHranicustom, exp_ranH.c:741-793): one warp (= one CTA, its own SM) per inner sample
__global__ void __launch_bounds__(32)
k_rs_lo(const double* __restrict__ u, int T, double th, unsigned long long seed, LoShare* sh, double* dscr, int* iscr) {
  if (!sh->run_lo) return;
  const int rep = blockIdx.x, lane = threadIdx.x;
  double* dW = dscr + (size_t)rep * 2 * T;
  int* iW = iscr + (size_t)rep * T;
  const int* inl0 = iscr + (size_t)RS_NW * T;
  const int n0 = sh->n0;
  int bI = 0; double bJ = 0; double Hb[9], h0[9];
  for (int i = 0; i < 9; i++) { h0[i] = sh->h0[i]; Hb[i] = h0[i]; }
  if (n0 >= 8) {
    int ssiz = n0 / 2; if (ssiz > 12) ssiz = 12;
    // randsubset (rtools.c:25-39) on a private copy of the inlier list
    for (int k = lane; k < n0; k += 32) iW[k] = inl0[k];
    __syncwarp();
    if (lane == 0) {
      const unsigned long long stream = 0x4C4F000000000000ull + (unsigned long long)sh->lo_id * 64 + rep;
      for (int i = 0; i < ssiz; i++) {
        const int s = (int)rs_rand(seed, stream, i, (unsigned)(n0 - i)), j = n0 - i - 1;
        const int q = iW[s]; iW[s] = iW[j]; iW[j] = q;
      }
    }
    __syncwarp();
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = h0[i];
    lsq_h(u, iW + n0 - ssiz, ssiz, h, lane);
    int I; double J;
    score_all(u, T, h, th, dW, lane, &I, &J);
    __syncwarp();
    lo_iterate(u, T, th, h, dW, dW + T, iW, lane, &bI, &bJ, Hb);
  }
  if (lane == 0) { sh->loI[rep] = bI; sh->loJ[rep] = bJ; for (int i = 0; i < 9; i++) sh->loH[rep][i] = Hb[i]; }
}

}  // namespace
