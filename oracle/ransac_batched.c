/* ransac_batched.c -- TEST INFRASTRUCTURE (oracle): a CPU restatement of the BATCHED LO-RANSAC(H) schedule that
 * mods_light_zmq_b200/csrc/ransac.cu runs on the device.  Only tests/ may load it.
 *
 * What it pins.  The estimator is the reference's exp_ransacHcustom (degensac/exp_ranH.c:796-1236: 4-point samples,
 * oriented constraint Htools.c:526-551, null space utools.c:97-167, determinant test exp_ranH.c:883-892, Sampson /
 * symmetric errors Htools.c:138-284, MSAC rtools.c truncQuad, symmetric check exp_ranH.c:905-947, local optimisation
 * exp_inHranicustom :741-793 + exp_iterHcustom :617-737, stopping rule rtools.c:196-224).  The reference consumes
 * libc rand() one sample at a time; the device consumes a counter-based generator in batches of 512 / 1024 / 4096
 * hypotheses and takes the best of a batch.  This file restates THAT schedule sequentially -- same generator, same batch
 * sizes, same order of every floating-point reduction (thread t of a 256-thread CTA adds the items t, t+256, ... in
 * ascending order; the 32 lanes of a warp are combined by the xor butterfly 16, 8, 4, 2, 1; the 8 warps are added in
 * order; a hypothesis warp strides by 32) -- so that inlier masks must be BYTE-equal and H equal to the last bit.
 * Compile with -ffp-contract=off (the device code is built --fmad=false).  Only log() in the stopping rule is a libm
 * call whose last bit may differ between glibc and CUDA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RS_MAX_B 4096
#define LO_REPS 10
#define ILSQ_ITERS 4
#define TC 4.0
#define MWM 2.0
#define ITER_SAM 50
#define NT 256
#define NWARP 8
#define CHECK_COEF 9.0
#define MIN_GOOD_SYM_PTS 5
enum { ERR_SAMPSON = 0, ERR_SYMM_MAX = 1, ERR_SYMM_SUM = 2 };

typedef struct {
  double H[9]; double J; int I;
  double Hs[9]; double Js; int Is;
  int max_sam, no_sam, lo_runs, oc_rejects, done, have_sample;
} State;
typedef struct { double H[9]; double J; int I; int flag; } HypOut;
typedef struct { double H[9], Hm[9], H1[9]; int type, ok; } HErr;

/* ---- counter-based generator (ransac_common.cuh) ---- */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static unsigned rs_rand(uint64_t seed, uint64_t stream, unsigned draw, unsigned range) {
  return (unsigned)(mix64(seed ^ mix64(stream * 0x100000001B3ull + draw)) % range);
}
/* rtools.c sample(): partial Fisher-Yates over a virtual pool */
static void draw_sample4(uint64_t seed, uint64_t stream, int T, int* idx) {
  int pos[4], val[4];
  for (int i = 0; i < 4; i++) {
    const int s = (int)rs_rand(seed, stream, i, (unsigned)(T - i)), last = T - i - 1;
    int vs = s, vl = last;
    for (int k = 0; k < i; k++) { if (pos[k] == s) vs = val[k]; if (pos[k] == last) vl = val[k]; }
    idx[i] = vs;
    pos[i] = s; val[i] = vl;
  }
}

/* ---- the reduction orders of the device ---- */
static double bfly_d(const double* v) {
  double a[32], b[32];
  memcpy(a, v, sizeof(a));
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; l++) b[l] = a[l] + a[l ^ o];
    memcpy(a, b, sizeof(a));
  }
  return a[0];
}
static double blk_sum_d(const double* part /* NT */) {
  double c[NWARP];
  for (int w = 0; w < NWARP; w++) c[w] = bfly_d(part + 32 * w);
  double s = c[0];
  for (int w = 1; w < NWARP; w++) s += c[w];
  return s;
}

/* ---- geometry (ransac_common.cuh / ransac_h.cuh) ---- */
static double truncQuad(double eps, double thr) {
  if (thr == 0) return 0;
  if (eps >= thr * 9 / 4) return 0;
  return 1 - (eps / (thr * 9 / 4));
}
static double det3(const double* A) {
  double r = (A[0] * A[4] * A[8] + A[2] * A[3] * A[7] + A[1] * A[5] * A[6]);
  r -= (A[2] * A[4] * A[6] + A[0] * A[5] * A[7] + A[1] * A[3] * A[8]);
  return r;
}
static int inv3(const double* A, double* R) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0 || !isfinite(det)) return 0;
  const double id = 1.0 / det;
  R[0] = c0 * id; R[1] = (A[2] * A[7] - A[1] * A[8]) * id; R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c1 * id; R[4] = (A[0] * A[8] - A[2] * A[6]) * id; R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c2 * id; R[7] = (A[1] * A[6] - A[0] * A[7]) * id; R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return 1;
}
static void cross3(double* o, const double* a, const double* b) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static int det_ok(const double* h) {
  double v = det3(h), tol = h[8];
  if (tol == 0) {
    for (int i = 0; i < 9; ++i) tol += h[i] * h[i];
    tol = sqrt(tol);
    tol *= 0.001;
  }
  tol = tol * tol * tol;
  return !(fabs(v / tol) < 10e-2);
}
static double sampson(const double* H, const double* u) {
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
  for (int j = 0; j < 4; j++) {
    double t = (pJ[j] / N) * r1 + (pJ[j + 4] / N) * r2;
    p += t * t;
  }
  return p;
}
static double sym_err(const double* Hm, const double* H1, const double* u) {
  const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8];
  const double b = Hm[6] * u[3] + Hm[7] * u[4] + Hm[8];
  double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a, ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
  double xd = u[3] - xa, yd = u[4] - ya;
  const double d1 = xd * xd + yd * yd;
  xa = (Hm[0] * u[3] + Hm[1] * u[4] + Hm[2]) / b; ya = (Hm[3] * u[3] + Hm[4] * u[4] + Hm[5]) / b;
  xd = u[0] - xa; yd = u[1] - ya;
  return d1 + (xd * xd + yd * yd);
}
static void herr_setup(HErr* e, const double* H, int type) {
  for (int i = 0; i < 9; i++) e->H[i] = H[i];
  e->type = type; e->ok = 1;
  if (type != ERR_SAMPSON) {
    const double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]};
    for (int i = 0; i < 9; i++) e->Hm[i] = Hm[i];
    e->ok = inv3(e->Hm, e->H1);
  }
}
static double herr(const HErr* e, const double* u) {
  if (e->type == ERR_SAMPSON) return sampson(e->H, u);
  if (!e->ok) return 1e300;
  const double a = e->H1[6] * u[0] + e->H1[7] * u[1] + e->H1[8];
  const double b = e->Hm[6] * u[3] + e->Hm[7] * u[4] + e->Hm[8];
  double xa = (e->H1[0] * u[0] + e->H1[1] * u[1] + e->H1[2]) / a, ya = (e->H1[3] * u[0] + e->H1[4] * u[1] + e->H1[5]) / a;
  double xd = u[3] - xa, yd = u[4] - ya;
  const double d1 = xd * xd + yd * yd;
  xa = (e->Hm[0] * u[3] + e->Hm[1] * u[4] + e->Hm[2]) / b; ya = (e->Hm[3] * u[3] + e->Hm[4] * u[4] + e->Hm[5]) / b;
  xd = u[0] - xa; yd = u[1] - ya;
  const double d2 = xd * xd + yd * yd;
  return e->type == ERR_SYMM_SUM ? d1 + d2 : (d1 > d2 ? d1 : d2);
}
/* utools.c:97-167 */
static int nullspace9(double* m, double* sol) {
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
static void dlt_rows(const double* u, double* r1, double* r2) {
  const double x1 = u[0], y1 = u[1], x2 = u[3], y2 = u[4], w2 = u[5];
  r1[0] = x2; r1[1] = 0; r1[2] = -x1 * x2; r1[3] = y2; r1[4] = 0; r1[5] = -x1 * y2; r1[6] = w2; r1[7] = 0; r1[8] = -x1 * w2;
  r2[0] = 0; r2[1] = x2; r2[2] = -y1 * x2; r2[3] = 0; r2[4] = y2; r2[5] = -y1 * y2; r2[6] = 0; r2[7] = w2; r2[8] = -y1 * w2;
}
static int all_Hori_valid(const double* us, const int* idx) {
  const double *a = us + 6 * idx[0], *b = us + 6 * idx[1], *c = us + 6 * idx[2], *d = us + 6 * idx[3];
  double p[3], q[3];
  cross3(p, a, b); cross3(q, a + 3, b + 3);
  if ((p[0] * c[0] + p[1] * c[1] + p[2] * c[2]) * (q[0] * c[3] + q[1] * c[4] + q[2] * c[5]) < 0) return 0;
  if ((p[0] * d[0] + p[1] * d[1] + p[2] * d[2]) * (q[0] * d[3] + q[1] * d[4] + q[2] * d[5]) < 0) return 0;
  cross3(p, c, d); cross3(q, c + 3, d + 3);
  if ((p[0] * a[0] + p[1] * a[1] + p[2] * a[2]) * (q[0] * a[3] + q[1] * a[4] + q[2] * a[5]) < 0) return 0;
  if ((p[0] * b[0] + p[1] * b[1] + p[2] * b[2]) * (q[0] * b[3] + q[1] * b[4] + q[2] * b[5]) < 0) return 0;
  return 1;
}
static int h_from_4(const double* u, const int* idx, double* h) {
  double M[81];
  for (int i = 0; i < 4; i++) dlt_rows(u + 6 * idx[i], M + 18 * i, M + 18 * i + 9);
  for (int i = 72; i < 81; i++) M[i] = 0.0;
  return nullspace9(M, h) == 1;
}
/* rtools.c:196-224 */
static int nsamples(int ninl, int ptNum, int samsiz, double conf) {
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

/* ---- block-wide stages ---- */
static void blk_score(const double* u, int T, const HErr* E, double th, double* d, int* I, double* J) {
  double part[NT];
  int ci = 0;
  for (int t = 0; t < NT; t++) {
    double cj = 0;
    for (int j = t; j < T; j += NT) {
      const double e = herr(E, u + 6 * j);
      if (d) d[j] = e;
      if (e <= th) ci++;
      cj += truncQuad(e, th);
    }
    part[t] = cj;
  }
  *I = ci;
  *J = blk_sum_d(part);
}
static int compact(const double* d, int T, double th, int* idx) {
  int n = 0;
  for (int j = 0; j < T; j++) if (d[j] <= th) idx[n++] = j;
  return n;
}
static void blk_lsq(const double* u, const int* idx, int n, double* H) {
  if (n < 4) return;
  if (n == 4) {
    int id4[4] = {idx[0], idx[1], idx[2], idx[3]};
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = H[i];
    h_from_4(u, id4, h);
    for (int i = 0; i < 9; i++) H[i] = h[i];
    return;
  }
  double part[4][NT];
  for (int t = 0; t < NT; t++) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int k = t; k < n; k += NT) { const double* p = u + 6 * idx[k]; s0 += p[0]; s1 += p[1]; s2 += p[3]; s3 += p[4]; }
    part[0][t] = s0; part[1][t] = s1; part[2][t] = s2; part[3][t] = s3;
  }
  const double m1x = blk_sum_d(part[0]) / n, m1y = blk_sum_d(part[1]) / n, m2x = blk_sum_d(part[2]) / n, m2y = blk_sum_d(part[3]) / n;
  for (int t = 0; t < NT; t++) {
    double q1 = 0, q2 = 0;
    for (int k = t; k < n; k += NT) {
      const double* p = u + 6 * idx[k];
      double a = p[0] - m1x, b = p[1] - m1y;
      q1 += sqrt(a * a + b * b);
      a = p[3] - m2x; b = p[4] - m2y;
      q2 += sqrt(a * a + b * b);
    }
    part[0][t] = q1; part[1][t] = q2;
  }
  const double q1 = blk_sum_d(part[0]), q2 = blk_sum_d(part[1]);
  double A1[3] = {q1, m1x, m1y}, A2[3] = {q2, m2x, m2y};
  if (A1[0] != 0) A1[0] = n * sqrt(2.0) / A1[0];
  if (A2[0] != 0) A2[0] = n * sqrt(2.0) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0]; A2[1] *= -A2[0]; A2[2] *= -A2[0];
  static double cpart[45][NT];
  for (int t = 0; t < NT; t++) {
    double C[45];
    for (int i = 0; i < 45; i++) C[i] = 0;
    for (int k = t; k < n; k += NT) {
      const double* p = u + 6 * idx[k];
      const double a0 = p[0] * A1[0] + A1[1], a1 = p[1] * A1[0] + A1[2];
      const double b0 = p[3] * A2[0] + A2[1], b1 = p[4] * A2[0] + A2[2], b2 = 1;
      const double r1[9] = {b0, 0, -a0 * b0, b1, 0, -a0 * b1, b2, 0, -a0 * b2};
      const double r2[9] = {0, b0, -a1 * b0, 0, b1, -a1 * b1, 0, b2, -a1 * b2};
      int tt = 0;
      for (int i = 0; i < 9; i++)
        for (int j = 0; j <= i; j++, tt++) C[tt] += r1[i] * r1[j] + r2[i] * r2[j];
    }
    for (int i = 0; i < 45; i++) cpart[i][t] = C[i];
  }
  double C[45];
  for (int i = 0; i < 45; i++) C[i] = blk_sum_d(cpart[i]);
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
    for (int i = 0; i < 9; i++) {
      double s = x[i];
      for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    for (int i = 8; i >= 0; i--) {
      double s = x[i];
      for (int k = i + 1; k < 9; k++) s -= L[k * (k + 1) / 2 + i] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    double nrm = 0;
    for (int i = 0; i < 9; i++) nrm += x[i] * x[i];
    nrm = sqrt(nrm);
    if (!(nrm > 0) || !isfinite(nrm)) return;
    for (int i = 0; i < 9; i++) x[i] /= nrm;
  }
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
static int sym_ok(const double* u, int T, const double* H, double th) {
  double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]}, H1[9];
  if (!inv3(Hm, H1)) return 0;
  int c = 0;
  for (int j = 0; j < T; j++) if (sym_err(Hm, H1, u + 6 * j) <= CHECK_COEF * th) c++;
  return c > MIN_GOOD_SYM_PTS;
}
static void lo_iterate(const double* u, int T, double th, int etype, double* h, const double* d0, double* d, int* idx,
                       int* bestI, double* bestJ, double* Hbest) {
  double part[NT];
  int mI = 0;
  for (int t = 0; t < NT; t++) {
    double s = 0;
    for (int j = t; j < T; j += NT) { if (d0[j] <= th) mI++; s += truncQuad(d0[j], th); }
    part[t] = s;
  }
  double mJ = blk_sum_d(part);
  *bestI = 0; *bestJ = 0;
  if (mI < 4) return;
  for (int i = 0; i < 9; i++) Hbest[i] = h[i];
  int n = compact(d0, T, th * MWM, idx);
  blk_lsq(u, idx, n, h);
  double ths = TC * th;
  const double dth = (ths - th) / ILSQ_ITERS;
  HErr E;
  for (int it = 0; it < ILSQ_ITERS; it++) {
    int sI; double sJ;
    herr_setup(&E, h, etype);
    blk_score(u, T, &E, th, d, &sI, &sJ);
    n = compact(d, T, ths * MWM, idx);
    if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
    if (n < 4) { *bestI = mI; *bestJ = mJ; return; }
    blk_lsq(u, idx, n, h);
    ths -= dth;
  }
  int sI; double sJ;
  herr_setup(&E, h, etype);
  blk_score(u, T, &E, th, NULL, &sI, &sJ);
  if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
  *bestI = mI; *bestJ = mJ;
}

typedef struct { double h0[9]; int n0, run_lo, lo_id; double loJ[LO_REPS]; int loI[LO_REPS]; double loH[LO_REPS][9]; } LoShare;

/* k_rs_hyp: one warp per hypothesis */
static void hyp_batch(const double* u, int T, double th, int etype, uint64_t seed, int base, int nhyp, HypOut* out) {
  for (int w = 0; w < nhyp; w++) {
    int idx[4];
    draw_sample4(seed, (uint64_t)(base + w), T, idx);
    double h[9];
    int flag = 0;
    if (!all_Hori_valid(u, idx)) flag = 1;
    else if (!h_from_4(u, idx, h) || !det_ok(h)) flag = 2;
    int I = 0; double J = 0;
    if (flag == 0) {
      HErr E;
      herr_setup(&E, h, etype);
      double part[32];
      for (int l = 0; l < 32; l++) {
        double cj = 0;
        for (int j = l; j < T; j += 32) { const double e = herr(&E, u + 6 * j); if (e <= th) I++; cj += truncQuad(e, th); }
        part[l] = cj;
      }
      J = bfly_d(part);
    }
    for (int i = 0; i < 9; i++) out[w].H[i] = flag == 0 ? h[i] : 0.0;
    out[w].I = I; out[w].J = J; out[w].flag = flag;
  }
}

/* k_rsb_select */
static void select_batch(const double* u, int T, double th, int etype, int do_sym, const HypOut* hyp, int nhyp, int force_lo,
                         State* st, LoShare* sh, double* dW, int* iW, double* dS, int* inl0) {
  double bj = -1; int bi = -1, rej = 0;
  for (int k = 0; k < nhyp; k++) {
    if (hyp[k].flag == 1) rej++;
    if (hyp[k].flag == 0 && hyp[k].J > bj) { bj = hyp[k].J; bi = k; }      /* max J, lowest index on ties */
  }
  const double curJ = st->J;
  double curJs = st->Js;
  int have = st->have_sample, curIs = st->Is;
  const int no_sam = st->no_sam, lo_runs = st->lo_runs;
  int run_lo = 0;
  if (bi >= 0) {
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = hyp[bi].H[i];
    const int I = hyp[bi].I;
    if (curJ < bj) {
      const int ok = !do_sym || sym_ok(u, T, h, th);
      if (ok) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bj; st->I = I; }
    }
    if (!have || curJs < bj) {
      for (int i = 0; i < 9; i++) st->Hs[i] = h[i];
      st->Js = bj; st->Is = I; st->have_sample = 1;
      have = 1; curJs = bj; curIs = I;
      run_lo = no_sam + nhyp > ITER_SAM;
    }
  }
  st->oc_rejects += rej;
  if (no_sam + nhyp >= ITER_SAM && lo_runs == 0 && have && curIs > 4) run_lo = 1;
  if (force_lo) run_lo = have && lo_runs == 0;
  if (run_lo) {
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = st->Hs[i];
    HErr E;
    int I; double J;
    herr_setup(&E, h, etype);
    blk_score(u, T, &E, th, dW, &I, &J);
    int n = compact(dW, T, TC * th * MWM, iW);
    blk_lsq(u, iW, n, h);
    herr_setup(&E, h, etype);
    blk_score(u, T, &E, th, dS, &I, &J);
    n = compact(dS, T, th, inl0);
    for (int i = 0; i < 9; i++) sh->h0[i] = h[i];
    sh->n0 = n; st->lo_runs = lo_runs + 1; sh->lo_id = lo_runs + 1;
  }
  sh->run_lo = run_lo;
}

/* k_rsb_lo */
static void lo_batch(const double* u, int T, double th, int etype, uint64_t seed, LoShare* sh, const int* inl0, double* dW, int* iW) {
  if (!sh->run_lo) return;
  for (int rep = 0; rep < LO_REPS; rep++) {
    const int n0 = sh->n0;
    int bI = 0; double bJ = 0; double Hb[9], h0[9];
    for (int i = 0; i < 9; i++) { h0[i] = sh->h0[i]; Hb[i] = h0[i]; }
    if (n0 >= 8) {
      int ssiz = n0 / 2; if (ssiz > 12) ssiz = 12;
      for (int k = 0; k < n0; k++) iW[k] = inl0[k];
      const uint64_t stream = 0x4C4F000000000000ull + (uint64_t)sh->lo_id * 64 + rep;
      for (int i = 0; i < ssiz; i++) {
        const int s = (int)rs_rand(seed, stream, i, (unsigned)(n0 - i)), j = n0 - i - 1;
        const int q = iW[s]; iW[s] = iW[j]; iW[j] = q;
      }
      double h[9];
      for (int i = 0; i < 9; i++) h[i] = h0[i];
      blk_lsq(u, iW + n0 - ssiz, ssiz, h);
      HErr E;
      herr_setup(&E, h, etype);
      int I; double J;
      blk_score(u, T, &E, th, dW, &I, &J);
      lo_iterate(u, T, th, etype, h, dW, dW + T, iW, &bI, &bJ, Hb);
    }
    sh->loI[rep] = bI; sh->loJ[rep] = bJ;
    for (int i = 0; i < 9; i++) sh->loH[rep][i] = Hb[i];
  }
}

/* k_rsb_accept */
static void accept_batch(const double* u, int T, double th, double conf, int do_sym, int nhyp, int closing, State* st, const LoShare* sh) {
  if (closing && !sh->run_lo) return;
  if (sh->run_lo) {
    int best = -1; double bJ = 0; int bI = 0;
    for (int k = 0; k < LO_REPS; k++) if (bJ < sh->loJ[k]) { bJ = sh->loJ[k]; bI = sh->loI[k]; best = k; }
    const double curJ = st->J;
    if (best >= 0 && curJ < bJ) {
      double h[9];
      for (int i = 0; i < 9; i++) h[i] = sh->loH[best][i];
      if (det_ok(h) && (!do_sym || sym_ok(u, T, h, th))) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bJ; st->I = bI; }
    }
  }
  if (!closing) {
    st->no_sam += nhyp;
    if (st->I > 0) { const int ns = nsamples(st->I + 1, T, 4, conf); if (ns < st->max_sam) st->max_sam = ns; }
    st->done = st->no_sam >= st->max_sam;
  }
}

/* u: T x 6 doubles.  stats: [0] I, [1] samples, [2] LO runs, [3] oriented-constraint rejects.  Returns 0. */
int orb_ransac_H(const double* u, int T, double th, double conf, int max_samples, int do_sym, uint64_t seed, int etype,
                 double* H, unsigned char* inl, int* stats, double* Jout, double* resid) {
  for (int i = 0; i < 9; i++) H[i] = 0;
  if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
  if (Jout) *Jout = 0;
  if (T < 4) { for (int i = 0; i < T; i++) { inl[i] = 0; if (resid) resid[i] = 0; } return 0; }
  State st;
  memset(&st, 0, sizeof(st));
  st.max_sam = max_samples;
  LoShare sh;
  memset(&sh, 0, sizeof(sh));
  HypOut* hyp = (HypOut*)malloc(sizeof(HypOut) * RS_MAX_B);
  double* dW = (double*)malloc(sizeof(double) * 2 * (size_t)T);
  double* dS = (double*)malloc(sizeof(double) * (size_t)T);
  int* iW = (int*)malloc(sizeof(int) * (size_t)T);
  int* inl0 = (int*)malloc(sizeof(int) * (size_t)T);
  int base = 0;
  for (int batch = 0; !st.done; batch++) {
    const int B = batch == 0 ? 512 : (batch == 1 ? 1024 : RS_MAX_B);
    hyp_batch(u, T, th, etype, seed, base, B, hyp);
    select_batch(u, T, th, etype, do_sym, hyp, B, 0, &st, &sh, dW, iW, dS, inl0);
    lo_batch(u, T, th, etype, seed, &sh, inl0, dW, iW);
    accept_batch(u, T, th, conf, do_sym, B, 0, &st, &sh);
    base += B;
  }
  if (st.lo_runs == 0) {     /* exp_ranH.c:1085-1197 */
    select_batch(u, T, th, etype, do_sym, hyp, 0, 1, &st, &sh, dW, iW, dS, inl0);
    lo_batch(u, T, th, etype, seed, &sh, inl0, dW, iW);
    accept_batch(u, T, th, conf, do_sym, 0, 1, &st, &sh);
  }
  HErr E;
  herr_setup(&E, st.H, etype);
  for (int j = 0; j < T; j++) {
    const double e = st.I > 0 ? herr(&E, u + 6 * j) : 0.0;
    if (resid) resid[j] = e;
    inl[j] = (st.I > 0 && e <= th) ? 1 : 0;
  }
  for (int i = 0; i < 9; i++) H[i] = st.H[i];
  if (stats) { stats[0] = st.I; stats[1] = st.no_sam; stats[2] = st.lo_runs; stats[3] = st.oc_rejects; }
  if (Jout) *Jout = st.J;
  free(hyp); free(dW); free(dS); free(iW); free(inl0);
  return 0;
}
