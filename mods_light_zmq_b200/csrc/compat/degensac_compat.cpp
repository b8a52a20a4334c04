// degensac_compat.cpp -- link-time drop-in for the reference's degensac entry points (seam S4).
//
// Exports exp_ransacHcustom (degensac/exp_ranH.h:32-36) and exp_ransacFcustom (degensac/exp_ranF.h:71-73) with
// the reference's exact C signatures, implemented on the device by modsgpu_ransac_H / modsgpu_ransac_F.
// Built as libmodsgpu_degensac.so (links libmodsgpu.so): the reference links it instead of its `degensac`
// target and LORANSACFiltering (matching.cpp:637-823) runs unchanged.
//
// Contract kept from the reference (matching.cpp:718-735):
//   u      : len x 6 doubles (x1 y1 1 x2 y2 1), caller owned
//   H / F  : 9 doubles out (degensac convention)
//   inl    : len bytes out
//   data_out[0] samples drawn, [1] LO runs, [2] oriented-constraint rejects (caller allocated len*18 ints)
//   *resids: malloc()ed here, free()d by the caller (matching.cpp:724,:732)
//   error-function pointers: the three known H sets are mapped to the library's error_type (matching.cpp:652-681:
//   HDs -> Sampson, HDsSymMax -> SymmMax, HDsSym -> SymmSum) by comparing against the symbols this shim exports under
//   the reference's names; an unknown H pointer, or the symmetric F pair (exFDsSym / FDsSym), makes the call FAIL loudly
//   (empty result + message) instead of silently scoring with Sampson
// Like the reference (global HASH_TABLE, libc rand) these entry points are NOT re-entrant: one call at a time.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <mutex>
#include "../../../include/modsgpu.h"

extern "C" {

typedef struct { unsigned I; double J; } Score;   // degensac/rtools.h:18-24
typedef void (*HDsPtr)(const double*, const double*, const double*, double*, int);
typedef void (*HDsiPtr)(const double*, const double*, const double*, double*, int, int*, int);
typedef void (*HDsidxPtr)(const double*, const double*, const double*, double*, int, int*, int);
typedef void (*FDsPtr)(const double*, const double*, double*, int);
typedef void (*exFDsPtr)(const double*, const double*, double*, double*, int);

// The reference passes the addresses of ITS error functions.  A program linked against this shim instead of the
// reference's degensac resolves those names here: the functions below exist to BE those addresses (and they compute the
// reference's formulas on the host should anyone call them: Htools.c:160-284, Ftools.c:83-124).
static void h_err(const double* u, const double* H, double* p, int len, int type) {
  const double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]};
  double H1[9];
  {
    const double* A = Hm;
    const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2, id = 1.0 / det;
    H1[0] = c0 * id; H1[1] = (A[2] * A[7] - A[1] * A[8]) * id; H1[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    H1[3] = c1 * id; H1[4] = (A[0] * A[8] - A[2] * A[6]) * id; H1[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    H1[6] = c2 * id; H1[7] = (A[1] * A[6] - A[0] * A[7]) * id; H1[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  }
  for (int i = 0; i < len; i++, u += 6) {
    const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8], b = Hm[6] * u[3] + Hm[7] * u[4] + Hm[8];
    double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a, ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
    double xd = u[3] - xa, yd = u[4] - ya;
    const double d1 = xd * xd + yd * yd;
    xa = (Hm[0] * u[3] + Hm[1] * u[4] + Hm[2]) / b; ya = (Hm[3] * u[3] + Hm[4] * u[4] + Hm[5]) / b;
    xd = u[0] - xa; yd = u[1] - ya;
    const double d2 = xd * xd + yd * yd;
    p[i] = type == 2 ? d1 + d2 : (d1 > d2 ? d1 : d2);
  }
}
void HDsSym(const double*, const double* u, const double* H, double* p, int len) { h_err(u, H, p, len, 2); }
void HDsSymMax(const double*, const double* u, const double* H, double* p, int len) { h_err(u, H, p, len, 1); }
void HDs(const double*, const double* u, const double* H, double* p, int len) {
  // Htools.c:138-198 (pinvJ + HDs)
  for (int i = 0; i < len; i++, u += 6) {
    const double x1 = u[0], y1 = u[1], x2 = u[3], y2 = u[4], w2 = u[5];
    double r1 = 0, r2 = 0;
    r1 += H[0] * x2; r1 += H[2] * (-x1 * x2); r1 += H[3] * y2; r1 += H[5] * (-x1 * y2); r1 += H[6] * w2; r1 += H[8] * (-x1 * w2);
    r2 += H[1] * x2; r2 += H[2] * (-y1 * x2); r2 += H[4] * y2; r2 += H[5] * (-y1 * y2); r2 += H[7] * w2; r2 += H[8] * (-y1 * w2);
    const double a = H[0] - H[2] * x1, b = H[3] - H[5] * x1, c = -H[8] - H[2] * x2 - H[5] * y2;
    const double d = H[1] - H[2] * y1, e = H[4] - H[5] * y1;
    const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
    const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
    double pJ[8];
    pJ[0] = -b * de + a * (c2 + e2); pJ[1] = b * c2pd2 - a * de; pJ[2] = c * (c2pd2 + e2); pJ[3] = -c * (a * d + b * e);
    pJ[4] = d * (b2 + c2) - ab * e; pJ[5] = -ab * d + e * (a2 + c2); pJ[6] = pJ[3]; pJ[7] = c * (a2 + b2 + c2);
    const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
    double s = 0;
    for (int j = 0; j < 4; j++) { const double t = (pJ[j] / N) * r1 + (pJ[j + 4] / N) * r2; s += t * t; }
    p[i] = s;
  }
}
static void sub_idx(void (*f)(const double*, const double*, const double*, double*, int), const double* u6, const double* H,
                    double* p, int* pts, int ni) {
  for (int k = 0; k < ni; k++) f(nullptr, u6 + 6 * pts[k], H, p + k, 1);
}
void HDsi(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDs, u6, H, p, pts, ni); }
void HDsiSym(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDsSym, u6, H, p, pts, ni); }
void HDsiSymMax(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDsSymMax, u6, H, p, pts, ni); }
void HDsidx(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDs, u6, H, p, pts, ni); }
void HDsSymidx(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDsSym, u6, H, p, pts, ni); }
void HDsSymidxMax(const double*, const double* u6, const double* H, double* p, int, int* pts, int ni) { sub_idx(HDsSymMax, u6, H, p, pts, ni); }
static void f_err(const double* u, const double* F, double* p, double* w, int len, int sym) {
  for (int i = 0; i < len; i++, u += 6) {
    const double rxc = F[0] * u[3] + F[3] * u[4] + F[6], ryc = F[1] * u[3] + F[4] * u[4] + F[7], rwc = F[2] * u[3] + F[5] * u[4] + F[8];
    const double r = (u[0] * rxc + u[1] * ryc + rwc);
    const double rx = F[0] * u[0] + F[1] * u[1] + F[2], ry = F[3] * u[0] + F[4] * u[1] + F[5];
    const double a = rxc * rxc + ryc * ryc, b = rx * rx + ry * ry;
    if (sym) { const double ww = (a * b) / (a + b); p[i] = r * r / ww; if (w) w[i] = ww; }
    else { const double ww = rxc * rxc + ryc * ryc + rx * rx + ry * ry; p[i] = r * r / ww; if (w) w[i] = 1 / std::sqrt(ww); }
  }
}
void FDs(const double* u, const double* F, double* p, int len) { f_err(u, F, p, nullptr, len, 0); }
void FDsSym(const double* u, const double* F, double* p, int len) { f_err(u, F, p, nullptr, len, 1); }
void exFDs(const double* u, const double* F, double* p, double* w, int len) { f_err(u, F, p, w, len, 0); }
void exFDsSym(const double* u, const double* F, double* p, double* w, int len) { f_err(u, F, p, w, len, 1); }

static std::mutex g_mu;
static modsgpu_ctx* g_ctx = nullptr;
static uint64_t g_seed = 0;
static bool g_seed_set = false;

// the reference seeds with time(NULL) on every call (exp_ranH.c:823); call this for reproducible runs
void modsgpu_ransac_set_seed(uint64_t seed) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_seed = seed;
  g_seed_set = true;
}

static modsgpu_ctx* ctx_locked() {
  if (!g_ctx) {
    const char* e = getenv("MODSGPU_DEVICE");
    int rc = modsgpu_create(e ? atoi(e) : 0, &g_ctx);
    if (rc) {
      fprintf(stderr, "modsgpu degensac shim: modsgpu_create failed (%d): no sm_100 device; there is no CPU path\n", rc);
      g_ctx = nullptr;
    }
  }
  return g_ctx;
}

static uint64_t next_seed() {
  if (g_seed_set) return g_seed++;          // successive calls draw different, reproducible streams
  return (uint64_t)time(nullptr);
}

Score exp_ransacHcustom(double* u, int len, double th, double conf, int max_sam, double* H, unsigned char* inl,
                        int iter_type, int* data_out, int oriented_constraint, unsigned inlLimit, double** resids,
                        HDsPtr hds, HDsiPtr, HDsidxPtr, int doSymCheck) {
  (void)iter_type; (void)oriented_constraint; (void)inlLimit;
  std::lock_guard<std::mutex> lk(g_mu);
  Score s = {0, 0};
  if (resids) *resids = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  if (data_out) data_out[0] = data_out[1] = data_out[2] = 0;
  // the caller (matching.cpp:694-760) reads H and inl[0..len) unconditionally: every failure leaves an EMPTY result
  if (H) memset(H, 0, 9 * sizeof(double));
  if (inl && len > 0) memset(inl, 0, (size_t)len);
  if (resids && *resids) memset(*resids, 0, sizeof(double) * (size_t)(len > 0 ? len : 1));
  modsgpu_ctx* ctx = ctx_locked();
  if (!ctx || !u || !H || !inl || len <= 0) return s;
  modsgpu_ransac_params p;
  p.th = th; p.conf = conf; p.max_samples = max_sam; p.do_sym_check = doSymCheck; p.seed = next_seed(); p._pad = 0;
  if (hds == nullptr || hds == &HDs) p.error_type = MODSGPU_ERR_SAMPSON;
  else if (hds == &HDsSymMax) p.error_type = MODSGPU_ERR_SYMM_MAX;
  else if (hds == &HDsSym) p.error_type = MODSGPU_ERR_SYMM_SUM;
  else {
    fprintf(stderr, "modsgpu degensac shim: exp_ransacHcustom called with an unknown error function; refusing to guess\n");
    return s;
  }
  modsgpu_ransac_result r;
  if (modsgpu_ransac_H_resid(ctx, u, len, &p, H, inl, &r, resids ? *resids : nullptr)) {
    fprintf(stderr, "modsgpu degensac shim: %s\n", modsgpu_last_error(ctx));
    memset(inl, 0, len);
    return s;
  }
  if (data_out) { data_out[0] = r.samples; data_out[1] = r.lo_runs; data_out[2] = r.oc_rejects; }
  s.I = (unsigned)r.n_inliers; s.J = r.J;
  return s;
}

int exp_ransacFcustom(double* u, int len, double th, double conf, int max_sam, double* F, unsigned char* inl,
                      int* data_out, int do_lo, unsigned inlLimit, double** resids, double* H_best, int* Ih,
                      exFDsPtr exfds, FDsPtr fdsp, int doSymCheck) {
  (void)do_lo; (void)inlLimit;
  std::lock_guard<std::mutex> lk(g_mu);
  if (resids) *resids = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  if (data_out) data_out[0] = data_out[1] = 0;
  if (Ih) *Ih = 0;
  if (F) memset(F, 0, 9 * sizeof(double));
  if (H_best) memset(H_best, 0, 9 * sizeof(double));
  if (inl && len > 0) memset(inl, 0, (size_t)len);
  if (resids && *resids) memset(*resids, 0, sizeof(double) * (size_t)(len > 0 ? len : 1));
  modsgpu_ctx* ctx = ctx_locked();
  if (!ctx || !u || !F || !inl || len <= 0) return 0;
  modsgpu_ransac_params p;
  p.th = th; p.conf = conf; p.max_samples = max_sam; p.do_sym_check = doSymCheck; p.seed = next_seed(); p._pad = 0;
  p.error_type = MODSGPU_ERR_SAMPSON;
  if ((exfds != nullptr && exfds != &exFDs) || (fdsp != nullptr && fdsp != &FDs)) {
    fprintf(stderr, "modsgpu degensac shim: exp_ransacFcustom supports the Sampson error (exFDs / FDs) only; empty result\n");
    return 0;
  }
  modsgpu_ransac_result r;
  if (modsgpu_ransac_F(ctx, u, len, &p, F, inl, &r)) {
    fprintf(stderr, "modsgpu degensac shim: %s\n", modsgpu_last_error(ctx));
    memset(inl, 0, len);
    return 0;
  }
  if (data_out) { data_out[0] = r.samples; data_out[1] = r.lo_runs; }
  if (Ih) *Ih = r.h_inliers;
  if (resids && *resids) FDs(u, F, *resids, len);      // the errors of the returned model (exp_ranF.c fills them the same way)
  return r.n_inliers;
}

}  // extern "C"
