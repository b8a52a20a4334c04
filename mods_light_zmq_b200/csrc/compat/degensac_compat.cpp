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
//   error-function pointers: accepted and ignored -- the library implements the Sampson error path that
//   LORANSACFiltering selects for errorType SAMPSON (matching.cpp:652-661); other error types are a documented
//   deviation
// Like the reference (global HASH_TABLE, libc rand) these entry points are NOT re-entrant: one call at a time.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
                        HDsPtr, HDsiPtr, HDsidxPtr, int doSymCheck) {
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
  p.th = th; p.conf = conf; p.max_samples = max_sam; p.do_sym_check = doSymCheck; p.seed = next_seed();
  modsgpu_ransac_result r;
  if (modsgpu_ransac_H(ctx, u, len, &p, H, inl, &r)) {
    fprintf(stderr, "modsgpu degensac shim: %s\n", modsgpu_last_error(ctx));
    memset(inl, 0, len);
    return s;
  }
  if (data_out) { data_out[0] = r.samples; data_out[1] = r.lo_runs; data_out[2] = r.oc_rejects; }
  if (resids && *resids) memset(*resids, 0, sizeof(double) * (size_t)len);
  s.I = (unsigned)r.n_inliers; s.J = r.J;
  return s;
}

int exp_ransacFcustom(double* u, int len, double th, double conf, int max_sam, double* F, unsigned char* inl,
                      int* data_out, int do_lo, unsigned inlLimit, double** resids, double* H_best, int* Ih,
                      exFDsPtr, FDsPtr, int doSymCheck) {
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
  p.th = th; p.conf = conf; p.max_samples = max_sam; p.do_sym_check = doSymCheck; p.seed = next_seed();
  modsgpu_ransac_result r;
  if (modsgpu_ransac_F(ctx, u, len, &p, F, inl, &r)) {
    fprintf(stderr, "modsgpu degensac shim: %s\n", modsgpu_last_error(ctx));
    memset(inl, 0, len);
    return 0;
  }
  if (data_out) { data_out[0] = r.samples; data_out[1] = r.lo_runs; }
  if (Ih) *Ih = r.h_inliers;
  if (resids && *resids) memset(*resids, 0, sizeof(double) * (size_t)len);
  return r.n_inliers;
}

}  // extern "C"
