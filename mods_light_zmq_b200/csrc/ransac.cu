// ransac.cu -- LO-RANSAC (homography) on the device.  Placeholder until the batched kernel lands.
#include "common.cuh"
extern "C" int modsgpu_ransac_H(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                                double* H, unsigned char* inl, modsgpu_ransac_result* res) {
  (void)u; (void)T; (void)p; (void)H; (void)inl; (void)res;
  if (!ctx) return MODSGPU_EINVAL;
  MG_FAIL(ctx, MODSGPU_ESTATE, "modsgpu_ransac_H: not built yet");
}
