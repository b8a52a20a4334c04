// api.cu -- context life cycle of libmodsgpu.so
#include "common.cuh"

void mg_free_nets(modsgpu_ctx* ctx);  // cnn.cu

extern "C" const char* modsgpu_version(void) { return "modsgpu 0.1 (sm_100a)"; }

extern "C" int modsgpu_create(int device, modsgpu_ctx** out) {
  if (!out) return MODSGPU_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MODSGPU_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MODSGPU_ENODEV;
  if (prop.major != 10) return MODSGPU_ENODEV;  // sm_100a code only; there is no other path
  if (cudaSetDevice(device) != cudaSuccess) return MODSGPU_ENODEV;
  modsgpu_ctx* ctx = new modsgpu_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    delete ctx;
    return MODSGPU_ECUDA;
  }
  *out = ctx;
  return 0;
}

extern "C" void modsgpu_destroy(modsgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mg_free_nets(ctx);
  DevBuf* bufs[] = {&ctx->det_pyr, &ctx->det_cand, &ctx->det_map, &ctx->det_out, &ctx->det_misc, &ctx->io_a, &ctx->io_b,
                    &ctx->io_c, &ctx->smp_regs, &ctx->smp_meta, &ctx->smp_taps, &ctx->smp_scratch, &ctx->smp_out,
                    &ctx->cnn_act0, &ctx->cnn_act1, &ctx->cnn_out, &ctx->mt_q, &ctx->mt_t, &ctx->mt_d, &ctx->mt_aux,
                    &ctx->mt_out, &ctx->rs_buf};
  for (DevBuf* b : bufs) b->release();
  ctx->h_stage.release();
  ctx->h_stage2.release();
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* modsgpu_last_error(const modsgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" float modsgpu_last_device_ms(const modsgpu_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }
extern "C" long long modsgpu_launch_count(const modsgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* modsgpu_stream(const modsgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
