// common.cuh -- context, workspace buffers and error plumbing shared by the kernels of libmodsgpu.so
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/modsgpu.h"

// A grow-only device buffer: cudaMalloc is far too slow to sit on the per-image path.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Pinned host staging buffer (grow-only).
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct modsgpu_image {
  float* d = nullptr;  // device, w*h, dense
  int w = 0, h = 0;
};

struct NetWeights;  // cnn.cu

// One region of a view while it lives on the device between the stages of the per-view chain (chain.cu): the
// AffineRegion fields the hot path reads (structures.hpp:185-229).  Same bytes as modsgpu_view_region.
struct DevRegion {
  modsgpu_region det;      // det_kp, view coordinates
  modsgpu_region reproj;   // reproj_kp, original image
  double response;
  int octave, type;
};
static_assert(sizeof(DevRegion) == 128, "DevRegion layout");

// Upper bounds the host needs to size the sampler's launches for ANY subset of a view's keypoints: the window size R of a
// region depends on its scale only, and the scale never changes between the three sampler passes of a view, so the
// figures of the first pass (all keypoints) bound the later ones.  Written by k_smp_stats, read back once per view.
struct SmpStats {
  int cls_cnt[6], cls_rmax[6], cls_kr[6];
  int nl, pre0, pre1, pre2, maxPS, maxR, maxks, _pad;
  long long scratch;
};

// Optional per-launch CUDA-event timing (bench.py's roofline block).  kind: 0 = HBM-bound (work in bytes),
// 1 = tensor-bound (work in flops), 2 = latency-bound (work = items).
struct ProfRec { const char* name; int kind; double work, bytes; cudaEvent_t e0, e1; };
struct ProfAgg { int kind = 0; long long launches = 0; double ms = 0, work = 0, bytes = 0; };
struct Profiler {
  bool on = false;
  bool open = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, ProfAgg> agg;
};

struct modsgpu_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;          // side stream of the detector's fork (detect.cu); events of the fork / join
  cudaEvent_t det_fork_ev = nullptr, det_join_ev = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // detector launch sequences captured as CUDA graphs, keyed by everything a launch argument depends on (mg_detect_graph)
  struct DetGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; int uses = 0; };
  std::map<std::string, DetGraph> det_graphs;
  HostBuf h_out;                   // pinned landing area of result read-backs (mg_read_back)
  cudaEvent_t ev_sync = nullptr;   // cudaEventBlockingSync: host waits sleep instead of spinning (mg_stream_sync)
  float last_ms = 0.f;
  long long launches = 0;
  std::string err;
  // workspaces (named by their user)
  DevBuf det_pyr, det_cand, det_map, det_out, det_misc, det_aff;
  HostBuf h_stage, h_stage2;
  DevBuf io_a, io_b, io_c;            // generic staging for the test-only entry points
  DevBuf smp_regs, smp_meta, smp_taps, smp_scratch, smp_out;
  DevBuf smp_bins;                    // histogram / bin-base scratch of the device-side sampler preparation
  DevBuf smp_prof;                    // 6 doubles: algorithmic bytes per sampler class of the device-prepared launches (profiler)
  DevBuf smp_taptab;                  // Gaussian taps of every even window size (patchSize 32), device-side sampler prep
  DevBuf chain_tmp;                   // uncompacted rows + verdict bytes between a net and the compaction
  DevBuf chain_a, chain_b, chain_misc;   // region lists (DevRegion) of the per-view chain, counters / stats
  DevBuf cnn_act0, cnn_act1, cnn_out;
  DevBuf cnn_stats;                   // normalised patches (fp32) for the fused conv1+conv2 kernel
  DevBuf mt_q, mt_t, mt_d, mt_aux, mt_out;
  DevBuf rs_buf;
  NetWeights* nets[3] = {nullptr, nullptr, nullptr};
  // second context of the same device for the second image of a pair (modsgpu_ctx_sibling): own stream and workspaces, the
  // weights of the nets shared with this context (mg_nets_share: the sibling owns only its activation buffers)
  modsgpu_ctx* sibling = nullptr;
  cudaEvent_t sib_ev = nullptr;
  bool nets_borrowed = false;
  int pair_overlap = 0;
  Profiler prof;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;   // modsgpu_timer_*
  DevBuf l2flush;
  std::vector<std::pair<float*, size_t>> img_pool;   // recycled image buffers (cudaFree would sync the device)
  std::vector<std::pair<float*, size_t>> desc_pool;  // recycled device descriptor blocks (modsgpu_describe_view_dev)
};

cudaError_t mg_image_alloc(modsgpu_ctx* ctx, size_t bytes, float** out);

int mg_nets_share(modsgpu_ctx* sib, const modsgpu_ctx* src);      // cnn.cu
struct modsgpu_devdesc;
const float* mg_devdesc_ptr(const modsgpu_devdesc* dd);            // chain.cu
void mg_prof_begin(modsgpu_ctx* ctx, const char* name, int kind, double work, double bytes = 0.0);
void mg_prof_end(modsgpu_ctx* ctx);
// call right before a kernel launch; MG_LAUNCHED closes the record
#define MG_PROF(ctx, name, kind, work) do { if ((ctx)->prof.on) mg_prof_begin(ctx, name, kind, work); } while (0)
// flop-counted kernels that also stream activations: `bytes` = algorithmic HBM bytes (inputs read once + outputs written once)
#define MG_PROF2(ctx, name, kind, work, bytes) do { if ((ctx)->prof.on) mg_prof_begin(ctx, name, kind, work, bytes); } while (0)

#define MG_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                   std::to_string(__LINE__) + ")";                                          \
      return MODSGPU_ECUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define MG_FAIL(ctx, code, msg) \
  do { (ctx)->err = (msg); return (code); } while (0)

// after every kernel launch
#define MG_LAUNCHED(ctx)                         \
  do {                                           \
    (ctx)->launches++;                           \
    if ((ctx)->prof.open) mg_prof_end(ctx);      \
    MG_CUDA(ctx, cudaGetLastError());            \
  } while (0)

// Host wait for the context's stream.  The wait sleeps on a blocking-sync event: a process usually drives several
// contexts from several threads (one per pair in flight), often more threads than it has cores (4 cores per GPU on the
// 8-GPU box), and a spinning cudaStreamSynchronize per thread starves the threads that have launches to issue.
// MODSGPU_SPIN_SYNC=1 restores the spinning wait (lowest latency for a single context).
// A process with one or two live contexts keeps the spinning wait (lowest latency: 89 vs 111 ms for the MODS loop of
// one hard pair); with more contexts than that -- one per pair in flight -- the waits sleep.
extern std::atomic<int> mg_live_contexts;   // api.cu
static inline bool mg_sleeping_waits(const modsgpu_ctx* ctx) {
  return ctx->ev_sync != nullptr && mg_live_contexts.load(std::memory_order_relaxed) > 2;
}
static inline cudaError_t mg_stream_sync(modsgpu_ctx* ctx) {
  if (!mg_sleeping_waits(ctx)) return cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->ev_sync);
}

static inline int mg_begin(modsgpu_ctx* ctx) {
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  MG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}
static inline int mg_end(modsgpu_ctx* ctx) {
  // ev1 is the entry point's end marker AND what the host sleeps on (it carries cudaEventBlockingSync unless
  // MODSGPU_SPIN_SYNC=1); the elapsed time is formed only when modsgpu_last_device_ms asks for it
  MG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  MG_CUDA(ctx, mg_sleeping_waits(ctx) ? cudaEventSynchronize(ctx->ev1) : cudaStreamSynchronize(ctx->stream));
  ctx->last_ms = -1.f;
  return 0;
}

// Device -> caller's (pageable) buffer at the end of an entry point: land in pinned memory, wait for the stream with the
// sleeping wait, then memcpy.  (A cudaMemcpyAsync to pageable memory would wait for the preceding kernels inside the
// driver, spinning on a core.)  Includes mg_end.
static inline int mg_read_back_end(modsgpu_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
  MG_CUDA(ctx, ctx->h_out.ensure(bytes));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->h_out.p, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  memcpy(dst, ctx->h_out.p, bytes);
  return 0;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device: a call site keeps one OnceFlags and asks whether this context's device still
// needs the call.  Contexts of several threads may race here; the attribute call is idempotent, the flag is atomic.
struct OnceFlags {
  std::atomic<unsigned long long> done{0};
  bool need(int device) {
    const unsigned long long bit = 1ull << (device & 63);
    return (done.load(std::memory_order_acquire) & bit) == 0;
  }
  void set(int device) { done.fetch_or(1ull << (device & 63), std::memory_order_release); }
};

// Gaussian taps exactly as cv::getGaussianKernel(ksize,(double)sigma,CV_32F) produces them for the
// ksize rule of helpers.cpp:717-731: ksize=(int)(2*3*sigma+1), forced odd.  Returns ksize.
int mg_gaussian_taps(float sigma, std::vector<float>& taps);
