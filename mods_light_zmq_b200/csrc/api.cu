// api.cu -- context life cycle of libmodsgpu.so
#include "common.cuh"
#include <cstdlib>

void mg_free_nets(modsgpu_ctx* ctx);  // cnn.cu

std::atomic<int> mg_live_contexts{0};

extern "C" const char* modsgpu_version(void) { return "modsgpu 0.1 (sm_100a)"; }

extern "C" int modsgpu_create(int device, modsgpu_ctx** out) {
  if (!out) return MODSGPU_EINVAL;
  *out = nullptr;
  // One hardware work queue per stream instead of the default 8 shared ones: with 16 contexts (32 streams) per GPU the
  // default made unrelated streams share queues -- no throughput change, but 24 % more host CPU per pair (launches wait
  // behind another stream's queue).  Only effective if no CUDA context exists yet; never overrides the caller's value.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MODSGPU_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MODSGPU_ENODEV;
  if (prop.major != 10) return MODSGPU_ENODEV;  // sm_100a code only; there is no other path
  if (cudaSetDevice(device) != cudaSuccess) return MODSGPU_ENODEV;
  modsgpu_ctx* ctx = new modsgpu_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  const char* spin = getenv("MODSGPU_SPIN_SYNC");
  const bool spinning = spin && atoi(spin) != 0;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev1, spinning ? cudaEventDefault : cudaEventBlockingSync) != cudaSuccess ||
      (!spinning &&
       cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)) {
    // whatever was created so far goes with the context (nothing else holds these handles yet)
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return MODSGPU_ECUDA;
  }
  {
    const char* e = getenv("MODSGPU_PAIR_OVERLAP");
    ctx->pair_overlap = e ? atoi(e) : 0;
  }
  mg_live_contexts.fetch_add(1);
  *out = ctx;
  return 0;
}

// ---- two images of a pair side by side ----------------------------------------------------------------------------
// The reference extracts the two images of a pair concurrently (OpenMP tasks, mods.cpp:234-251).  One context is one
// stream and one set of workspaces, so the pair-level entry points can borrow a SIBLING context for the second image:
// same device, own stream / workspaces / detector graphs, the nets shared.  Off by default (with many contexts per GPU
// the device is already full and the sibling only costs memory); MODSGPU_PAIR_OVERLAP=1 or modsgpu_set_pair_overlap.
extern "C" int modsgpu_set_pair_overlap(modsgpu_ctx* ctx, int on) {
  if (!ctx) return MODSGPU_EINVAL;
  ctx->pair_overlap = on ? 1 : 0;
  return 0;
}
extern "C" int modsgpu_get_pair_overlap(const modsgpu_ctx* ctx) { return ctx ? ctx->pair_overlap : 0; }

// The sibling (created on first use, owned by ctx).  Its stream is ordered after everything enqueued on ctx's stream so
// far -- images converted on ctx's stream may be read by the sibling straight away.
extern "C" int modsgpu_ctx_sibling(modsgpu_ctx* ctx, modsgpu_ctx** out) {
  if (!ctx || !out) return MODSGPU_EINVAL;
  *out = nullptr;
  if (ctx->nets_borrowed) MG_FAIL(ctx, MODSGPU_ESTATE, "a sibling context has no sibling of its own");
  if (!ctx->sibling) {
    modsgpu_ctx* sib = nullptr;
    const int rc = modsgpu_create(ctx->device, &sib);
    if (rc) MG_FAIL(ctx, rc, "sibling context could not be created");
    sib->nets_borrowed = true;
    sib->pair_overlap = 0;
    ctx->sibling = sib;
    MG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->sib_ev, cudaEventDisableTiming));
  }
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = mg_nets_share(ctx->sibling, ctx)) MG_FAIL(ctx, rc, ctx->sibling->err);
  MG_CUDA(ctx, cudaEventRecord(ctx->sib_ev, ctx->stream));
  MG_CUDA(ctx, cudaStreamWaitEvent(ctx->sibling->stream, ctx->sib_ev, 0));
  *out = ctx->sibling;
  return 0;
}
// after the sibling's work has been collected: its launches count as ctx's, its error (if any) becomes ctx's
extern "C" int modsgpu_ctx_sibling_join(modsgpu_ctx* ctx) {
  if (!ctx || !ctx->sibling) return MODSGPU_EINVAL;
  ctx->launches += ctx->sibling->launches;
  ctx->sibling->launches = 0;
  if (!ctx->sibling->err.empty()) { ctx->err = ctx->sibling->err; ctx->sibling->err.clear(); }
  return 0;
}

extern "C" void modsgpu_destroy(modsgpu_ctx* ctx) {
  if (!ctx) return;
  if (ctx->sibling) {
    modsgpu_destroy(ctx->sibling);      // frees its activation buffers only (the weights are ctx's)
    ctx->sibling = nullptr;
  }
  mg_live_contexts.fetch_sub(1);
  cudaSetDevice(ctx->device);
  mg_stream_sync(ctx);
  if (ctx->sib_ev) cudaEventDestroy(ctx->sib_ev);
  mg_free_nets(ctx);
  DevBuf* bufs[] = {&ctx->det_pyr, &ctx->det_cand, &ctx->det_map, &ctx->det_out, &ctx->det_misc, &ctx->det_aff, &ctx->io_a, &ctx->io_b,
                    &ctx->io_c, &ctx->smp_regs, &ctx->smp_meta, &ctx->smp_taps, &ctx->smp_scratch, &ctx->smp_out,
                    &ctx->cnn_act0, &ctx->cnn_act1, &ctx->cnn_out, &ctx->mt_q, &ctx->mt_t, &ctx->mt_d, &ctx->mt_aux,
                    &ctx->mt_out, &ctx->rs_buf, &ctx->cnn_stats, &ctx->smp_bins, &ctx->smp_prof, &ctx->smp_taptab, &ctx->chain_a, &ctx->chain_b, &ctx->chain_misc, &ctx->chain_tmp};
  for (DevBuf* b : bufs) b->release();
  ctx->h_stage.release();
  ctx->h_stage2.release();
  ctx->h_out.release();
  for (auto& g : ctx->det_graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  for (auto e : ctx->prof.pool) cudaEventDestroy(e);
  for (auto& r : ctx->prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  ctx->l2flush.release();
  for (auto& b : ctx->img_pool) cudaFree(b.first);
  for (auto& b : ctx->desc_pool) cudaFree(b.first);
  if (ctx->tm0) cudaEventDestroy(ctx->tm0);
  if (ctx->tm1) cudaEventDestroy(ctx->tm1);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
  if (ctx->det_fork_ev) cudaEventDestroy(ctx->det_fork_ev);
  if (ctx->det_join_ev) cudaEventDestroy(ctx->det_join_ev);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* modsgpu_last_error(const modsgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" float modsgpu_last_device_ms(const modsgpu_ctx* ctx) {
  if (!ctx) return 0.f;
  float ms = ctx->last_ms;
  if (ms < 0.f && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
  return ms;
}
extern "C" long long modsgpu_launch_count(const modsgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* modsgpu_stream(const modsgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- measurement helpers (bench.py) -------------------------------------------------------------------
void mg_prof_begin(modsgpu_ctx* ctx, const char* name, int kind, double work, double bytes) {
  Profiler& P = ctx->prof;
  ProfRec r;
  r.name = name; r.kind = kind; r.work = work; r.bytes = bytes;
  for (cudaEvent_t* e : {&r.e0, &r.e1}) {
    if (!P.pool.empty()) { *e = P.pool.back(); P.pool.pop_back(); }
    else cudaEventCreate(e);
  }
  cudaEventRecord(r.e0, ctx->stream);
  P.recs.push_back(r);
  P.open = true;
}
void mg_prof_end(modsgpu_ctx* ctx) {
  Profiler& P = ctx->prof;
  if (!P.open || P.recs.empty()) return;
  cudaEventRecord(P.recs.back().e1, ctx->stream);
  P.open = false;
}
static void prof_collect(modsgpu_ctx* ctx) {
  Profiler& P = ctx->prof;
  mg_stream_sync(ctx);
  for (auto& r : P.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      ProfAgg& a = P.agg[r.name];
      a.kind = r.kind; a.launches++; a.ms += ms; a.work += r.work; a.bytes += r.bytes;
    }
    P.pool.push_back(r.e0); P.pool.push_back(r.e1);
  }
  P.recs.clear();
  P.open = false;
}
extern "C" int modsgpu_profile_enable(modsgpu_ctx* ctx, int on) {
  if (!ctx) return MODSGPU_EINVAL;
  prof_collect(ctx);
  if (on) {
    ctx->prof.agg.clear();
    if (ctx->smp_prof.p) cudaMemsetAsync(ctx->smp_prof.p, 0, 64, ctx->stream);
  }
  ctx->prof.on = on != 0;
  return 0;
}
// JSON: {"kernel": {"kind": k, "launches": n, "ms": total, "work": total, "bytes": total (flop-kind kernels)}, ...}; returns the length needed
extern "C" int modsgpu_profile_report(modsgpu_ctx* ctx, char* buf, int cap) {
  if (!ctx) return MODSGPU_EINVAL;
  prof_collect(ctx);
  if (ctx->smp_prof.p) {     // sampler launches prepared on the device: their algorithmic bytes were summed there (sampler.cu)
    double acc[6] = {0, 0, 0, 0, 0, 0};
    static const char* names[6] = {"k_sample_small", "k_sample_a<R<=40>", "k_sample_a<R<=65>", "k_sample_b<R<=100>", "k_sample_b<R<=160>",
                                   "k_large_resample"};
    if (cudaMemcpy(acc, ctx->smp_prof.p, sizeof(acc), cudaMemcpyDeviceToHost) == cudaSuccess) {
      for (int c = 0; c < 6; c++) {
        auto it = ctx->prof.agg.find(names[c]);
        if (it != ctx->prof.agg.end() && acc[c] > 0) it->second.work += acc[c];
      }
      cudaMemset(ctx->smp_prof.p, 0, sizeof(acc));
    }
  }
  std::string s = "{";
  bool first = true;
  for (auto& kv : ctx->prof.agg) {
    char tmp[512];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"kind\": %d, \"launches\": %lld, \"ms\": %.6f, \"work\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.kind, kv.second.launches, kv.second.ms, kv.second.work, kv.second.bytes);
    s += tmp;
    first = false;
  }
  s += "}";
  if (buf && cap > 0) { strncpy(buf, s.c_str(), cap - 1); buf[cap - 1] = 0; }
  return (int)s.size() + 1;
}
extern "C" int modsgpu_timer_start(modsgpu_ctx* ctx) {
  if (!ctx) return MODSGPU_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->tm0) { MG_CUDA(ctx, cudaEventCreate(&ctx->tm0)); MG_CUDA(ctx, cudaEventCreate(&ctx->tm1)); }
  MG_CUDA(ctx, mg_stream_sync(ctx));
  MG_CUDA(ctx, cudaEventRecord(ctx->tm0, ctx->stream));
  return 0;
}
extern "C" int modsgpu_timer_stop(modsgpu_ctx* ctx, float* ms) {
  if (!ctx || !ms || !ctx->tm0) return MODSGPU_EINVAL;
  MG_CUDA(ctx, cudaEventRecord(ctx->tm1, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  MG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1));
  return 0;
}
// write a buffer larger than the 126 MB L2 so the next step starts cold
extern "C" int modsgpu_flush_l2(modsgpu_ctx* ctx) {
  if (!ctx) return MODSGPU_EINVAL;
  const size_t bytes = (size_t)256 << 20;
  MG_CUDA(ctx, ctx->l2flush.ensure(bytes));
  MG_CUDA(ctx, cudaMemsetAsync(ctx->l2flush.p, 0, bytes, ctx->stream));
  return 0;
}
