// chain.cu -- one synthesised view from pixels to described regions without leaving the device.
//
// Replaces the per-view body of ImageRepresentation::SynthDetectDescribeKeypoints (imagerepresentation.cpp:704-1006,
// HessianAffine + AffNet + OriNet + HardNet++):
//   DetectAffineRegions (:739)                          -> detector graph, keypoints stay in ctx->det_out
//   DescribeWithZmq(AffNet) + post-processing (:797-845) -> sampler, AffNet, k_chain_affnet_apply + k_chain_compact
//   ReprojectRegionsAndRemoveTouchBoundary (:868, synth-detection.cpp:151-190, dontRemove)   (same kernel)
//   DescribeWithZmq(OriNet) + rotation (:876-899)        -> sampler, OriNet, k_chain_orinet_apply + k_chain_compact
//   ReprojectRegions (:951, synth-detection.cpp:631-706)                                     (same kernel)
//   DescribeWithZmq(desc) (:992-1006)                    -> sampler, HardNet++
// The seam-by-seam route (modsgpu_detect + 3 x modsgpu_describe with the host arithmetic of mods_host.cpp in between)
// needs four host round trips per view and rebuilds the sampler's work lists on the host three times; here the region
// list is filtered and compacted (order preserving) by one CTA between the nets, the nets and the sampler take their
// live counts from device memory, and the host synchronises twice: once after the detector for the launch bounds
// (count + SmpStats, 120 bytes) and once for the result.
//
// Arithmetic: the fp64 / fp32 mix of the host mirror, operation by operation (compiled --fmad=false).  The only
// functions that are not IEEE basic operations are atan2 / cos / sin of the OriNet rotation: CUDA's double versions may
// differ from glibc's in the last bit (<= 2 ulp), which no decision of the chain can see except through a float cast
// sitting exactly on a rounding boundary; tests compare the chain with the seam route at 4e-16 relative on A.
#include "common.cuh"
#include <cmath>
#include <algorithm>

int mg_detect_device(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p, int cap);
int mg_keys_to_export(const modsgpu_keypoint* k, int n, const modsgpu_pyr_params* p);
int mg_sample_stats_enqueue(modsgpu_ctx* ctx, const DevRegion* regs, const int* cnt_dev, double mrSize, SmpStats* d_stats);
int mg_sample_enqueue_dev(modsgpu_ctx* ctx, const modsgpu_image* img, const DevRegion* regs, const int* cnt_dev, int n_ub,
                          const SmpStats& st, double mrSize, uint8_t* d_out);
int mg_net_forward_enqueue(modsgpu_ctx* ctx, modsgpu_net net, const uint8_t* d_patches, int n, float* d_out, const int* cnt_dev);

namespace {

struct Mat3 { double m[9]; };

// helpers.cpp:524-549 (callers pass doubles into the int res_w / res_h: truncation)
__device__ __forceinline__ bool check_borders(int img_w, int img_h, float ofsx, float ofsy, float a11, float a12, float a21,
                                              float a22, int res_w, int res_h) {
  const int width = img_w - 2, height = img_h - 2;
  const float halfWidth = (float)ceil((double)((float)res_w) / 2.0);
  const float halfHeight = (float)ceil((double)((float)res_h) / 2.0);
  const float x[4] = {-halfWidth, -halfWidth, +halfWidth, +halfWidth};
  const float y[4] = {-halfHeight, +halfHeight, -halfHeight, +halfHeight};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float imx = ofsx + x[i] * a11 + y[i] * a12;
    const float imy = ofsy + x[i] * a21 + y[i] * a22;
    if (floorf(imx) <= 0 || floorf(imy) <= 0 || ceilf(imx) >= (float)width || ceilf(imy) >= (float)height) return true;
  }
  return false;
}

// synth-detection.cpp:578-587 on a copy of det_kp
__device__ __forceinline__ void reproject(const modsgpu_region& in, modsgpu_region& out, const Mat3& H, int eye) {
  out = in;
  if (eye) return;
  out.x = (H.m[0] * in.x + H.m[1] * in.y + H.m[2]);
  out.y = (H.m[3] * in.x + H.m[4] * in.y + H.m[5]);
  out.a11 = (H.m[0] * in.a11 + H.m[1] * in.a21);
  out.a12 = (H.m[0] * in.a12 + H.m[1] * in.a22);
  out.a21 = (H.m[3] * in.a11 + H.m[4] * in.a21);
  out.a22 = (H.m[3] * in.a12 + H.m[4] * in.a22);
}
__device__ __forceinline__ bool centre_inside(const modsgpu_region& p, int orig_w, int orig_h) {
  return (p.x < orig_w) && (p.y < orig_h) && (p.x > 0) && (p.y > 0);
}

// imagerepresentation.cpp:803-845 for one region: A from the three AffNet outputs, rectifyAffineTransformationUpIsUp
// (helpers.cpp:401-410), getEigenvalues (:504-515) ratio test, frame test against the VIEW; then the centre test of
// ReprojectRegionsAndRemoveTouchBoundary against the ORIGINAL image.  keepA: survives AffNet's tests (n_affine counts
// these); keepB: also inside the original image.
__device__ __forceinline__ void affnet_apply(const DevRegion& r, const float* __restrict__ o, int w, int h, int orig_w, int orig_h,
                                             double mrSize, const Mat3& Hinv, int eye, DevRegion& t, bool& keepA, bool& keepB) {
  t = r;
  const double a = (double)o[0], b = 0.0, c = (double)o[1], d = (double)o[2];
  const double det = sqrt(fabs(a * d - b * c));
  const double b2a2 = sqrt(b * b + a * a);
  t.det.a11 = b2a2 / det;
  t.det.a12 = 0;
  t.det.a21 = (d * b + c * a) / (b2a2 * det);
  t.det.a22 = det / b2a2;
  const float fa = (float)t.det.a11, fb = (float)t.det.a12, fc = (float)t.det.a21, fd = (float)t.det.a22;
  const float trace = fa + fd;
  const float delta1 = (trace * trace - 4 * (fa * fd - fb * fc));
  bool keep = !(delta1 < 0);
  if (keep) {
    const float delta = sqrtf(delta1);
    const float l1 = (trace + delta) / 2.0f, l2 = (trace - delta) / 2.0f;
    keep = !((l1 / l2 > 6) || (l2 / l1 > 6));
  }
  const int fs = (int)(mrSize * t.det.s);
  keep = keep && !check_borders(w, h, (float)t.det.x, (float)t.det.y, fa, fb, fc, fd, fs, fs);
  keepA = keep;
  reproject(t.det, t.reproj, Hinv, eye);
  keepB = keep && centre_inside(t.reproj, orig_w, orig_h);
}

// imagerepresentation.cpp:881-899 (A <- A * R(angle), angle = atan2(o0, o1)), then ReprojectRegions
// (synth-detection.cpp:631-706): centre + k_sigma * s frame inside the ORIGINAL image
__device__ __forceinline__ void orinet_apply(const DevRegion& r, const float* __restrict__ o, int orig_w, int orig_h, double k_sigma,
                                             const Mat3& Hinv, int eye, DevRegion& t, bool& keep) {
  t = r;
  const double angle = atan2((double)o[0], (double)o[1]);
  const double ci = cos(angle), si = sin(angle);
  const double a11 = r.det.a11, a12 = r.det.a12, a21 = r.det.a21, a22 = r.det.a22;
  t.det.a11 = a11 * ci - a12 * si;
  t.det.a12 = a11 * si + a12 * ci;
  t.det.a21 = a21 * ci - a22 * si;
  t.det.a22 = a21 * si + a22 * ci;
  reproject(t.det, t.reproj, Hinv, eye);
  const modsgpu_region& p = t.reproj;
  const int fs = (int)(k_sigma * p.s);
  keep = centre_inside(p, orig_w, orig_h) &&
         !check_borders(orig_w, orig_h, (float)p.x, (float)p.y, (float)p.a11, (float)p.a12, (float)p.a21, (float)p.a22, fs, fs);
}

// exclusive prefix of one int per thread over a 1024-thread CTA; returns the total through `total`
__device__ __forceinline__ int block_excl_scan(int v, int& total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    warp_sums[lane] = wi - w;
    if (lane == 31) total = wi;
  }
  __syncthreads();
  const int r = warp_sums[warp] + incl - v;
  __syncthreads();
  return r;
}

// keypoints of the detector -> region rows: DetectAffineRegions (synth-detection.hpp:79-112) with A = I
__global__ void k_chain_init(const modsgpu_keypoint* __restrict__ kp, const int* __restrict__ det_counters, int cap,
                             DevRegion* __restrict__ out, int* __restrict__ cnt) {
  const int n = min(det_counters[1], cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) cnt[0] = n;
  if (i >= n) return;
  const modsgpu_keypoint k = kp[i];
  DevRegion r;
  r.det.x = k.x; r.det.y = k.y; r.det.s = k.s;
  r.det.a11 = 1; r.det.a12 = 0; r.det.a21 = 0; r.det.a22 = 1;
  r.reproj = r.det;
  r.response = k.response; r.octave = k.octave; r.type = k.type;
  out[i] = r;
}

// Post-processing of a net's output in two launches: the per-region arithmetic runs on as many CTAs as the list needs
// (k_chain_*_apply: new row into tmp[i], verdict into flag[i]); one CTA then compacts the survivors in list order
// (k_chain_compact: thread t owns the contiguous run [t*per, (t+1)*per) of the flags; rows are copied 16 bytes per lane).
// flag bits: 1 = counted in cnt_out[0] (AffNet: survives AffNet's own tests, ImageRepresentation::n_affine), 2 = kept.
__global__ void __launch_bounds__(128)
k_chain_affnet_apply(const DevRegion* __restrict__ in, const float* __restrict__ aff, const int* __restrict__ cnt_in,
                     DevRegion* __restrict__ tmp, unsigned char* __restrict__ flag, int w, int h, int orig_w, int orig_h, double mrSize,
                     Mat3 Hinv, int eye) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *cnt_in) return;
  DevRegion t; bool ka, kb;
  affnet_apply(in[i], aff + 3 * (size_t)i, w, h, orig_w, orig_h, mrSize, Hinv, eye, t, ka, kb);
  tmp[i] = t;
  flag[i] = (unsigned char)((ka ? 1 : 0) | (kb ? 2 : 0));
}
__global__ void __launch_bounds__(128)
k_chain_orinet_apply(const DevRegion* __restrict__ in, const float* __restrict__ ori, const int* __restrict__ cnt_in,
                     DevRegion* __restrict__ tmp, unsigned char* __restrict__ flag, int orig_w, int orig_h, double k_sigma, Mat3 Hinv, int eye) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *cnt_in) return;
  DevRegion t; bool k;
  orinet_apply(in[i], ori + 2 * (size_t)i, orig_w, orig_h, k_sigma, Hinv, eye, t, k);
  tmp[i] = t;
  flag[i] = (unsigned char)(k ? 3 : 0);
}
// cnt_out[0] = #(flag & 1), cnt_out[n_cnt - 1] = #(flag & 2) = length of the compacted list (n_cnt = 2: both counters)
__global__ void __launch_bounds__(1024)
k_chain_compact(const DevRegion* __restrict__ tmp, const unsigned char* __restrict__ flag, const int* __restrict__ cnt_in,
                DevRegion* __restrict__ out, int* __restrict__ cnt_out, int n_cnt) {
  __shared__ int s_total, s_totalA;
  __shared__ int s_src[8192];          // source index of every survivor of the current 8192-row window
  const int n = *cnt_in;
  int done = 0, doneA = 0;
  for (int base = 0; base < n; base += 8192) {
    const int m = min(8192, n - base);
    const int per = (m + 1023) / 1024, i0 = threadIdx.x * per, i1 = min(m, i0 + per);
    int nA = 0, nB = 0;
    for (int i = i0; i < i1; i++) { const int f = flag[base + i]; nA += f & 1; nB += (f >> 1) & 1; }
    block_excl_scan(nA, s_totalA);
    int pos = block_excl_scan(nB, s_total);
    for (int i = i0; i < i1; i++) if (flag[base + i] & 2) s_src[pos++] = base + i;
    __syncthreads();
    const int kept = s_total;
    // rows of 128 bytes: 8 lanes x 16 bytes per row, coalesced
    const uint4* src4 = reinterpret_cast<const uint4*>(tmp);
    uint4* dst4 = reinterpret_cast<uint4*>(out);
    for (int it = threadIdx.x; it < kept * 8; it += 1024) {
      const int k = it >> 3, q = it & 7;
      dst4[(size_t)(done + k) * 8 + q] = src4[(size_t)s_src[k] * 8 + q];
    }
    done += kept; doneA += s_totalA;
    __syncthreads();
  }
  if (threadIdx.x == 0) { cnt_out[0] = doneA; cnt_out[n_cnt - 1] = done; }
}

bool invert3h(const double* A, double* R) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0 || !std::isfinite(det)) { for (int i = 0; i < 9; i++) R[i] = 0; return false; }
  const double id = 1.0 / det;
  R[0] = c0 * id; R[1] = (A[2] * A[7] - A[1] * A[8]) * id; R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c1 * id; R[4] = (A[0] * A[8] - A[2] * A[6]) * id; R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c2 * id; R[7] = (A[1] * A[6] - A[0] * A[7]) * id; R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}
// synth-detection.cpp:143-149
bool h_is_eye(const double* H) {
  double s = 0;
  for (int i = 0; i < 9; i++) s += std::fabs(H[i] - ((i % 4 == 0) ? 1.0 : 0.0));
  return s < 0.01;
}

// layout of ctx->chain_misc: [0..15] ints: cnt[0] keypoints, [1] n_affine, [2] after the centre test, [3] described;
// SmpStats at byte 64
constexpr size_t MISC_STATS_OFF = 64, MISC_BYTES = 64 + sizeof(SmpStats) + 64;

}  // namespace

// test-only seams of the two post-processing kernels: host arrays in, survivors out (order preserved).
// regions: n rows of modsgpu_view_region (det filled; reproj ignored on input).
extern "C" int modsgpu_debug_affnet_post(modsgpu_ctx* ctx, const modsgpu_view_region* regs, const float* aff, int n, int w, int h,
                                         int orig_w, int orig_h, double mrSize, const double* H, modsgpu_view_region* out,
                                         int* n_affine, int* n_out) {
  if (!ctx || (n > 0 && (!regs || !aff || !out)) || n < 0 || !n_affine || !n_out) return MODSGPU_EINVAL;
  static_assert(sizeof(modsgpu_view_region) == sizeof(DevRegion), "view region layout");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  Mat3 Hinv; const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  invert3h(H ? H : I3, Hinv.m);
  const int eye = h_is_eye(H ? H : I3);
  MG_CUDA(ctx, ctx->chain_a.ensure((size_t)(n + 1) * sizeof(DevRegion)));
  MG_CUDA(ctx, ctx->chain_b.ensure((size_t)(n + 1) * sizeof(DevRegion)));
  MG_CUDA(ctx, ctx->cnn_out.ensure((size_t)(n + 1) * 3 * 4));
  MG_CUDA(ctx, ctx->chain_misc.ensure(MISC_BYTES));
  int* cnt = ctx->chain_misc.as<int>();
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->chain_a.p, regs, (size_t)n * sizeof(DevRegion), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->cnn_out.p, aff, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(cnt, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, ctx->chain_tmp.ensure((size_t)(n + 1) * (sizeof(DevRegion) + 1)));
  {
    DevRegion* tmp = ctx->chain_tmp.as<DevRegion>();
    unsigned char* flag = reinterpret_cast<unsigned char*>(tmp + n + 1);
    k_chain_affnet_apply<<<ceil_div(std::max(n, 1), 128), 128, 0, ctx->stream>>>(ctx->chain_a.as<DevRegion>(), ctx->cnn_out.as<float>(), cnt, tmp, flag,
                                                                                 w, h, orig_w, orig_h, mrSize, Hinv, eye);
    MG_LAUNCHED(ctx);
    k_chain_compact<<<1, 1024, 0, ctx->stream>>>(tmp, flag, cnt, ctx->chain_b.as<DevRegion>(), cnt + 1, 2);
    MG_LAUNCHED(ctx);
  }
  int hc[3] = {0, 0, 0};
  MG_CUDA(ctx, cudaMemcpyAsync(hc, cnt, 12, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  *n_affine = hc[1]; *n_out = hc[2];
  if (hc[2] > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->chain_b.p, (size_t)hc[2] * sizeof(DevRegion), cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

extern "C" int modsgpu_debug_orinet_post(modsgpu_ctx* ctx, const modsgpu_view_region* regs, const float* ori, int n, int orig_w,
                                         int orig_h, const double* H, modsgpu_view_region* out, int* n_out) {
  if (!ctx || (n > 0 && (!regs || !ori || !out)) || n < 0 || !n_out) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  Mat3 Hinv; const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  invert3h(H ? H : I3, Hinv.m);
  const int eye = h_is_eye(H ? H : I3);
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);   // synth-detection.cpp:21
  MG_CUDA(ctx, ctx->chain_a.ensure((size_t)(n + 1) * sizeof(DevRegion)));
  MG_CUDA(ctx, ctx->chain_b.ensure((size_t)(n + 1) * sizeof(DevRegion)));
  MG_CUDA(ctx, ctx->cnn_out.ensure((size_t)(n + 1) * 2 * 4));
  MG_CUDA(ctx, ctx->chain_misc.ensure(MISC_BYTES));
  int* cnt = ctx->chain_misc.as<int>();
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->chain_a.p, regs, (size_t)n * sizeof(DevRegion), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->cnn_out.p, ori, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(cnt, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, ctx->chain_tmp.ensure((size_t)(n + 1) * (sizeof(DevRegion) + 1)));
  {
    DevRegion* tmp = ctx->chain_tmp.as<DevRegion>();
    unsigned char* flag = reinterpret_cast<unsigned char*>(tmp + n + 1);
    k_chain_orinet_apply<<<ceil_div(std::max(n, 1), 128), 128, 0, ctx->stream>>>(ctx->chain_a.as<DevRegion>(), ctx->cnn_out.as<float>(), cnt, tmp, flag,
                                                                                 orig_w, orig_h, k_sigma, Hinv, eye);
    MG_LAUNCHED(ctx);
    k_chain_compact<<<1, 1024, 0, ctx->stream>>>(tmp, flag, cnt, ctx->chain_b.as<DevRegion>(), cnt + 1, 1);
    MG_LAUNCHED(ctx);
  }
  int hc[2] = {0, 0};
  MG_CUDA(ctx, cudaMemcpyAsync(hc, cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  *n_out = hc[1];
  if (hc[1] > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->chain_b.p, (size_t)hc[1] * sizeof(DevRegion), cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

// The whole view.  *regions (n rows) and *desc (n x 128 floats holding integers 0..255) are malloc()ed -> modsgpu_free.
// counts: [0] raw keypoints, [1] regions after AffNet's eigen-ratio / frame tests, [2] described regions (= n).
// Descriptor block left on the device (modsgpu_describe_view_dev): n x dim floats, buffers recycled through the context
// (cudaFree would synchronise the device).
struct modsgpu_devdesc { float* d = nullptr; int n = 0, dim = 128; size_t cap = 0; };

static cudaError_t devdesc_alloc(modsgpu_ctx* ctx, size_t bytes, modsgpu_devdesc* dd) {
  bytes = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);       // 1 MB granules: same-sized views hit the pool
  for (size_t i = 0; i < ctx->desc_pool.size(); i++)
    if (ctx->desc_pool[i].second >= bytes && ctx->desc_pool[i].second <= 2 * bytes) {
      dd->d = ctx->desc_pool[i].first; dd->cap = ctx->desc_pool[i].second;
      ctx->desc_pool.erase(ctx->desc_pool.begin() + i);
      return cudaSuccess;
    }
  dd->cap = bytes;
  return cudaMalloc((void**)&dd->d, bytes);
}
extern "C" void modsgpu_devdesc_free(modsgpu_ctx* ctx, modsgpu_devdesc* dd) {
  if (!dd) return;
  if (dd->d) {
    if (ctx && ctx->desc_pool.size() < 8) ctx->desc_pool.emplace_back(dd->d, dd->cap);
    else cudaFree(dd->d);
  }
  delete dd;
}
extern "C" int modsgpu_devdesc_size(const modsgpu_devdesc* dd) { return dd ? dd->n : 0; }
const float* mg_devdesc_ptr(const modsgpu_devdesc* dd) { return dd ? dd->d : nullptr; }
// n x 128 floats to the host (tests; the product path never needs them there)
extern "C" int modsgpu_devdesc_download(modsgpu_ctx* ctx, const modsgpu_devdesc* dd, float* out) {
  if (!ctx || !dd || (dd->n > 0 && !out)) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (dd->n > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, dd->d, (size_t)dd->n * dd->dim * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

static int describe_view_impl(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                              const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                              float** desc, modsgpu_devdesc** devdesc, int* n, int* counts);

extern "C" int modsgpu_describe_view(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                                     const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                                     float** desc, int* n, int* counts) {
  if (!desc) return MODSGPU_EINVAL;
  return describe_view_impl(ctx, view, H, orig_w, orig_h, p, mrSize, patchSize, regions, desc, nullptr, n, counts);
}
// The same with the descriptors LEFT ON THE DEVICE (*devdesc -> modsgpu_devdesc_free): what a caller that only matches them
// (modsgpu_match_fginn_dev) wants -- 0.5 MB instead of 2.7 MB read back per 4k-keypoint view, nothing uploaded again.
extern "C" int modsgpu_describe_view_dev(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                                         const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                                         modsgpu_devdesc** devdesc, int* n, int* counts) {
  if (!devdesc) return MODSGPU_EINVAL;
  return describe_view_impl(ctx, view, H, orig_w, orig_h, p, mrSize, patchSize, regions, nullptr, devdesc, n, counts);
}

static int describe_view_impl(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                              const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                              float** desc, modsgpu_devdesc** devdesc, int* n, int* counts) {
  if (!ctx || !view || !p || !regions || !n) return MODSGPU_EINVAL;
  if (devdesc) *devdesc = nullptr;
  if (patchSize != 32) MG_FAIL(ctx, MODSGPU_EINVAL, "the networks take 32x32 patches");
  if (p->detectorMode < MODSGPU_FIXED_TH || p->detectorMode > MODSGPU_NOT_LESS_THAN_REGIONS)
    MG_FAIL(ctx, MODSGPU_EINVAL, "unknown detectorMode");
  for (int i = 0; i < 3; i++) if (!ctx->nets[i]) MG_FAIL(ctx, MODSGPU_ESTATE, "modsgpu_load_weights has not been called for every net");
  *regions = nullptr; *n = 0;
  if (desc) *desc = nullptr;
  if (counts) counts[0] = counts[1] = counts[2] = 0;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  Mat3 Hinv;
  invert3h(H ? H : I3, Hinv.m);
  const int eye = h_is_eye(H ? H : I3);
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);   // synth-detection.cpp:21
  MG_CUDA(ctx, ctx->chain_misc.ensure(MISC_BYTES));
  int* cnt = ctx->chain_misc.as<int>();
  SmpStats* d_stats = reinterpret_cast<SmpStats*>(ctx->chain_misc.as<uint8_t>() + MISC_STATS_OFF);
  MG_CUDA(ctx, ctx->h_stage.ensure(256));
  int* hc = ctx->h_stage.as<int>();                                   // [0..1] detector counters, [2..5] chain counters
  SmpStats* hst = reinterpret_cast<SmpStats*>(ctx->h_stage.as<uint8_t>() + 64);
  int cap = 1 << 16, n0 = 0;
  for (;;) {
    int rc = mg_detect_device(ctx, view, p, cap);
    if (rc) return rc;
    MG_CUDA(ctx, ctx->chain_a.ensure((size_t)cap * sizeof(DevRegion)));
    MG_CUDA(ctx, ctx->chain_b.ensure((size_t)cap * sizeof(DevRegion)));
    k_chain_init<<<ceil_div(cap, 256), 256, 0, ctx->stream>>>(ctx->det_out.as<modsgpu_keypoint>(), ctx->det_misc.as<int>(), cap,
                                                              ctx->chain_a.as<DevRegion>(), cnt);
    MG_LAUNCHED(ctx);
    rc = mg_sample_stats_enqueue(ctx, ctx->chain_a.as<DevRegion>(), cnt, mrSize, d_stats);
    if (rc) return rc;
    MG_CUDA(ctx, cudaMemcpyAsync(hc, ctx->det_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, cudaMemcpyAsync(hst, d_stats, sizeof(SmpStats), cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));                                // host round trip 1 of 2
    if (hc[0] > cap) { cap = hc[0] + hc[0] / 8; continue; }           // candidate list overflowed: redo with room
    n0 = hc[1];
    break;
  }
  if (p->detectorMode != MODSGPU_FIXED_TH && n0 > 0) {
    // the other detection modes truncate the |response|-sorted list (prepareKeysForExport, scale-space-detector.hpp:125-198):
    // the rule needs the responses on the host; the statistics of the untruncated list stay valid upper bounds
    std::vector<modsgpu_keypoint> k(n0);
    MG_CUDA(ctx, cudaMemcpyAsync(k.data(), ctx->det_out.p, sizeof(modsgpu_keypoint) * (size_t)n0, cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    n0 = mg_keys_to_export(k.data(), n0, p);
    hc[8] = n0;
    MG_CUDA(ctx, cudaMemcpyAsync(cnt, hc + 8, 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (counts) counts[0] = n0;
  if (n0 <= 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const SmpStats st = *hst;
  DevRegion* ra = ctx->chain_a.as<DevRegion>();
  DevRegion* rb = ctx->chain_b.as<DevRegion>();
  MG_CUDA(ctx, ctx->smp_out.ensure((size_t)n0 * 1024 + 16));
  MG_CUDA(ctx, ctx->cnn_out.ensure((size_t)n0 * 128 * 4 + 16));
  uint8_t* patches = ctx->smp_out.as<uint8_t>();
  float* nout = ctx->cnn_out.as<float>();
  int rc;
  // ---- AffNet
  if ((rc = mg_sample_enqueue_dev(ctx, view, ra, cnt, n0, st, mrSize, patches))) return rc;
  if ((rc = mg_net_forward_enqueue(ctx, MODSGPU_AFFNET, patches, n0, nout, cnt))) return rc;
  MG_CUDA(ctx, ctx->chain_tmp.ensure((size_t)(n0 + 1) * (sizeof(DevRegion) + 1)));
  DevRegion* tmp = ctx->chain_tmp.as<DevRegion>();
  unsigned char* flag = reinterpret_cast<unsigned char*>(tmp + n0 + 1);
  MG_PROF(ctx, "k_chain_affnet_apply", 2, (double)n0);
  k_chain_affnet_apply<<<ceil_div(n0, 128), 128, 0, ctx->stream>>>(ra, nout, cnt, tmp, flag, view->w, view->h, orig_w, orig_h, mrSize, Hinv, eye);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_chain_compact", 2, (double)n0);
  k_chain_compact<<<1, 1024, 0, ctx->stream>>>(tmp, flag, cnt, rb, cnt + 1, 2);
  MG_LAUNCHED(ctx);
  // ---- OriNet
  if ((rc = mg_sample_enqueue_dev(ctx, view, rb, cnt + 2, n0, st, mrSize, patches))) return rc;
  if ((rc = mg_net_forward_enqueue(ctx, MODSGPU_ORINET, patches, n0, nout, cnt + 2))) return rc;
  MG_PROF(ctx, "k_chain_orinet_apply", 2, (double)n0);
  k_chain_orinet_apply<<<ceil_div(n0, 128), 128, 0, ctx->stream>>>(rb, nout, cnt + 2, tmp, flag, orig_w, orig_h, k_sigma, Hinv, eye);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_chain_compact", 2, (double)n0);
  k_chain_compact<<<1, 1024, 0, ctx->stream>>>(tmp, flag, cnt + 2, ra, cnt + 3, 1);
  MG_LAUNCHED(ctx);
  // ---- HardNet++ (its output lands in the block that stays on the device, if the caller asked for one)
  modsgpu_devdesc* dd = nullptr;
  if (devdesc) {
    dd = new modsgpu_devdesc();
    if (devdesc_alloc(ctx, (size_t)n0 * 128 * 4 + 16, dd) != cudaSuccess) {
      cudaGetLastError();
      delete dd;
      MG_FAIL(ctx, MODSGPU_ECUDA, "descriptor block could not be allocated");
    }
    nout = dd->d;
  }
  auto fail = [&](int code) { if (dd) modsgpu_devdesc_free(ctx, dd); return code; };
  if ((rc = mg_sample_enqueue_dev(ctx, view, ra, cnt + 3, n0, st, mrSize, patches))) return fail(rc);
  if ((rc = mg_net_forward_enqueue(ctx, MODSGPU_HARDNET, patches, n0, nout, cnt + 3))) return fail(rc);
  // ---- one read-back: counters, region rows, descriptors (sized by the upper bound n0; the live rows are the first n3)
  const size_t row_bytes = (size_t)n0 * sizeof(DevRegion), desc_bytes = desc ? (size_t)n0 * 128 * 4 : 0;
  if (cudaSuccess != ctx->h_out.ensure(64 + row_bytes + desc_bytes)) { cudaGetLastError(); ctx->err = "pinned read-back buffer"; return fail(MODSGPU_ECUDA); }
  uint8_t* ho = ctx->h_out.as<uint8_t>();
  bool ok = cudaMemcpyAsync(ho, cnt, 16, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaMemcpyAsync(ho + 64, ra, row_bytes, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
  if (ok && desc) ok = cudaMemcpyAsync(ho + 64 + row_bytes, nout, desc_bytes, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
  if (!ok) { ctx->err = std::string("describe_view read-back: ") + cudaGetErrorString(cudaGetLastError()); return fail(MODSGPU_ECUDA); }
  if (mg_end(ctx)) return fail(MODSGPU_ECUDA);                        // host round trip 2 of 2
  const int* fc = reinterpret_cast<const int*>(ho);
  const int n3 = fc[3];
  if (counts) { counts[1] = fc[1]; counts[2] = n3; }
  modsgpu_view_region* r = (modsgpu_view_region*)malloc(sizeof(modsgpu_view_region) * (size_t)std::max(n3, 1));
  float* d = desc ? (float*)malloc(sizeof(float) * 128 * (size_t)std::max(n3, 1)) : nullptr;
  if (!r || (desc && !d)) { free(r); free(d); ctx->err = "out of host memory"; return fail(MODSGPU_ECUDA); }
  memcpy(r, ho + 64, (size_t)n3 * sizeof(DevRegion));
  if (desc) { memcpy(d, ho + 64 + row_bytes, (size_t)n3 * 512); *desc = d; }
  if (dd) { dd->n = n3; *devdesc = dd; }
  *regions = r; *n = n3;
  return 0;
}
