// classic.cu -- the "classic" per-region stages of config_affori_classic.ini on the device:
//   k_dom_ori  dominant gradient orientation (DetectOrientation synth-detection.cpp:1039-1149 +
//              EstimateDominantAnglesFunctor :836-929): 32x32 direct affine patch, gradient magnitude / LUT
//              orientation, 36-bin histogram weighted by a circular Gaussian, 6 circular box smoothings, peaks
//   k_sift     (Root)SIFT (SIFTDescriptor matching/siftdesc.cpp + DescribeRegions synth-detection.hpp:170-263 +
//              photometricallyNormalize helpers.cpp:666-716) on the 41x41 float patch of the 3-step sampler
// Both accumulate floats / doubles in the reference's raster order: one warp per region; the per-pixel quantities
// are computed by all lanes, the order-dependent sums by one lane (histogram, photometric sums) or by the 8 lanes
// that own the 8 distinct bins a pixel updates (SIFT).  --fmad=false; sqrtf / division are IEEE on the device.
// The atan LUT (helpers.cpp:30-72) is rebuilt as round(atan(i/255), 1e-10) plus the reference's three typo entries.
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace modsb200 {
bool interpolateCheckBorders(int orig_img_w, int orig_img_h, float ofsx, float ofsy, float a11, float a12, float a21,
                             float a22, int res_w, int res_h);
}
int mg_sample_enqueue(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                      double mrSize, int ps, uint8_t* d_out, float* d_outf);

namespace {

constexpr int ORI_BINS = 36;
constexpr int ORI_MAX_PS = 32;
constexpr int SIFT_MAX_PS = 41;

struct OriMeta { float x, y, a11, a12, a21, a22; int out; };

__device__ __forceinline__ float sample_image_c(const float* im, int w, int h, float WX, float WY) {
  const int x = (int)floorf(WX), y = (int)floorf(WY);
  if (WX >= 0 && WY >= 0 && x < w - 1 && y < h - 1) {
    const float wx = WX - (float)x;
    const float* Row0 = im + (size_t)y * w;
    const float* Row1 = Row0 + w;
    const float I1 = wx * (Row0[x + 1] - Row0[x]) + Row0[x];
    return (WY - (float)y) * (wx * (Row1[x + 1] - Row1[x]) + Row1[x] - I1) + I1;
  }
  return 0.f;
}

// helpers.cpp:160-207 atan2LUTff (LUT entries are doubles; the sums are evaluated in double and returned as float)
__device__ __forceinline__ float atan2LUTff(const double* __restrict__ LUT, float y, float x) {
  const float PI2f = 1.57079632679489661923f, PIf = 3.14159265358979323846f;
  if (x > 0.f) {
    if (y > 0.f) {
      if (x > y) return (float)LUT[(int)(255.f * y / x)];
      return (float)(PI2f - LUT[(int)(255 * x / y)]);
    } else {
      const float absy = fabsf(y);
      if (x > absy) return (float)(-LUT[(int)(255.f * absy / x)]);
      return (float)(-PI2f + LUT[(int)(255.f * x / absy)]);
    }
  } else if (y > 0.f) {
    const float absx = fabsf(x);
    if (absx > y) return (float)(PIf - LUT[(int)(255.f * y / absx)]);
    return (float)(PI2f + LUT[(int)(255.f * absx / y)]);
  } else {
    const float absx = fabsf(x), absy = fabsf(y);
    if (absx > absy) return (float)(-PIf + LUT[(int)(255.f * absy / absx)]);
    if (x == 0.f) return 0.f;
    return (float)(-PI2f - LUT[(int)(255.f * absx / absy)]);
  }
}

// ---- dominant orientation: one warp per region ------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_dom_ori(const float* __restrict__ img, int w, int h, const OriMeta* __restrict__ metas, int n, int ps,
          const float* __restrict__ orimask, const double* __restrict__ lut_g, int maxAngles, double th,
          int* __restrict__ n_ang, float* __restrict__ angles) {
  __shared__ double LUT[256];
  __shared__ float patch_s[4][ORI_MAX_PS * ORI_MAX_PS];
  __shared__ float contrib_s[4][ORI_MAX_PS * ORI_MAX_PS];
  __shared__ signed char bin_s[4][ORI_MAX_PS * ORI_MAX_PS];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) LUT[i] = lut_g[i];
  __syncthreads();
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int reg = blockIdx.x * 4 + wl;
  if (reg >= n) return;
  const OriMeta m = metas[reg];
  float* patch = patch_s[wl];
  float* contrib = contrib_s[wl];
  signed char* bin = bin_s[wl];
  // interpolate(img, x, y, A*curr_sc, patch) (helpers.cpp:551-626): lane = patch row, coordinates accumulated in order
  const int half = ps / 2;
  for (int j = lane; j < ps; j += 32) {
    float rx = m.x - (float)half * m.a12, ry = m.y - (float)half * m.a22;
    for (int t = 0; t < j; t++) { rx += m.a12; ry += m.a22; }
    float WX = rx - (float)half * m.a11, WY = ry - (float)half * m.a21;
    for (int i = 0; i < ps; i++) {
      patch[j * ps + i] = sample_image_c(img, w, h, WX, WY);
      WX += m.a11; WY += m.a21;
    }
  }
  __syncwarp();
  // per-pixel magnitude / orientation bin (helpers.cpp:840-863 + synth-detection.cpp:868-880), all lanes
  const int maskPixels = ps * (ps - 2);
  for (int q = lane; q < maskPixels; q += 32) {
    const int r = 1 + q / ps, c = q - (r - 1) * ps;
    signed char b = -1;
    float v = 0.f;
    if (c >= 1 && c < ps - 1) {
      const float xgrad = patch[r * ps + c + 1] - patch[r * ps + c - 1];
      const float ygrad = patch[(r + 1) * ps + c] - patch[(r - 1) * ps + c];
      const float g = sqrtf(xgrad * xgrad + ygrad * ygrad);
      const float mk = orimask[ps + q];
      if (mk > 0 && g > 1.0f) {
        const float o = atan2LUTff(LUT, ygrad, xgrad);
        b = (signed char)(int)((float)ORI_BINS * (o / 3.14159265358979323846f + 1.0f) / 2.0f);
        v = g * mk;
      }
    }
    bin[q] = b; contrib[q] = v;
  }
  __syncwarp();
  if (lane != 0) return;
  float hist[ORI_BINS + 1];
  for (int b = 0; b <= ORI_BINS; b++) hist[b] = 0.0f;
  for (int q = 0; q < maskPixels; q++) {        // raster order: the float sums are order dependent
    const int b = bin[q];
    if (b >= 0) hist[b] += contrib[q];
  }
  for (int it = 0; it < 6; it++) {              // smoothCircularBuffer, synth-detection.cpp:811-822
    float first = hist[0], prev = hist[ORI_BINS - 1];
    for (int b = 0; b < ORI_BINS - 1; b++) { float cur = hist[b]; hist[b] = prev + cur + hist[b + 1]; prev = cur; }
    hist[ORI_BINS - 1] = prev + hist[ORI_BINS - 1] + first;
  }
  float thresh = 0.0f;
  for (int b = 0; b < ORI_BINS; b++) if (hist[b] > thresh) thresh = hist[b];
  thresh = (float)((double)thresh * th);
  int cnt = 0, seen = 0;
  bool stop = false;
  auto addPeak = [&](int a, int b, int c) {
    if (stop) return;
    if (hist[b] >= thresh && hist[b] > hist[a] && hist[b] > hist[c]) {
      // peaks are taken in bin order; every accepted peak satisfies value >= thresh
      if (seen < maxAngles) {
        const float pp = (hist[a] - hist[c]) / (hist[a] - 2.0f * hist[b] + hist[c]) / 2.0f;
        angles[(size_t)m.out * maxAngles + cnt++] = 2.0f * 3.14159265358979323846f * ((float)b + 0.5f + pp) / (float)ORI_BINS - 3.14159265358979323846f;
      } else stop = true;
      seen++;
    }
  };
  addPeak(ORI_BINS - 1, 0, 1);
  for (int b = 1; b < ORI_BINS - 1; b++) addPeak(b - 1, b, b + 1);
  addPeak(ORI_BINS - 2, ORI_BINS - 1, 0);
  n_ang[m.out] = cnt;
}

// ---- (Root)SIFT: one warp per region ------------------------------------------------------------------------------
struct SiftTab {
  int bin0[SIFT_MAX_PS], bin1[SIFT_MAX_PS];
  double w0[SIFT_MAX_PS], w1[SIFT_MAX_PS];
};

__global__ void __launch_bounds__(128)
k_sift(const float* __restrict__ patches, int n, int ps, const float* __restrict__ mask, const double* __restrict__ lut_g,
       SiftTab tab, int photoNorm, int rootSift, double maxBinValue, float* __restrict__ out) {
  __shared__ double LUT[256];
  __shared__ float patch_s[4][SIFT_MAX_PS * SIFT_MAX_PS];
  __shared__ double vec_s[4][128];
  __shared__ float stat_s[4][2];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) LUT[i] = lut_g[i];
  __syncthreads();
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int reg = blockIdx.x * 4 + wl;
  if (reg >= n) return;
  float* patch = patch_s[wl];
  double* vec = vec_s[wl];
  const int npx = ps * ps;
  const float* src = patches + (size_t)reg * npx;
  for (int q = lane; q < npx; q += 32) patch[q] = src[q];
  for (int q = lane; q < 128; q += 32) vec[q] = 0.0;
  __syncwarp();
  if (photoNorm) {   // helpers.cpp:666-716: float sums in raster order over the circular mask
    if (lane == 0) {
      float sum = 0.f, gsum = 0.f;
      for (int q = 0; q < npx; q++) if (mask[q] > 0) { sum += patch[q]; gsum += 1.f; }
      sum = sum / gsum;
      float var = 0.f;
      for (int q = 0; q < npx; q++) if (mask[q] > 0) var += (sum - patch[q]) * (sum - patch[q]);
      var = sqrtf(var / gsum);
      stat_s[wl][0] = sum; stat_s[wl][1] = var;
    }
    __syncwarp();
    const float sum = stat_s[wl][0], var = stat_s[wl][1];
    if (!((double)var < 0.0001)) {
      const float fac = 50.0f / var;
      for (int q = lane; q < npx; q += 32) {
        float v = 128 + fac * (patch[q] - sum);
        if (v > 255) v = 255;
        if (v < 0) v = 0;
        patch[q] = v;
      }
    }
    __syncwarp();
  }
  // gradients (siftdesc.cpp:279-302) + trilinear histogram (samplePatch :73-130).  32 pixels at a time: every lane
  // prepares one pixel, then the pixels are applied in raster order; the (up to) 8 bins one pixel updates are
  // distinct, lanes 0..7 own one (row bin, column bin, orientation bin) combination each.
  const double M_PI_DOUBLED = 6.28318530718;
  for (int base = 0; base < npx; base += 32) {
    const int q = base + lane;
    float val = 0.f, wo0 = 0.f, wo1 = 0.f;
    int bo0 = 0, r = 0, c = 0;
    if (q < npx) {
      r = q / ps; c = q - r * ps;
      float xgrad, ygrad;
      if (c == 0) xgrad = patch[q + 1] - patch[q];
      else if (c == ps - 1) xgrad = patch[q] - patch[q - 1];
      else xgrad = patch[q + 1] - patch[q - 1];
      if (r == 0) ygrad = patch[q + ps] - patch[q];
      else if (r == ps - 1) ygrad = patch[q] - patch[q - ps];
      else ygrad = patch[q + ps] - patch[q - ps];
      const float g = sqrtf(xgrad * xgrad + ygrad * ygrad);
      const float oriv = atan2LUTff(LUT, ygrad, xgrad);
      val = (float)(0.0 * 1.0 + (1.0 - 0.0) * mask[q] * g);     // magnLess = false
      const float o = (float)(8.0f * ((double)oriv + M_PI_DOUBLED) / M_PI_DOUBLED);
      bo0 = (int)o;
      wo1 = o - (float)bo0;
      bo0 %= 8;
      wo0 = 1.0f - wo1;
    }
    const int cnt = min(32, npx - base);
    for (int k = 0; k < cnt; k++) {
      const float v_k = __shfl_sync(0xffffffffu, val, k), wo0_k = __shfl_sync(0xffffffffu, wo0, k), wo1_k = __shfl_sync(0xffffffffu, wo1, k);
      const int bo0_k = __shfl_sync(0xffffffffu, bo0, k), r_k = __shfl_sync(0xffffffffu, r, k), c_k = __shfl_sync(0xffffffffu, c, k);
      if (lane < 8) {
        const int ri = lane >> 2, ci = (lane >> 1) & 1, oi = lane & 1;
        const int br = 4 * (ri ? tab.bin1[r_k] : tab.bin0[r_k]);
        const float wr = (float)(ri ? tab.w1[r_k] : tab.w0[r_k]);
        const int bc = ci ? tab.bin1[c_k] : tab.bin0[c_k];
        const float wc = (float)((ci ? tab.w1[c_k] : tab.w0[c_k]) * v_k);
        const int bo = oi ? (bo0_k + 1) % 8 : bo0_k;
        const float wo = oi ? wo1_k : wo0_k;
        const float vv = wr * wc;
        if (vv > 0) vec[br + bc + bo] += (double)(vv * wo);
      }
      __syncwarp();
    }
  }
  // SIFTnorm / RootSIFTnorm on doubles (siftdesc.cpp:132-159, :196-249): sequential sums by lane 0
  if (lane == 0) {
    for (int pass = 0; pass < 2; pass++) {
      double len = 0.0;
      for (int i = 0; i < 128; i += 4) {
        const double sq0 = vec[i] * vec[i], sq1 = vec[i + 1] * vec[i + 1], sq2 = vec[i + 2] * vec[i + 2], sq3 = vec[i + 3] * vec[i + 3];
        len += sq0 + sq1 + sq2 + sq3;
      }
      len = sqrt(len);
      const double fac = 1.0 / len;
      for (int i = 0; i < 128; i++) vec[i] *= fac;
      if (pass == 1) break;
      bool changed = false;
      for (int i = 0; i < 128; i++) if (vec[i] > maxBinValue) { vec[i] = maxBinValue; changed = true; }
      if (!changed) break;
    }
    if (rootSift) {
      double sum = 0.;
      for (int i = 0; i < 128; i++) sum += fabs(vec[i]);
      for (int i = 0; i < 128; i++) vec[i] = sqrt(vec[i] / sum);
    }
  }
  __syncwarp();
  for (int i = lane; i < 128; i += 32) {
    int b = (int)(512.0 * vec[i] + 0.5);     // 0.5 "for appropriate rounding" (siftdesc.cpp:218, :246)
    b = max(0, min(b, 255));
    out[(size_t)reg * 128 + i] = (float)b;
  }
}

void build_lut(double* lut) {
  char buf[64];
  for (int i = 0; i < 256; i++) {
    snprintf(buf, sizeof(buf), "%.10f", std::atan(i / 255.0));
    lut[i] = strtod(buf, nullptr);
  }
  // typos of the reference's literal table (helpers.cpp:30-72), reproduced
  lut[32] = strtod("0.1248376255", nullptr);
  lut[83] = strtod("0.3146752558", nullptr);
  lut[100] = strtod("0.3737268255", nullptr);
}
void circular_mask(std::vector<float>& mask, int size, float sigma) {   // helpers.cpp:442-459
  mask.resize((size_t)size * size);
  const int halfSize = size >> 1;
  const float r2 = float(halfSize * halfSize);
  const float sigma2 = sigma == 0 ? 0.9f * r2 : 2 * sigma * sigma;
  float* mp = mask.data();
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++) {
      const float disq = float((i - halfSize) * (i - halfSize) + (j - halfSize) * (j - halfSize));
      *mp++ = (disq < r2) ? std::exp(-disq / sigma2) : 0;
    }
}

}  // namespace

extern "C" int modsgpu_dominant_orientation(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                            double mrSize, int patchSize, int maxAngles, double th, int* n_ang, float* angles) {
  if (!ctx || !img || n < 0 || (n > 0 && (!regs || !n_ang || !angles)) || maxAngles < 0) return MODSGPU_EINVAL;
  if (patchSize < 3 || patchSize > ORI_MAX_PS) MG_FAIL(ctx, MODSGPU_EINVAL, "dominant orientation: patchSize must be in [3, 32]");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  const double mrScale = (double)mrSize;
  const int patchImageSize = 2 * int(mrScale) + 1;
  const double imageToPatchScale = double(patchImageSize) / (double)patchSize;
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
  std::vector<OriMeta> metas;
  metas.reserve(n);
  for (int i = 0; i < n; i++) {
    const modsgpu_region& k = regs[i];
    n_ang[i] = 0;
    if (modsb200::interpolateCheckBorders(img->w, img->h, (float)k.x, (float)k.y, (float)k.a11, (float)k.a12, (float)k.a21,
                                          (float)k.a22, (int)(k_sigma * k.s), (int)(k_sigma * k.s))) { n_ang[i] = -1; continue; }
    if (maxAngles <= 0) continue;
    const float curr_sc = imageToPatchScale * k.s;
    OriMeta m;
    m.x = (float)k.x; m.y = (float)k.y;
    m.a11 = (float)k.a11 * curr_sc; m.a12 = (float)k.a12 * curr_sc; m.a21 = (float)k.a21 * curr_sc; m.a22 = (float)k.a22 * curr_sc;
    m.out = i;
    metas.push_back(m);
  }
  const int nk = (int)metas.size();
  if (nk == 0 || maxAngles == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  std::vector<float> mask;
  circular_mask(mask, patchSize, patchSize / 3.0f);
  double lut[256];
  build_lut(lut);
  const size_t mb = ((size_t)nk * sizeof(OriMeta) + 15) & ~(size_t)15, kb = mask.size() * 4, lb = 256 * 8;
  const size_t ab = (size_t)n * maxAngles * 4, cb = ((size_t)n * 4 + 15) & ~(size_t)15;
  MG_CUDA(ctx, ctx->io_a.ensure(lb + mb + kb + 64));
  MG_CUDA(ctx, ctx->io_b.ensure(cb + ab + 64));
  uint8_t* da = ctx->io_a.as<uint8_t>();
  double* d_lut = reinterpret_cast<double*>(da);
  OriMeta* d_m = reinterpret_cast<OriMeta*>(da + lb);
  float* d_mask = reinterpret_cast<float*>(da + lb + mb);
  int* d_cnt = ctx->io_b.as<int>();
  float* d_ang = reinterpret_cast<float*>(ctx->io_b.as<uint8_t>() + cb);
  MG_CUDA(ctx, cudaMemcpyAsync(d_lut, lut, lb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(d_m, metas.data(), (size_t)nk * sizeof(OriMeta), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(d_mask, mask.data(), kb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, (size_t)n * 4, ctx->stream));
  MG_CUDA(ctx, cudaMemsetAsync(d_ang, 0, ab, ctx->stream));
  MG_PROF(ctx, "k_dom_ori", 2, (double)nk);
  k_dom_ori<<<ceil_div(nk, 4), 128, 0, ctx->stream>>>(img->d, img->w, img->h, d_m, nk, patchSize, d_mask, d_lut, maxAngles, th, d_cnt, d_ang);
  MG_LAUNCHED(ctx);
  std::vector<int> hc(n);
  MG_CUDA(ctx, cudaMemcpyAsync(hc.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(angles, d_ang, ab, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  for (int i = 0; i < n; i++) if (n_ang[i] >= 0) n_ang[i] = hc[i];
  return 0;
}

extern "C" int modsgpu_describe_sift(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                     double mrSize, int patchSize, int photoNorm, int rootSift, float* out) {
  if (!ctx || !img || n < 0 || (n > 0 && (!regs || !out))) return MODSGPU_EINVAL;
  if (patchSize < 3 || patchSize > SIFT_MAX_PS) MG_FAIL(ctx, MODSGPU_EINVAL, "SIFT: patchSize must be in [3, 41]");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (n == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const int ps = patchSize, npx = ps * ps;
  // siftdesc.cpp:22-71 precomputeBinsAndWeights (spatialBins 4, orientationBins 8)
  SiftTab tab;
  memset(&tab, 0, sizeof(tab));
  {
    const int spatialBins = 4, orientationBins = 8, halfSize = ps >> 1;
    const float step = float(spatialBins + 1) / (2 * halfSize);
    for (int i = 0; i < ps; i++) {
      float x = step * i;
      int xi = (int)(x);
      tab.bin0[i] = xi - 1; tab.bin1[i] = xi;
      tab.w1[i] = x - xi;
      tab.w0[i] = 1.0f - tab.w1[i];
      if (tab.bin0[i] < 0) { tab.bin0[i] = 0; tab.w0[i] = 0; }
      if (tab.bin0[i] >= spatialBins) { tab.bin0[i] = spatialBins - 1; tab.w0[i] = 0; }
      if (tab.bin1[i] < 0) { tab.bin1[i] = 0; tab.w1[i] = 0; }
      if (tab.bin1[i] >= spatialBins) { tab.bin1[i] = spatialBins - 1; tab.w1[i] = 0; }
      tab.bin0[i] *= orientationBins; tab.bin1[i] *= orientationBins;
    }
  }
  std::vector<float> mask;
  circular_mask(mask, ps, 0.f);
  double lut[256];
  build_lut(lut);
  const size_t pb = (size_t)n * npx * 4, kb = (size_t)npx * 4, lb = 256 * 8, ob = (size_t)n * 128 * 4;
  MG_CUDA(ctx, ctx->smp_regs.ensure(pb + 64));                   // float patches
  MG_CUDA(ctx, ctx->io_a.ensure(lb + kb + 64));
  MG_CUDA(ctx, ctx->io_b.ensure(ob + 64));
  float* d_patches = ctx->smp_regs.as<float>();
  double* d_lut = ctx->io_a.as<double>();
  float* d_mask = reinterpret_cast<float*>(ctx->io_a.as<uint8_t>() + lb);
  float* d_out = ctx->io_b.as<float>();
  int rc = mg_sample_enqueue(ctx, img, regs, n, mrSize, ps, nullptr, d_patches);
  if (rc) return rc;
  MG_CUDA(ctx, cudaMemcpyAsync(d_lut, lut, lb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(d_mask, mask.data(), kb, cudaMemcpyHostToDevice, ctx->stream));
  MG_PROF(ctx, "k_sift", 2, (double)n);
  k_sift<<<ceil_div(n, 4), 128, 0, ctx->stream>>>(d_patches, n, ps, d_mask, d_lut, tab, photoNorm, rootSift, 0.2, d_out);
  MG_LAUNCHED(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}
