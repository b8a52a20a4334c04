// sampler.cu -- affine patch sampler on the device (seam S5).
//
// Restates ExtractPatchesColumn (synth-detection.cpp:38-132) + interpolate (helpers.cpp:551-626)
// + gaussianBlurInplace (helpers.cpp:726-731) + the float->u8 conversion of cv::imencode
// (imagerepresentation.cpp:45) with the exact arithmetic order of the CPU oracle
// (oracle/mods_oracle.cpp).  Compiled with --fmad=false; fused ops only as explicit fmaf().
//
// One CTA per region:
//   1. resample an R x R window (R = 2*ceil(s*mrSize)+2) with the region's affine frame.
//      The reference accumulates sample coordinates incrementally in float (WX += a11), so each
//      work item replays the additions from the row start -- the rounding sequence is the contract.
//   2. separable Gaussian (sigma = 1.5*R0/patchSize, cv::GaussianBlur order), evaluated only at the
//      rows/columns the final resampling touches when R is large,
//   3. resample to patchSize x patchSize, round-half-even to u8.
// Small windows (R <= 66) live entirely in shared memory; large ones use an HBM scratch slab.
#include "common.cuh"
#include <cmath>
#include <map>
#include <algorithm>

namespace {

constexpr int SMALL_R = 66;     // R0 <= 64
constexpr int MAX_R = 2048;
constexpr int MAX_PS = 64;
constexpr int CHUNK = 16;

struct PatchMeta {
  float x, y, a11, a12, a21, a22;   // region frame as the reference casts it to float
  float scale;                      // imageToPatchScale
  int R;                            // resampling window (0: direct mode, scale <= 0.4)
  int ks, tap_off;                  // Gaussian taps
  int out_index;                    // patch slot in the output
  long long scratch_off;            // floats, large windows only
};

__device__ __forceinline__ float bilinear(const float* im, int pitch, int x, int y, float WX, float WY) {
  const float wx = WX - (float)x;
  const float* Row0 = im + (size_t)y * pitch;
  const float* Row1 = Row0 + pitch;
  const float I1 = wx * (Row0[x + 1] - Row0[x]) + Row0[x];
  return (WY - (float)y) * (wx * (Row1[x + 1] - Row1[x]) + Row1[x] - I1) + I1;
}

// helpers.cpp:551-626 for one sample (uniform form of the fast and the border-checking path)
__device__ __forceinline__ float sample_image(const float* im, int w, int h, float WX, float WY) {
  const int x = (int)floorf(WX), y = (int)floorf(WY);
  if (WX >= 0 && WY >= 0 && x < w - 1 && y < h - 1) return bilinear(im, w, x, y, WX, WY);
  return 0.f;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cv::GaussianBlur row pass at column x of a row of length R (replicate border); see detect.cu
__device__ __forceinline__ float row_pass_at(const float* row, int R, int x, const float* k, int ks) {
  const int r = ks >> 1;
  auto PX = [&](int xx) { return row[clampi(xx, 0, R - 1)]; };
  float s;
  if (ks == 5) {
    float p1 = PX(x + 1) + PX(x - 1), p2 = PX(x + 2) + PX(x - 2), x0 = PX(x);
    if (x < (R & ~1)) {
      s = p1 * k[3];
      s = fmaf(x0, k[2], s);
      s = fmaf(p2, k[4], s);
    } else {
      s = x0 * k[2] + p1 * k[3];
      s = s + p2 * k[4];
    }
  } else if (ks < 5) {
    s = PX(x) * k[r];
    for (int t = 1; t <= r; t++) s = s + (PX(x + t) + PX(x - t)) * k[r + t];
  } else if (x < (R & ~3)) {
    s = 0.f;
    for (int t = 0; t < ks; t++) s = fmaf(PX(x + t - r), k[t], s);
  } else {
    const int nf = (ks - 1) % 4;
    s = PX(x - r) * k[0];
    for (int t = 1; t < ks; t++) {
      if (t >= ks - nf) s = fmaf(PX(x + t - r), k[t], s);
      else s = s + PX(x + t - r) * k[t];
    }
  }
  return s;
}

template <bool LARGE>
__global__ void __launch_bounds__(LARGE ? 256 : 128)
k_sample(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas,
         const float* __restrict__ taps_all, float* __restrict__ scratch, uint8_t* __restrict__ out, int ps) {
  extern __shared__ float sm[];
  const PatchMeta m = metas[blockIdx.x];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int R = m.R;
  uint8_t* dst = out + (size_t)m.out_index * ps * ps;

  if (R == 0) {
    // scale <= 0.4: one direct resampling with A*scale (synth-detection.cpp:117-127)
    const float a11 = m.a11 * m.scale, a12 = m.a12 * m.scale, a21 = m.a21 * m.scale, a22 = m.a22 * m.scale;
    const int half = ps / 2;
    for (int j = tid; j < ps; j += nth) {
      float rx = m.x - (float)half * a12, ry = m.y - (float)half * a22;
      for (int t = 0; t < j; t++) { rx += a12; ry += a22; }
      float WX = rx - (float)half * a11, WY = ry - (float)half * a21;
      for (int i = 0; i < ps; i++) {
        float v = sample_image(img, w, h, WX, WY);
        int q = __float2int_rn(v);
        dst[j * ps + i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
        WX += a11; WY += a21;
      }
    }
    return;
  }

  // ---- shared / scratch carve-up
  float* P = sm;                    // ps sample positions of the final resampling
  float* RX = P + MAX_PS;           // per-row start coordinates of the first resampling
  float* RY = RX + (LARGE ? MAX_R : SMALL_R);
  float* kk = RY + (LARGE ? MAX_R : SMALL_R);   // taps (<= 6*1.5*MAX_R/32+3)
  float* S; float* T; float* B;
  int nc;                           // number of columns (= rows) the blur is evaluated at
  if (LARGE) {
    B = kk + 640;
    S = scratch + m.scratch_off;
    T = S + (size_t)R * R;
    nc = 2 * ps;
  } else {
    S = kk + 32;
    T = S + SMALL_R * SMALL_R;
    B = S;
    nc = R;
  }
  const int ks = m.ks;
  for (int i = tid; i < ks; i += nth) kk[i] = taps_all[m.tap_off + i];

  // ---- 1. first resampling: S = interpolate(img; centre (x,y), A, R x R)
  const int half = R / 2;
  for (int j = tid; j < R; j += nth) {
    float rx = m.x - (float)half * m.a12, ry = m.y - (float)half * m.a22;
    for (int t = 0; t < j; t++) { rx += m.a12; ry += m.a22; }
    RX[j] = rx; RY[j] = ry;
  }
  if (tid == 0) {
    // positions of the second resampling: centre R/2 (integer division), diag(scale)
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; p += m.scale; }
  }
  __syncthreads();
  const int nchunk = (R + CHUNK - 1) / CHUNK;
  for (int it = tid; it < R * nchunk; it += nth) {
    const int j = it / nchunk, q = it - j * nchunk;
    float WX = RX[j] - (float)half * m.a11, WY = RY[j] - (float)half * m.a21;
    const int i0 = q * CHUNK;
    for (int t = 0; t < i0; t++) { WX += m.a11; WY += m.a21; }
    const int i1 = min(R, i0 + CHUNK);
    for (int i = i0; i < i1; i++) {
      S[(size_t)j * R + i] = sample_image(img, w, h, WX, WY);
      WX += m.a11; WY += m.a21;
    }
  }
  __syncthreads();

  // ---- 2. Gaussian blur (row pass -> T, column pass -> B) at the needed columns / rows
  // column / row list: identity (small) or {floor(P[i]), floor(P[i])+1} (large)
  auto pos_of = [&](int ci) -> int {
    if (!LARGE) return ci;
    int x = (int)floorf(P[ci >> 1]) + (ci & 1);
    return clampi(x, 0, R - 1);
  };
  for (int it = tid; it < R * nc; it += nth) {
    const int y = it / nc, ci = it - y * nc;
    T[(size_t)y * nc + ci] = row_pass_at(S + (size_t)y * R, R, pos_of(ci), kk, ks);
  }
  __syncthreads();
  {
    const int r = ks >> 1;
    const int wc = R & ~7;
    for (int it = tid; it < nc * nc; it += nth) {
      const int ri = it / nc, ci = it - ri * nc;
      const int y = pos_of(ri), x = pos_of(ci);
      auto TY = [&](int yy) { return T[(size_t)clampi(yy, 0, R - 1) * nc + ci]; };
      float s = TY(y) * kk[r];
      if (x < wc) {
        for (int t = 1; t <= r; t++) s = fmaf(TY(y - t) + TY(y + t), kk[r + t], s);
      } else {
        for (int t = 1; t <= r; t++) s = s + (TY(y - t) + TY(y + t)) * kk[r + t];
      }
      // small: B aliases S, which the row pass no longer needs -- but other threads still read T only
      B[(size_t)ri * nc + ci] = s;
    }
  }
  __syncthreads();

  // ---- 3. second resampling to ps x ps + u8 quantisation
  for (int it = tid; it < ps * ps; it += nth) {
    const int j = it / ps, i = it - j * ps;
    const float WX = P[i], WY = P[j];
    const int x = (int)floorf(WX), y = (int)floorf(WY);
    float v = 0.f;
    if (WX >= 0 && WY >= 0 && x < R - 1 && y < R - 1) {
      float v00, v01, v10, v11;
      if (LARGE) {
        const float* r0 = B + (size_t)(2 * j) * nc + 2 * i;
        const float* r1 = r0 + nc;
        v00 = r0[0]; v01 = r0[1]; v10 = r1[0]; v11 = r1[1];
      } else {
        const float* r0 = B + (size_t)y * nc + x;
        const float* r1 = r0 + nc;
        v00 = r0[0]; v01 = r0[1]; v10 = r1[0]; v11 = r1[1];
      }
      const float wx = WX - (float)x;
      const float I1 = wx * (v01 - v00) + v00;
      v = (WY - (float)y) * (wx * (v11 - v10) + v10 - I1) + I1;
    }
    int q = __float2int_rn(v);
    dst[it] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

constexpr int SMEM_SMALL = (MAX_PS + 2 * SMALL_R + 32 + 2 * SMALL_R * SMALL_R) * 4;
constexpr int SMEM_LARGE = (MAX_PS + 2 * MAX_R + 640 + (2 * MAX_PS) * (2 * MAX_PS)) * 4;

}  // namespace

// Enqueue the sampler for n regions (host array) on ctx->stream; u8 patches land in d_out
// (n * ps * ps bytes, region order).  No host synchronisation.
int mg_sample_enqueue(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                      double mrSize, int ps, uint8_t* d_out) {
  if (ps < 2 || ps > MAX_PS) MG_FAIL(ctx, MODSGPU_EINVAL, "patchSize out of range");
  if (n <= 0) return 0;
  std::vector<PatchMeta> small, large;
  small.reserve(n);
  std::map<int, std::pair<int, int>> tap_index;  // R0 -> (offset, ks)
  std::vector<float> taps_all;
  long long scratch = 0;
  for (int i = 0; i < n; i++) {
    const modsgpu_region& k = regs[i];
    PatchMeta m;
    m.x = (float)k.x; m.y = (float)k.y;
    m.a11 = (float)k.a11; m.a12 = (float)k.a12; m.a21 = (float)k.a21; m.a22 = (float)k.a22;
    float mrScale = (float)std::ceil(k.s * mrSize);
    int R0 = ps % 2 != 0 ? 2 * int(mrScale) + 1 : 2 * int(mrScale);
    m.scale = float(R0) / float(ps);
    m.out_index = i;
    m.scratch_off = 0;
    m.ks = 0; m.tap_off = 0;
    if (m.scale > 0.4) {
      m.R = R0 + 2;
      if (m.R > MAX_R) MG_FAIL(ctx, MODSGPU_EINVAL, "region too large for the sampler (R > 2048)");
      auto it = tap_index.find(R0);
      if (it == tap_index.end()) {
        std::vector<float> t;
        int ks = mg_gaussian_taps(1.5f * m.scale, t);
        it = tap_index.emplace(R0, std::make_pair((int)taps_all.size(), ks)).first;
        taps_all.insert(taps_all.end(), t.begin(), t.end());
      }
      m.tap_off = it->second.first; m.ks = it->second.second;
      if (m.ks > 600) MG_FAIL(ctx, MODSGPU_EINVAL, "sampler blur too wide");
      if (m.R <= SMALL_R && m.ks <= 31) small.push_back(m);
      else {
        m.scratch_off = scratch;
        scratch += (long long)m.R * m.R + (long long)m.R * 2 * ps;
        large.push_back(m);
      }
    } else {
      m.R = 0;
      small.push_back(m);
    }
  }
  // biggest windows first so the tail of the launch is short
  std::sort(large.begin(), large.end(), [](const PatchMeta& a, const PatchMeta& b) { return a.R > b.R; });
  const size_t nm = small.size() + large.size();
  MG_CUDA(ctx, ctx->h_stage2.ensure(nm * sizeof(PatchMeta) + taps_all.size() * 4 + 64));
  PatchMeta* hm = ctx->h_stage2.as<PatchMeta>();
  if (!small.empty()) memcpy(hm, small.data(), small.size() * sizeof(PatchMeta));
  if (!large.empty()) memcpy(hm + small.size(), large.data(), large.size() * sizeof(PatchMeta));
  float* ht = reinterpret_cast<float*>(hm + nm);
  if (!taps_all.empty()) memcpy(ht, taps_all.data(), taps_all.size() * 4);
  MG_CUDA(ctx, ctx->smp_meta.ensure(nm * sizeof(PatchMeta)));
  MG_CUDA(ctx, ctx->smp_taps.ensure(taps_all.size() * 4 + 16));
  MG_CUDA(ctx, ctx->smp_scratch.ensure((size_t)scratch * 4 + 16));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_meta.p, hm, nm * sizeof(PatchMeta), cudaMemcpyHostToDevice, ctx->stream));
  if (!taps_all.empty())
    MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_taps.p, ht, taps_all.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LARGE));
    attr_set = true;
  }
  const PatchMeta* dm = ctx->smp_meta.as<PatchMeta>();
  double bytes_small = 0, bytes_large = 0;   // algorithmic: R*R*4 read + ps*ps written per region
  for (const PatchMeta& m : small) bytes_small += (double)m.R * m.R * 4.0 + (double)ps * ps;
  for (const PatchMeta& m : large) bytes_large += (double)m.R * m.R * 4.0 + (double)ps * ps;
  if (!small.empty()) {
    MG_PROF(ctx, "k_sample<small>", 0, bytes_small);
    k_sample<false><<<(unsigned)small.size(), 128, SMEM_SMALL, ctx->stream>>>(
        img->d, img->w, img->h, dm, ctx->smp_taps.as<float>(), nullptr, d_out, ps);
    MG_LAUNCHED(ctx);
  }
  if (!large.empty()) {
    MG_PROF(ctx, "k_sample<large>", 0, bytes_large);
    k_sample<true><<<(unsigned)large.size(), 256, SMEM_LARGE, ctx->stream>>>(
        img->d, img->w, img->h, dm + small.size(), ctx->smp_taps.as<float>(), ctx->smp_scratch.as<float>(), d_out, ps);
    MG_LAUNCHED(ctx);
  }
  return 0;
}

extern "C" int modsgpu_extract_patches(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                       double mrSize, int patchSize, uint8_t* out) {
  if (!ctx || !img || (n > 0 && (!regs || !out)) || n < 0) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  size_t bytes = (size_t)n * patchSize * patchSize;
  MG_CUDA(ctx, ctx->smp_out.ensure(bytes + 16));
  int rc = mg_sample_enqueue(ctx, img, regs, n, mrSize, patchSize, ctx->smp_out.as<uint8_t>());
  if (rc) return rc;
  if (n > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->smp_out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}
