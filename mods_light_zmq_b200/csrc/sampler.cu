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
// Small windows (R <= 66): one CTA per region, everything in shared memory.
// Large windows: the three phases are separate launches over flattened (region, row-block) work lists, so
// a 600-px window is spread over hundreds of CTAs instead of serialising one; S and the row-filtered
// columns live in an HBM scratch slab (L2 resident).
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


// ---- small windows: one CTA per region -----------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_sample_small(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas,
               const float* __restrict__ taps_all, uint8_t* __restrict__ out, int ps) {
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
  float* P = sm;                    // ps sample positions of the final resampling
  float* RX = P + MAX_PS;           // per-row start coordinates of the first resampling
  float* RY = RX + SMALL_R;
  float* kk = RY + SMALL_R;         // taps (<= 31)
  float* S = kk + 32;
  float* T = S + SMALL_R * SMALL_R;
  float* B = S;                     // the column pass overwrites S (only T is read by then)
  const int nc = R;
  const int ks = m.ks;
  for (int i = tid; i < ks; i += nth) kk[i] = taps_all[m.tap_off + i];

  // 1. first resampling: S = interpolate(img; centre (x,y), A, R x R)
  const int half = R / 2;
  for (int j = tid; j < R; j += nth) {
    float rx = m.x - (float)half * m.a12, ry = m.y - (float)half * m.a22;
    for (int t = 0; t < j; t++) { rx += m.a12; ry += m.a22; }
    RX[j] = rx; RY[j] = ry;
  }
  if (tid == 0) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;   // centre R/2 (integer division), diag(scale)
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
      S[j * R + i] = sample_image(img, w, h, WX, WY);
      WX += m.a11; WY += m.a21;
    }
  }
  __syncthreads();
  // 2. Gaussian blur: row pass -> T, column pass -> B
  for (int it = tid; it < R * nc; it += nth) {
    const int y = it / nc, x = it - y * nc;
    T[y * nc + x] = row_pass_at(S + y * R, R, x, kk, ks);
  }
  __syncthreads();
  {
    const int r = ks >> 1;
    const int wc = R & ~7;
    for (int it = tid; it < nc * nc; it += nth) {
      const int y = it / nc, x = it - y * nc;
      auto TY = [&](int yy) { return T[clampi(yy, 0, R - 1) * nc + x]; };
      float s = TY(y) * kk[r];
      if (x < wc) {
        for (int t = 1; t <= r; t++) s = fmaf(TY(y - t) + TY(y + t), kk[r + t], s);
      } else {
        for (int t = 1; t <= r; t++) s = s + (TY(y - t) + TY(y + t)) * kk[r + t];
      }
      B[y * nc + x] = s;
    }
  }
  __syncthreads();
  // 3. second resampling to ps x ps + u8 quantisation
  for (int it = tid; it < ps * ps; it += nth) {
    const int j = it / ps, i = it - j * ps;
    const float WX = P[i], WY = P[j];
    const int x = (int)floorf(WX), y = (int)floorf(WY);
    float v = 0.f;
    if (WX >= 0 && WY >= 0 && x < R - 1 && y < R - 1) {
      const float* r0 = B + y * nc + x;
      const float* r1 = r0 + nc;
      const float wx = WX - (float)x;
      const float I1 = wx * (r0[1] - r0[0]) + r0[0];
      v = (WY - (float)y) * (wx * (r1[1] - r1[0]) + r1[0] - I1) + I1;
    }
    int q = __float2int_rn(v);
    dst[it] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

// ---- large windows: flattened (region, row-block) work lists -------------------------------------------
// pre[] = exclusive prefix sums of the per-region block counts; binary search maps blockIdx -> region
__device__ __forceinline__ int find_region(const int* __restrict__ pre, int n, int b) {
  int lo = 0, hi = n;   // largest k with pre[k] <= b
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pre[mid] <= b) lo = mid; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int needed_pos(float p, int odd, int R) {
  return clampi((int)floorf(p) + odd, 0, R - 1);
}

constexpr int L1_ROWS = 4, L2_ROWS = 8, L3_OUT_ROWS = 4;

// phase 1: S[j][i] for L1_ROWS rows of one region
__global__ void __launch_bounds__(128)
k_large_resample(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas, int nreg,
                 const int* __restrict__ pre, float* __restrict__ scratch) {
  const int reg = find_region(pre, nreg, blockIdx.x);
  const PatchMeta m = metas[reg];
  const int R = m.R, half = R / 2;
  const int row0 = (blockIdx.x - pre[reg]) * L1_ROWS;
  float* S = scratch + m.scratch_off;
  const int nchunk = (R + CHUNK - 1) / CHUNK;
  const int nrows = min(L1_ROWS, R - row0);
  for (int it = threadIdx.x; it < nrows * nchunk; it += blockDim.x) {
    const int jr = it / nchunk, q = it - jr * nchunk, j = row0 + jr;
    float rx = m.x - (float)half * m.a12, ry = m.y - (float)half * m.a22;
    for (int t = 0; t < j; t++) { rx += m.a12; ry += m.a22; }
    float WX = rx - (float)half * m.a11, WY = ry - (float)half * m.a21;
    const int i0 = q * CHUNK;
    for (int t = 0; t < i0; t++) { WX += m.a11; WY += m.a21; }
    const int i1 = min(R, i0 + CHUNK);
    for (int i = i0; i < i1; i++) {
      S[(size_t)j * R + i] = sample_image(img, w, h, WX, WY);
      WX += m.a11; WY += m.a21;
    }
  }
}

// phase 2a: row pass at the 2*ps needed columns for L2_ROWS rows of one region: T[y][ci]
__global__ void __launch_bounds__(256)
k_large_rowpass(const PatchMeta* __restrict__ metas, int nreg, const int* __restrict__ pre,
                const float* __restrict__ taps_all, float* __restrict__ scratch, int ps) {
  extern __shared__ float sm[];
  const int reg = find_region(pre, nreg, blockIdx.x);
  const PatchMeta m = metas[reg];
  const int R = m.R, ks = m.ks, nc = 2 * ps;
  const int row0 = (blockIdx.x - pre[reg]) * L2_ROWS;
  const int nrows = min(L2_ROWS, R - row0);
  float* P = sm;              // ps
  float* kk = P + MAX_PS;     // ks <= 640
  float* rows = kk + 640;     // nrows x R
  const float* S = scratch + m.scratch_off;
  float* T = scratch + m.scratch_off + (size_t)R * R;
  for (int i = threadIdx.x; i < ks; i += blockDim.x) kk[i] = taps_all[m.tap_off + i];
  for (int i = threadIdx.x; i < nrows * R; i += blockDim.x) rows[i] = S[(size_t)row0 * R + i];
  if (threadIdx.x == 0) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; p += m.scale; }
  }
  __syncthreads();
  for (int it = threadIdx.x; it < nrows * nc; it += blockDim.x) {
    const int yr = it / nc, ci = it - yr * nc;
    const int x = needed_pos(P[ci >> 1], ci & 1, R);
    T[(size_t)(row0 + yr) * nc + ci] = row_pass_at(rows + yr * R, R, x, kk, ks);
  }
}

// phase 2b + 3: column pass at the needed rows of L3_OUT_ROWS output rows, then the final resampling
__global__ void __launch_bounds__(256)
k_large_colpass_final(const PatchMeta* __restrict__ metas, const float* __restrict__ taps_all,
                      const float* __restrict__ scratch, uint8_t* __restrict__ out, int ps) {
  extern __shared__ float sm[];
  const int nblk = (ps + L3_OUT_ROWS - 1) / L3_OUT_ROWS;
  const int reg = blockIdx.x / nblk, j0 = (blockIdx.x - reg * nblk) * L3_OUT_ROWS;
  const PatchMeta m = metas[reg];
  const int R = m.R, ks = m.ks, nc = 2 * ps, r = ks >> 1;
  const int nout = min(L3_OUT_ROWS, ps - j0);
  float* P = sm;
  float* kk = P + MAX_PS;
  float* B = kk + 640;        // (2*nout) x nc
  const float* T = scratch + m.scratch_off + (size_t)R * R;
  for (int i = threadIdx.x; i < ks; i += blockDim.x) kk[i] = taps_all[m.tap_off + i];
  if (threadIdx.x == 0) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; p += m.scale; }
  }
  __syncthreads();
  const int wc = R & ~7;
  for (int it = threadIdx.x; it < 2 * nout * nc; it += blockDim.x) {
    const int rl = it / nc, ci = it - rl * nc;
    const int ri = 2 * j0 + rl;
    const int y = needed_pos(P[ri >> 1], ri & 1, R), x = needed_pos(P[ci >> 1], ci & 1, R);
    auto TY = [&](int yy) { return T[(size_t)clampi(yy, 0, R - 1) * nc + ci]; };
    float s = TY(y) * kk[r];
    if (x < wc) {
      for (int t = 1; t <= r; t++) s = fmaf(TY(y - t) + TY(y + t), kk[r + t], s);
    } else {
      for (int t = 1; t <= r; t++) s = s + (TY(y - t) + TY(y + t)) * kk[r + t];
    }
    B[rl * nc + ci] = s;
  }
  __syncthreads();
  uint8_t* dst = out + (size_t)m.out_index * ps * ps;
  for (int it = threadIdx.x; it < nout * ps; it += blockDim.x) {
    const int jl = it / ps, i = it - jl * ps, j = j0 + jl;
    const float WX = P[i], WY = P[j];
    const int x = (int)floorf(WX), y = (int)floorf(WY);
    float v = 0.f;
    if (WX >= 0 && WY >= 0 && x < R - 1 && y < R - 1) {
      const float* r0 = B + (2 * jl) * nc + 2 * i;
      const float* r1 = r0 + nc;
      const float wx = WX - (float)x;
      const float I1 = wx * (r0[1] - r0[0]) + r0[0];
      v = (WY - (float)y) * (wx * (r1[1] - r1[0]) + r1[0] - I1) + I1;
    }
    int q = __float2int_rn(v);
    dst[j * ps + i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

constexpr int SMEM_SMALL = (MAX_PS + 2 * SMALL_R + 32 + 2 * SMALL_R * SMALL_R) * 4;
constexpr int SMEM_L3 = (MAX_PS + 640 + 2 * L3_OUT_ROWS * 2 * MAX_PS) * 4;

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
  // prefix sums of the per-region block counts of the two row-blocked phases
  const int nl = (int)large.size();
  std::vector<int> pre1(nl + 1, 0), pre2(nl + 1, 0);
  int maxR = 0;
  for (int i = 0; i < nl; i++) {
    pre1[i + 1] = pre1[i] + ceil_div(large[i].R, L1_ROWS);
    pre2[i + 1] = pre2[i] + ceil_div(large[i].R, L2_ROWS);
    maxR = std::max(maxR, large[i].R);
  }
  const size_t nm = small.size() + large.size();
  const size_t meta_bytes = nm * sizeof(PatchMeta), taps_bytes = taps_all.size() * 4, pre_bytes = (size_t)(nl + 1) * 4;
  MG_CUDA(ctx, ctx->h_stage2.ensure(meta_bytes + taps_bytes + 2 * pre_bytes + 64));
  uint8_t* hb = ctx->h_stage2.as<uint8_t>();
  if (!small.empty()) memcpy(hb, small.data(), small.size() * sizeof(PatchMeta));
  if (!large.empty()) memcpy(hb + small.size() * sizeof(PatchMeta), large.data(), large.size() * sizeof(PatchMeta));
  if (!taps_all.empty()) memcpy(hb + meta_bytes, taps_all.data(), taps_bytes);
  memcpy(hb + meta_bytes + taps_bytes, pre1.data(), pre_bytes);
  memcpy(hb + meta_bytes + taps_bytes + pre_bytes, pre2.data(), pre_bytes);
  // one upload: [metas | taps | pre1 | pre2]
  MG_CUDA(ctx, ctx->smp_meta.ensure(meta_bytes + taps_bytes + 2 * pre_bytes + 64));
  MG_CUDA(ctx, ctx->smp_scratch.ensure((size_t)scratch * 4 + 16));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_meta.p, hb, meta_bytes + taps_bytes + 2 * pre_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const PatchMeta* dm = ctx->smp_meta.as<PatchMeta>();
  const float* dtaps = reinterpret_cast<const float*>(ctx->smp_meta.as<uint8_t>() + meta_bytes);
  const int* dpre1 = reinterpret_cast<const int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes + taps_bytes);
  const int* dpre2 = reinterpret_cast<const int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes + taps_bytes + pre_bytes);
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_large_rowpass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (MAX_PS + 640 + L2_ROWS * MAX_R) * 4));
    attr_set = true;
  }
  double bytes_small = 0, bytes_large = 0;   // algorithmic: R*R*4 read + ps*ps written per region
  for (const PatchMeta& m : small) bytes_small += (double)m.R * m.R * 4.0 + (double)ps * ps;
  for (const PatchMeta& m : large) bytes_large += (double)m.R * m.R * 4.0 + (double)ps * ps;
  if (!small.empty()) {
    MG_PROF(ctx, "k_sample_small", 0, bytes_small);
    k_sample_small<<<(unsigned)small.size(), 128, SMEM_SMALL, ctx->stream>>>(img->d, img->w, img->h, dm, dtaps, d_out, ps);
    MG_LAUNCHED(ctx);
  }
  if (nl > 0) {
    const PatchMeta* dl = dm + small.size();
    float* scr = ctx->smp_scratch.as<float>();
    MG_PROF(ctx, "k_large_resample", 0, bytes_large);
    k_large_resample<<<pre1[nl], 128, 0, ctx->stream>>>(img->d, img->w, img->h, dl, nl, dpre1, scr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_rowpass", 2, (double)nl);
    k_large_rowpass<<<pre2[nl], 256, (MAX_PS + 640 + L2_ROWS * maxR) * 4, ctx->stream>>>(dl, nl, dpre2, dtaps, scr, ps);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_colpass_final", 2, (double)nl);
    k_large_colpass_final<<<nl * ceil_div(ps, L3_OUT_ROWS), 256, SMEM_L3, ctx->stream>>>(dl, dtaps, scr, d_out, ps);
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
