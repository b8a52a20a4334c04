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
               const float* __restrict__ taps_all, uint8_t* __restrict__ out, int ps, float* __restrict__ outf,
               const int* __restrict__ cnt) {
  extern __shared__ float sm[];
  if (cnt != nullptr && (int)blockIdx.x >= *cnt) return;      // grid sized by the host's upper bound (device-side prep)
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
        if (outf) outf[(size_t)m.out_index * ps * ps + j * ps + i] = v;
        else dst[j * ps + i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
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
    if (outf) outf[(size_t)m.out_index * ps * ps + it] = v;
    else dst[it] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

// =================================================================================================
// Register-blocked kernels (the common case: ks >= 7).
//
// What bounds the sampler is not HBM: the image (3 MB) is L2 resident and a region costs R^2 bilinear gathers plus
// ~18 R^2 (pairs form) / 2 ks R^2 (full form) sequentially ordered FMAs.  Measured (profiles/r02_ncu_hot_k_sample_*.txt):
// class B issues on two thirds of the cycles, the rest is split between barriers, gather latency, shared-memory latency
// inside the FMA chains and the sequentially rounded coordinate replay; class A is barrier bound.  Shared-memory bank
// conflicts matter more than instruction counts (the class B row pass lost 20 % to them).  The rules that shape the kernels:
//  (1) gathers are issued by 8 lanes walking 8 CONSECUTIVE samples of a row (a warp = 4 such groups, usually
//      32 consecutive samples): a load touches 2-4 cache lines instead of 32.  The reference accumulates
//      sample coordinates sequentially in float (WX += a11), so one thread per row first walks its whole row
//      and records the coordinate of every 8th (4th: class B1) sample (C2 table); a lane then replays <= 7 (3) additions.
//      The same walk proves a row interior (first and last sample inside the image => all of them: the sequence is
//      monotone), and interior rows are sampled without any bounds handling.
//  (2) the separable blur is register blocked: a thread owns CB adjacent columns x RB rows, loads one tap and
//      RB samples per step and issues RB*CB FMAs on them (taps rotate through registers), instead of
//      two shared-memory loads and a clamp per FMA.  Borders are replicated into padding so the inner loops
//      have no clamps; zero taps appended to the tap array keep the FMA chains bit-exact (fma(d, 0, s) == s).
//      Lanes are laid out so that a warp's shared-memory accesses fall into 32 different banks (class B: lane = row of
//      the block with an odd row pitch).
//  (3) CTA size follows the work per phase: 128 threads for R <= 40, 256 above.
// Only the columns/rows the final 32x32 resampling reads are filtered when R >= 66 (pairs x_i, x_i + 1).
// =================================================================================================
constexpr int SEG = 8;
constexpr int NT = 256;

__device__ __forceinline__ int fast_div(int a, float inv_b) { return (int)(((float)a + 0.5f) * inv_b); }

// Per-row coordinate table.  One thread per row walks the float coordinate sequence of interpolate() (helpers.cpp:551-626:
// WX += a11 per sample, sequentially rounded -- that sequence is the contract) and keeps every TS-th coordinate; a lane of
// the samplers below replays at most TS - 1 additions from the nearest entry.  TS = 4 where the table fits (classes A, B1
// and the large-window slab), 8 for B2.
// The walker also decides whether the row is INTERIOR: the sequence is monotone in x and in y (fl(x + a) >= x for a >= 0),
// so if its first and last sample pass the bounds test of interpolate(), every sample of the row does, and the samplers
// skip the per-sample test, the selects and the predicate spills that came with them (rows of a warp vote).
__device__ __forceinline__ bool sample_inside(float WX, float WY, int w, int h) {
  return WX >= 0 && WY >= 0 && (int)floorf(WX) < w - 1 && (int)floorf(WY) < h - 1;
}
template <int TS>
__device__ __forceinline__ bool gen_row_starts(const PatchMeta& m, int j, int R, int w, int h, float2* __restrict__ dst) {
  const int half = R / 2, nte = (R + TS - 1) / TS;
  float rx = m.x - (float)half * m.a12, ry = m.y - (float)half * m.a22;
  for (int t = 0; t < j; t++) { rx += m.a12; ry += m.a22; }
  float WX = rx - (float)half * m.a11, WY = ry - (float)half * m.a21;
  bool ok = sample_inside(WX, WY, w, h);
  for (int e = 0; e < nte - 1; e++) {
    dst[e] = make_float2(WX, WY);
#pragma unroll
    for (int t = 0; t < TS; t++) { WX += m.a11; WY += m.a21; }
  }
  dst[nte - 1] = make_float2(WX, WY);
  for (int i = (nte - 1) * TS; i < R - 1; i++) { WX += m.a11; WY += m.a21; }      // on to the row's last sample
  return ok && sample_inside(WX, WY, w, h);
}

// one sample: `rep` sequential additions from table entry c, then interpolate()'s bilinear form.
// INTERIOR: the row passed the walker's test, no bounds handling.  Otherwise branch-free (out-of-image samples read pixel
// (0,0) and are zeroed by the select) so that the gathers of several unrolled segments can be in flight together.
template <int TS, bool INTERIOR>
__device__ __forceinline__ float sample_seg(const float* __restrict__ img, int w, int h, float2 c, float a11, float a21, int rep) {
  float WX = c.x, WY = c.y;
#pragma unroll
  for (int t = 0; t < TS - 1; t++)
    if (t < rep) { WX += a11; WY += a21; }
  if (INTERIOR) {
    const int x = (int)WX, y = (int)WY;                  // both >= 0: truncation is floor
    const float* p = img + (y * w + x);
    const float v00 = p[0], v01 = p[1], v10 = p[w], v11 = p[w + 1];
    const float wx = WX - (float)x;
    const float I1 = wx * (v01 - v00) + v00;
    return (WY - (float)y) * (wx * (v11 - v10) + v10 - I1) + I1;
  } else {
    const int x = (int)floorf(WX), y = (int)floorf(WY);
    const bool ok = WX >= 0 && WY >= 0 && x < w - 1 && y < h - 1;
    const float v = bilinear(img, w, ok ? x : 0, ok ? y : 0, ok ? WX : 0.f, ok ? WY : 0.f);
    return ok ? v : 0.f;
  }
}

// one row (or row block) of the first resampling: an 8-lane group walks the row in segments of 8 consecutive samples, four
// segments per step -- coordinates first, then the 4 x 4 gathers, then the stores (the stores may alias the coordinate
// table as far as the compiler knows, so the order is spelled out).  Lanes beyond the row's end re-sample the last table
// entry (a real sample of the row, so the interior path never reads outside the image) and store nothing.
template <int TS, bool INTERIOR, int UNR>
__device__ __forceinline__ void sample_row_impl(const float* __restrict__ img, int w, int h, const float2* cs, int nseg, float a11, float a21,
                                                int sub, float* row, int R) {
  const int nte = (R + TS - 1) / TS;
  float2 c[UNR];
  int rep[UNR];
  auto fetch = [&](int s0) {               // table entries of the UNR segments starting at s0
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = (s0 + u) * SEG + sub;
      const bool live = i < R;
      c[u] = cs[live ? i / TS : nte - 1];
      rep[u] = live ? i % TS : 0;
    }
  };
  // PIPE (slab path, table in global memory): the next step's entries travel while this step's gathers do.  With the table
  // in shared memory the loop-carried registers cost class B 8 %, so there every step fetches its own entries first.
  constexpr bool PIPE = UNR == 8;
  if (PIPE) fetch(0);
  for (int s0 = 0; s0 < nseg; s0 += UNR) {
    if (!PIPE) fetch(s0);
    float v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) v[u] = sample_seg<TS, INTERIOR>(img, w, h, c[u], a11, a21, rep[u]);
    if (PIPE && s0 + UNR < nseg) fetch(s0 + UNR);
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const int i = (s0 + u) * SEG + sub;
      if (i < R) row[i] = v[u];
    }
  }
}
// `fast` must be warp uniform (the caller votes over the rows of its warp at a convergent point)
template <int TS, int UNR = 4>
__device__ __forceinline__ void sample_row(const float* __restrict__ img, int w, int h, const float2* cs, int nseg, float a11, float a21,
                                           int sub, float* row, int R, bool fast) {
  if (fast) sample_row_impl<TS, true, UNR>(img, w, h, cs, nseg, a11, a21, sub, row, R);
  else sample_row_impl<TS, false, 4>(img, w, h, cs, nseg, a11, a21, sub, row, R);
}

// Row pass, vector form (x < R & ~3, ks >= 7): s = 0; s = fma(p[t], k[t], s), t = 0..ks-1, for CB adjacent
// columns x0..x0+CB-1 of RB rows.  rows[b] points at padded index x0 (= sample x0 - r); kp has CB-1 zeros appended.
template <int CB, int RB>
__device__ __forceinline__ void rowpass_block(const float* (&rows)[RB], const float* __restrict__ kp, int ks,
                                              float (&acc)[RB][CB]) {
  float kq[CB];
#pragma unroll
  for (int c = 0; c < CB; c++) kq[c] = 0.f;
#pragma unroll
  for (int b = 0; b < RB; b++)
#pragma unroll
    for (int c = 0; c < CB; c++) acc[b][c] = 0.f;
  const int nstep = ks + CB - 1;
#pragma unroll 4
  for (int u = 0; u < nstep; u++) {
#pragma unroll
    for (int c = CB - 1; c > 0; c--) kq[c] = kq[c - 1];
    kq[0] = kp[u];
#pragma unroll
    for (int b = 0; b < RB; b++) {
      const float d = rows[b][u];
#pragma unroll
      for (int c = 0; c < CB; c++) acc[b][c] = fmaf(d, kq[c], acc[b][c]);
    }
  }
}

// Column pass for RB vertically adjacent outputs y0..y0+RB-1 of one column; Tc points at padded row y0 (+r).
// vec (x < R & ~7): s = T[y]*k[r]; s = fma(T[y-t] + T[y+t], k[r+t], s); else the unfused scalar tail.
template <int RB>
__device__ __forceinline__ void colpass_block(const float* __restrict__ Tc, int pitch, const float* __restrict__ k, int r,
                                              bool vec, float (&out)[RB]) {
  float dn[RB], up[RB];
#pragma unroll
  for (int i = 0; i < RB; i++) { dn[i] = up[i] = Tc[i * pitch]; out[i] = dn[i] * k[r]; }
#pragma unroll 2
  for (int t = 1; t <= r; t++) {
#pragma unroll
    for (int i = RB - 1; i > 0; i--) dn[i] = dn[i - 1];
    dn[0] = Tc[-t * pitch];
#pragma unroll
    for (int i = 0; i < RB - 1; i++) up[i] = up[i + 1];
    up[RB - 1] = Tc[(RB - 1 + t) * pitch];
    const float kt = k[r + t];
    if (vec) {
#pragma unroll
      for (int i = 0; i < RB; i++) out[i] = fmaf(dn[i] + up[i], kt, out[i]);
    } else {
#pragma unroll
      for (int i = 0; i < RB; i++) out[i] = out[i] + (dn[i] + up[i]) * kt;
    }
  }
}

__device__ __forceinline__ uint8_t quant_u8(float v) {
  int q = __float2int_rn(v);
  return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

// ---- class A: R <= 65, the whole R x R window is filtered ------------------------------------------------
constexpr int A_TS = 8;      // coordinate table spacing of class A (4 was measured: fewer replay additions, no gain -- the phase is gather-latency bound)
__host__ __device__ inline int a_smem_floats(int R, int r) {
  const int nte = (R + A_TS - 1) / A_TS, PS = (R + 2 * r + 4) | 1, PT = R | 1;
  return 64 + 64 + 72 + 2 * R * nte + R * PS + (R + 2 * r + 4) * PT;
}

template <int NT>
__global__ void __launch_bounds__(NT)
k_sample_a(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas,
           const float* __restrict__ taps_all, uint8_t* __restrict__ out, int ps, float* __restrict__ outf,
           const int* __restrict__ cnt) {
  extern __shared__ float sm[];
  if (cnt != nullptr && (int)blockIdx.x >= *cnt) return;
  const PatchMeta m = metas[blockIdx.x];
  const int tid = threadIdx.x;
  const int R = m.R, ks = m.ks, r = ks >> 1;
  const int nseg = (R + SEG - 1) / SEG, nte = (R + A_TS - 1) / A_TS, PS = (R + 2 * r + 4) | 1, PT = R | 1;
  float* P = sm;
  float* kp = P + 64;
  int* rowok = reinterpret_cast<int*>(kp + 64);       // 72: per-row interior flags
  float2* C2 = reinterpret_cast<float2*>(kp + 64 + 72);
  float* Sp = reinterpret_cast<float*>(C2 + R * nte);
  float* Tp = Sp + R * PS;
  float* B = Sp;   // R x R, written after the row pass has consumed Sp

  if (tid < ks + 3) kp[tid] = tid < ks ? taps_all[m.tap_off + tid] : 0.f;
  if (tid == NT - 1) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; p += m.scale; }
  }
  if (tid < R) rowok[tid] = gen_row_starts<A_TS>(m, tid, R, w, h, C2 + tid * nte);
  __syncthreads();
  // 1. first resampling: an 8-lane group walks one row segment by segment (no index division per sample)
  {
    const int g = tid >> 3, sub = tid & 7;
    for (int j0 = 0; j0 < R; j0 += NT / SEG) {
      const int j = j0 + g;
      const bool act = j < R;
      const bool fast = __all_sync(0xffffffffu, act ? rowok[j] != 0 : true);
      if (act) sample_row<A_TS>(img, w, h, C2 + j * nte, nseg, m.a11, m.a21, sub, Sp + j * PS + r, R, fast);
    }
  }
  __syncthreads();
  // replicate the borders into the padding (all threads; r left, r + 4 right entries per row)
  {
    const int np = 2 * r + 4;
    const float inv = 1.0f / (float)np;
    for (int it = tid; it < R * np; it += NT) {
      const int j = fast_div(it, inv), t = it - j * np;
      float* row = Sp + j * PS;
      if (t < r) row[t] = row[r];
      else row[R + t] = row[r + R - 1];
    }
  }
  __syncthreads();
  // 2. row pass: blocks of 4 columns x 4 rows (rows rg + k*NRG: conflict-free with the odd pitch)
  {
    const int ncb = R >> 2, NRG = (R + 3) >> 2, nmain = ncb * NRG;
    const float inv = 1.0f / (float)NRG;
    for (int it = tid; it < nmain; it += NT) {
      const int cb = fast_div(it, inv), rg = it - cb * NRG, x0 = cb * 4;
      const float* rows[4];
      int rowi[4];
#pragma unroll
      for (int b = 0; b < 4; b++) { rowi[b] = rg + b * NRG; rows[b] = Sp + min(rowi[b], R - 1) * PS + x0; }
      float acc[4][4];
      rowpass_block<4, 4>(rows, kp, ks, acc);
#pragma unroll
      for (int b = 0; b < 4; b++)
        if (rowi[b] < R) {
          float* t = Tp + (rowi[b] + r) * PT + x0;
#pragma unroll
          for (int c = 0; c < 4; c++) t[c] = acc[b][c];
        }
    }
    // scalar tail columns x >= R & ~3
    const int xt0 = R & ~3, nt = R - xt0;
    for (int it = tid; it < nt * R; it += NT) {
      const int y = it / nt, x = xt0 + (it - y * nt);
      Tp[(y + r) * PT + x] = row_pass_at(Sp + y * PS + r, R, x, kp, ks);
    }
  }
  __syncthreads();
  // 3. replicate T above and below
  for (int it = tid; it < (2 * r + 4) * R; it += NT) {
    const int pr = it / R, x = it - pr * R;
    if (pr < r) Tp[pr * PT + x] = Tp[r * PT + x];
    else Tp[(R + pr) * PT + x] = Tp[(R - 1 + r) * PT + x];
  }
  __syncthreads();
  // 4. column pass: 4 adjacent rows per thread, lanes along x; vector columns (x < R & ~7) and the scalar tail
  //    columns are separate item lists so that no warp executes both forms
  {
    const int NRG = (R + 3) >> 2, wc = R & ~7, nt = R - wc;
    if (wc > 0) {
      const float inv = 1.0f / (float)wc;
      for (int it = tid; it < wc * NRG; it += NT) {
        const int rg = fast_div(it, inv), x = it - rg * wc, y0 = rg * 4;
        float o[4];
        colpass_block<4>(Tp + (y0 + r) * PT + x, PT, kp, r, true, o);
#pragma unroll
        for (int i = 0; i < 4; i++) if (y0 + i < R) B[(y0 + i) * R + x] = o[i];
      }
    }
    for (int it = tid; it < nt * NRG; it += NT) {
      const int rg = it / nt, x = wc + (it - rg * nt), y0 = rg * 4;
      float o[4];
      colpass_block<4>(Tp + (y0 + r) * PT + x, PT, kp, r, false, o);
#pragma unroll
      for (int i = 0; i < 4; i++) if (y0 + i < R) B[(y0 + i) * R + x] = o[i];
    }
  }
  __syncthreads();
  // 5. second resampling + u8
  uint8_t* dst = out + (size_t)m.out_index * ps * ps;
  for (int it = tid; it < ps * ps; it += NT) {
    const int j = it / ps, i = it - j * ps;
    const float WX = P[i], WY = P[j];
    const int x = (int)floorf(WX), y = (int)floorf(WY);
    float v = 0.f;
    if (WX >= 0 && WY >= 0 && x < R - 1 && y < R - 1) {
      const float* r0 = B + y * R + x;
      const float* r1 = r0 + R;
      const float wx = WX - (float)x;
      const float I1 = wx * (r0[1] - r0[0]) + r0[0];
      v = (WY - (float)y) * (wx * (r1[1] - r1[0]) + r1[0] - I1) + I1;
    }
    if (outf) outf[(size_t)m.out_index * ps * ps + it] = v;
    else dst[it] = quant_u8(v);
  }
}

// ---- class B: 66 <= R <= 160 (scale >= 2): only the 64 columns / rows the final resampling reads ----------
constexpr int B_NB = 32;     // rows sampled + row-filtered per iteration
constexpr int B_TP = 65;     // pitch of T (64 needed columns)
__host__ __device__ inline int b_smem_floats(int R, int r, int ts) {
  const int nte = (R + ts - 1) / ts, PS = (R + 2 * r + 2) | 1;
  int u = 2 * R * nte + B_NB * PS;      // C2 + S block, later reused for B (64 x 64)
  if (u < 4096) u = 4096;
  return 64 + 64 + 64 + 160 + u + (R + 2 * r + 2) * B_TP;
}

template <int TS, int UNR>
__global__ void __launch_bounds__(NT)
k_sample_b(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas,
           const float* __restrict__ taps_all, uint8_t* __restrict__ out, const int* __restrict__ cnt) {
  constexpr int ps = 32;
  extern __shared__ float sm[];
  if (cnt != nullptr && (int)blockIdx.x >= *cnt) return;
  const PatchMeta m = metas[blockIdx.x];
  const int tid = threadIdx.x;
  const int R = m.R, ks = m.ks, r = ks >> 1;
  const int nseg = (R + SEG - 1) / SEG, nte = (R + TS - 1) / TS, PS = (R + 2 * r + 2) | 1;
  float* P = sm;
  int* X = reinterpret_cast<int*>(P + 64);
  float* kp = P + 128;
  int* rowok = reinterpret_cast<int*>(kp + 64);     // 160: per-row interior flags
  float2* C2 = reinterpret_cast<float2*>(kp + 64 + 160);
  float* Sb = reinterpret_cast<float*>(C2 + R * nte);
  float* B = reinterpret_cast<float*>(C2);          // 64 x 64 once C2 / Sb are dead
  int un = 2 * R * nte + B_NB * PS;
  if (un < 4096) un = 4096;
  float* Tp = reinterpret_cast<float*>(C2) + un;

  if (tid < ks + 1) kp[tid] = tid < ks ? taps_all[m.tap_off + tid] : 0.f;
  if (tid == NT - 1) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; X[i] = (int)floorf(p); p += m.scale; }
  }
  if (tid < R) rowok[tid] = gen_row_starts<TS>(m, tid, R, w, h, C2 + tid * nte);
  __syncthreads();
  const int xv = R & ~3;
  for (int j0 = 0; j0 < R; j0 += B_NB) {
    const int nrows = min(B_NB, R - j0);
    // 1. sample nrows rows: group g (8 lanes) walks row j0 + g segment by segment
    {
      const int g = tid >> 3, sub = tid & 7;
      const bool act = g < nrows;
      const bool fast = __all_sync(0xffffffffu, act ? rowok[j0 + g] != 0 : true);
      if (act) sample_row<TS, UNR>(img, w, h, C2 + (j0 + g) * nte, nseg, m.a11, m.a21, sub, Sb + g * PS + r, R, fast);
    }
    __syncthreads();
    // replicate the borders into the padding (r left, r + 2 right entries per row)
    {
      const int np = 2 * r + 2;
      const float inv = 1.0f / (float)np;
      for (int it = tid; it < nrows * np; it += NT) {
        const int jr = fast_div(it, inv), t = it - jr * np;
        float* row = Sb + jr * PS;
        if (t < r) row[t] = row[r];
        else row[R + t] = row[r + R - 1];
      }
    }
    __syncthreads();
    // 2. row pass at the 32 column pairs (x_i, x_i + 1).  LANE = ROW of the block, a warp owns the pairs w, w + 8, w + 16,
    //    w + 24: with the odd row pitch every sample load and every T store of a warp is bank-conflict free.  (The first
    //    mapping -- 4 pairs x 8 rows per warp -- measured 2.2 wavefronts per sample load and 3.8 per store: the row
    //    pass sits on the shared-memory pipe, so that was most of its time.)
    {
      const int wq = tid >> 5, row = tid & 31, rowc = min(row, nrows - 1);
      const float* rows[4];
      int x0[4];
#pragma unroll
      for (int b = 0; b < 4; b++) { x0[b] = X[wq + 8 * b]; rows[b] = Sb + rowc * PS + x0[b]; }
      float acc[4][2];
      rowpass_block<2, 4>(rows, kp, ks, acc);
      if (row < nrows) {
        float* t = Tp + (j0 + row + r) * B_TP + 2 * wq;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          if (x0[b] + 1 < xv) {           // warp uniform
            t[16 * b] = acc[b][0]; t[16 * b + 1] = acc[b][1];
          } else {                        // OpenCV's scalar tail columns (the last pair at most)
            const float* rp = Sb + row * PS + r;
            t[16 * b] = row_pass_at(rp, R, x0[b], kp, ks);
            t[16 * b + 1] = row_pass_at(rp, R, x0[b] + 1, kp, ks);
          }
        }
      }
    }
    __syncthreads();
  }
  // 3. replicate T above and below
  for (int it = tid; it < (2 * r + 2) * 64; it += NT) {
    const int pr = it >> 6, c = it & 63;
    if (pr < r) Tp[pr * B_TP + c] = Tp[r * B_TP + c];
    else Tp[(R + pr) * B_TP + c] = Tp[(R - 1 + r) * B_TP + c];
  }
  __syncthreads();
  // 4. column pass at the 32 row pairs (y_j, y_j + 1) x 64 columns: the fused (vector) form for every column first,
  //    then the few columns in OpenCV's scalar tail (x >= R & ~7) are redone with the unfused form -- a per-lane
  //    branch inside the tap loop would make half of the warps execute both forms
  {
    const int wc = R & ~7;
    for (int it = tid; it < 32 * 64; it += NT) {
      const int pj = it >> 6, ci = it & 63;
      const int y0 = X[pj];
      float o[2];
      colpass_block<2>(Tp + (y0 + r) * B_TP + ci, B_TP, kp, r, true, o);
      B[(2 * pj) * 64 + ci] = o[0];
      B[(2 * pj + 1) * 64 + ci] = o[1];
    }
    int ntail = 0;                       // needed columns are ascending: the tail is a suffix
    while (ntail < 64 && X[(63 - ntail) >> 1] + ((63 - ntail) & 1) >= wc) ntail++;
    __syncthreads();
    for (int it = tid; it < 32 * ntail; it += NT) {
      const int pj = it / ntail, ci = 64 - ntail + (it - pj * ntail);
      const int y0 = X[pj];
      float o[2];
      colpass_block<2>(Tp + (y0 + r) * B_TP + ci, B_TP, kp, r, false, o);
      B[(2 * pj) * 64 + ci] = o[0];
      B[(2 * pj + 1) * 64 + ci] = o[1];
    }
  }
  __syncthreads();
  // 5. second resampling + u8
  uint8_t* dst = out + (size_t)m.out_index * ps * ps;
  for (int it = tid; it < ps * ps; it += NT) {
    const int j = it >> 5, i = it & 31;
    const float WX = P[i], WY = P[j];
    const int x = X[i], y = X[j];
    float v = 0.f;
    if (WX >= 0 && WY >= 0 && x < R - 1 && y < R - 1) {
      const float* r0 = B + (2 * j) * 64 + 2 * i;
      const float* r1 = r0 + 64;
      const float wx = WX - (float)x;
      const float I1 = wx * (r0[1] - r0[0]) + r0[0];
      v = (WY - (float)y) * (wx * (r1[1] - r1[0]) + r1[0] - I1) + I1;
    }
    dst[it] = quant_u8(v);
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

constexpr int L0_ROWS = 128, L1_ROWS = 16, L2_ROWS = 16, L3_OUT_ROWS = 4;

// per-region scratch (floats): [C: R x nseg float2 row-segment start coordinates | S: R x R | T: R x 2ps]
constexpr int L_TS = 8;      // table spacing of the slab path; every table row carries one extra entry: .x = interior flag
__host__ __device__ inline int large_row_entries(int R) { return (R + L_TS - 1) / L_TS + 1; }
__host__ __device__ inline long long large_c_floats(int R) { return 2LL * R * large_row_entries(R); }

// phase 0: one thread per window row walks the row's float coordinate sequence once (the sequential `WX += a11` of
// interpolate(), helpers.cpp:551-626, is what makes the samples bit-exact) and keeps every SEG-th coordinate
__global__ void __launch_bounds__(L0_ROWS)
k_large_starts(const PatchMeta* __restrict__ metas, int nreg, const int* __restrict__ pre, float* __restrict__ scratch,
               const int* __restrict__ nreg_dev, int w, int h) {
  if (nreg_dev != nullptr) { nreg = *nreg_dev; if ((int)blockIdx.x >= pre[nreg]) return; }
  const int reg = find_region(pre, nreg, blockIdx.x);
  const PatchMeta m = metas[reg];
  const int R = m.R, nent = large_row_entries(R);
  const int j = (blockIdx.x - pre[reg]) * L0_ROWS + threadIdx.x;
  if (j < R) {
    float2* dst = reinterpret_cast<float2*>(scratch + m.scratch_off) + (size_t)j * nent;
    const bool ok = gen_row_starts<L_TS>(m, j, R, w, h, dst);
    dst[nent - 1] = make_float2(ok ? 1.f : 0.f, 0.f);
  }
}

// phase 1: S[j][i] for L1_ROWS rows of one region; an 8-lane group walks one row segment by segment (as class A / B)
__global__ void __launch_bounds__(128)
k_large_resample(const float* __restrict__ img, int w, int h, const PatchMeta* __restrict__ metas, int nreg,
                 const int* __restrict__ pre, float* __restrict__ scratch, const int* __restrict__ nreg_dev) {
  if (nreg_dev != nullptr) { nreg = *nreg_dev; if ((int)blockIdx.x >= pre[nreg]) return; }
  const int reg = find_region(pre, nreg, blockIdx.x);
  const PatchMeta m = metas[reg];
  const int R = m.R, nseg = (R + SEG - 1) / SEG, nent = large_row_entries(R);
  const int j = (blockIdx.x - pre[reg]) * L1_ROWS + (threadIdx.x >> 3), sub = threadIdx.x & 7;
  const bool act = j < R;
  const float2* cs = reinterpret_cast<const float2*>(scratch + m.scratch_off) + (size_t)(act ? j : 0) * nent;
  const bool fast = __all_sync(0xffffffffu, act ? cs[nent - 1].x != 0.f : true);
  if (!act) return;
  float* row = scratch + m.scratch_off + large_c_floats(R) + (size_t)j * R;
  sample_row<L_TS, 8>(img, w, h, cs, nseg, m.a11, m.a21, sub, row, R, fast);      // no shared memory here: 8 segments in flight per lane
}

// phase 2a: row pass at the 2*ps needed columns for L2_ROWS rows of one region: T[y][ci].  Rows are staged in shared
// memory with their replicated borders, then the register-blocked vector form of class B (a column pair x 2 rows per
// thread); the few columns in OpenCV's scalar tail go through row_pass_at.
__global__ void __launch_bounds__(256)
k_large_rowpass(const PatchMeta* __restrict__ metas, int nreg, const int* __restrict__ pre,
                const float* __restrict__ taps_all, float* __restrict__ scratch, int ps, const int* __restrict__ nreg_dev) {
  extern __shared__ float sm[];
  if (nreg_dev != nullptr) { nreg = *nreg_dev; if ((int)blockIdx.x >= pre[nreg]) return; }
  const int reg = find_region(pre, nreg, blockIdx.x);
  const PatchMeta m = metas[reg];
  const int R = m.R, ks = m.ks, r = ks >> 1, nc = 2 * ps, PS = (R + 2 * r + 2) | 1;
  const int row0 = (blockIdx.x - pre[reg]) * L2_ROWS;
  const int nrows = min(L2_ROWS, R - row0);
  float* P = sm;                                      // ps
  int* X = reinterpret_cast<int*>(P + MAX_PS);        // ps
  float* kk = P + 2 * MAX_PS;                         // ks + 1 <= 642
  float* rows = kk + 642;                             // nrows x PS (padded)
  const float* S = scratch + m.scratch_off + large_c_floats(R);
  float* T = scratch + m.scratch_off + large_c_floats(R) + (size_t)R * R;
  for (int i = threadIdx.x; i < ks + 1; i += blockDim.x) kk[i] = i < ks ? taps_all[m.tap_off + i] : 0.f;
  if (threadIdx.x == 0) {
    float p = (float)(R / 2) - (float)(ps / 2) * m.scale;
    for (int i = 0; i < ps; i++) { P[i] = p; X[i] = (int)floorf(p); p += m.scale; }
  }
  for (int jr = threadIdx.x >> 5; jr < nrows; jr += 8) {          // a warp per row: coalesced reads, clamped index = border replication
    const float* src = S + (size_t)(row0 + jr) * R;
    float* dst = rows + jr * PS;
    // four loads in flight per lane before the first store (the plain loop was one L2 round trip per iteration: 45 % of
    // this kernel's stall samples)
    const int len = R + 2 * r + 2;
    for (int t0 = threadIdx.x & 31; t0 < len; t0 += 128) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = src[clampi(min(t0 + 32 * q, len - 1) - r, 0, R - 1)];
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (t0 + 32 * q < len) dst[t0 + 32 * q] = v[q];
    }
  }
  __syncthreads();
  const int xv = R & ~3;
  const int rg = threadIdx.x & 7;
  for (int pi = threadIdx.x >> 3; pi < ps; pi += 32) {
    const int x0 = X[pi];
    int rowi[2] = {rg, rg + 8};
    if (x0 >= 0 && x0 + 1 < xv && ks >= 7) {
      const float* rp[2];
#pragma unroll
      for (int b = 0; b < 2; b++) rp[b] = rows + min(rowi[b], nrows - 1) * PS + x0;
      float acc[2][2];
      rowpass_block<2, 2>(rp, kk, ks, acc);
#pragma unroll
      for (int b = 0; b < 2; b++)
        if (rowi[b] < nrows) {
          float* t = T + (size_t)(row0 + rowi[b]) * nc + 2 * pi;
          t[0] = acc[b][0]; t[1] = acc[b][1];
        }
    } else {
#pragma unroll 1
      for (int b = 0; b < 2; b++)
        if (rowi[b] < nrows) {
          const float* row = rows + rowi[b] * PS + r;
          float* t = T + (size_t)(row0 + rowi[b]) * nc + 2 * pi;
          t[0] = row_pass_at(row, R, needed_pos(P[pi], 0, R), kk, ks);
          t[1] = row_pass_at(row, R, needed_pos(P[pi], 1, R), kk, ks);
        }
    }
  }
}

// phase 2b + 3: column pass at the needed rows of L3_OUT_ROWS output rows, then the final resampling
__global__ void __launch_bounds__(256)
k_large_colpass_final(const PatchMeta* __restrict__ metas, const float* __restrict__ taps_all,
                      const float* __restrict__ scratch, uint8_t* __restrict__ out, int ps, float* __restrict__ outf,
                      const int* __restrict__ nreg_dev) {
  extern __shared__ float sm[];
  const int nblk = (ps + L3_OUT_ROWS - 1) / L3_OUT_ROWS;
  const int reg = blockIdx.x / nblk, j0 = (blockIdx.x - reg * nblk) * L3_OUT_ROWS;
  if (nreg_dev != nullptr && reg >= *nreg_dev) return;
  const PatchMeta m = metas[reg];
  const int R = m.R, ks = m.ks, nc = 2 * ps, r = ks >> 1;
  const int nout = min(L3_OUT_ROWS, ps - j0);
  float* P = sm;
  float* kk = P + MAX_PS;
  float* B = kk + 640;        // (2*nout) x nc
  const float* T = scratch + m.scratch_off + large_c_floats(R) + (size_t)R * R;
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
    // the symmetric pairs T[y - t] + T[y + t] of 8 taps are loaded (16 L2 / L1 reads in flight) before their FMAs: the
    // one-tap-at-a-time loop was 60 % of this kernel's stall samples
    const float* Tc = T + ci;
    float s = Tc[(size_t)y * nc] * kk[r];
    const bool vec = x < wc;
    for (int t0 = 1; t0 <= r; t0 += 8) {
      float lo[8], hi[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int t = min(t0 + q, r);
        lo[q] = Tc[(size_t)max(y - t, 0) * nc];
        hi[q] = Tc[(size_t)min(y + t, R - 1) * nc];
      }
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (t0 + q <= r) {
          if (vec) s = fmaf(lo[q] + hi[q], kk[r + t0 + q], s);
          else s = s + (lo[q] + hi[q]) * kk[r + t0 + q];
        }
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
    if (outf) outf[(size_t)m.out_index * ps * ps + j * ps + i] = v;
    else dst[j * ps + i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  }
}

constexpr int SMEM_SMALL = (MAX_PS + 2 * SMALL_R + 32 + 2 * SMALL_R * SMALL_R) * 4;
constexpr int SMEM_L3 = (MAX_PS + 640 + 2 * L3_OUT_ROWS * 2 * MAX_PS) * 4;

}  // namespace

// Enqueue the sampler for n regions (host array) on ctx->stream; u8 patches land in d_out
// (n * ps * ps bytes, region order) or, when d_outf is given, unquantised float patches in d_outf (what
// DescribeRegions feeds the SIFT descriptor, synth-detection.hpp:170-263).  No host synchronisation.
// Regions are dealt into size classes (one launch each, shared memory sized for the class so that small
// windows keep 4-8 CTAs per SM) and sorted by decreasing R inside a class (largest CTAs start first):
//   A1 R <= 40, A2 R <= 65 : k_sample_a (whole window filtered)        B1 R <= 100, B2 R <= 160 : k_sample_b
//   odd cases (direct mode, ks < 7, patchSize != 32 for B) : k_sample_small;  R > 160 : the 3-launch slab path
constexpr int A1_R = 40, A2_R = 65, B1_R = 100, B2_R = 160;
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
// class B1 keeps a coordinate table entry for every 4th sample (3 % faster than 8; 72 KB still gives 3 CTAs per SM), B2 for
// every 8th (a finer table would drop it to one CTA per SM).  8 instead of 4 segments in flight per lane: measured slower.
static int smp_b1_ts() { static const int v = env_int("MODSGPU_B1_TS", 4); return v == 8 ? 8 : 4; }
// class A1 (R <= 40) runs 128-thread CTAs: its phases have 100-400 work items, half of a 256-thread CTA sat at the barriers
static int smp_a1_nt() { static const int v = env_int("MODSGPU_A1_NT", 128); return v == 256 ? 256 : 128; }
template <typename... A>
static void launch_sample_a(bool a1, unsigned grid, int smem, cudaStream_t st, A... a) {
  static const int a2nt = env_int("MODSGPU_A2_NT", 256);
  if ((a1 && smp_a1_nt() == 128) || (!a1 && a2nt == 128)) k_sample_a<128><<<grid, 128, smem, st>>>(a...);
  else k_sample_a<256><<<grid, 256, smem, st>>>(a...);
}
template <typename... A>
static void launch_sample_b(bool b1, unsigned grid, int smem, cudaStream_t st, A... a) {
  if (b1 && smp_b1_ts() == 4) k_sample_b<4, 4><<<grid, NT, smem, st>>>(a...);
  else k_sample_b<8, 4><<<grid, NT, smem, st>>>(a...);
}

int mg_sample_enqueue(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                      double mrSize, int ps, uint8_t* d_out, float* d_outf) {
  if (ps < 2 || ps > MAX_PS) MG_FAIL(ctx, MODSGPU_EINVAL, "patchSize out of range");
  if (n <= 0) return 0;
  enum { C_SMALL = 0, C_A1, C_A2, C_B1, C_B2, C_LARGE, NCLS };
  // host scratch is per thread and keeps its capacity: this runs three times per image on the calling thread
  static thread_local std::vector<PatchMeta> cls[NCLS], sort_tmp;
  static thread_local std::vector<float> taps_all;
  static thread_local std::vector<int> tap_off_of, tap_ks_of, r_hist;   // indexed by R0 / R
  for (int c = 0; c < NCLS; c++) cls[c].clear();
  taps_all.clear();
  tap_off_of.assign(MAX_R + 4, -1);
  tap_ks_of.assign(MAX_R + 4, 0);
  int cls_r[NCLS] = {0, 0, 0, 0, 0, 0};           // max blur radius per class
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
    int c = C_SMALL;
    if (m.scale > 0.4) {
      m.R = R0 + 2;
      if (m.R > MAX_R) MG_FAIL(ctx, MODSGPU_EINVAL, "region too large for the sampler (R > 2048)");
      if (tap_off_of[R0] < 0) {
        std::vector<float> t;
        tap_ks_of[R0] = mg_gaussian_taps(1.5f * m.scale, t);
        tap_off_of[R0] = (int)taps_all.size();
        taps_all.insert(taps_all.end(), t.begin(), t.end());
      }
      m.tap_off = tap_off_of[R0]; m.ks = tap_ks_of[R0];
      if (m.ks > 600) MG_FAIL(ctx, MODSGPU_EINVAL, "sampler blur too wide");
      const bool blocked = m.ks >= 7 && m.ks <= 60 && ps <= 64;
      if (blocked && m.R <= A1_R) c = C_A1;
      else if (blocked && m.R <= A2_R) c = C_A2;
      else if (blocked && ps == 32 && !d_outf && R0 >= 2 * ps && m.R <= B1_R) c = C_B1;
      else if (blocked && ps == 32 && !d_outf && R0 >= 2 * ps && m.R <= B2_R) c = C_B2;
      else if (m.R <= SMALL_R && m.ks <= 31) c = C_SMALL;
      else c = C_LARGE;
    } else {
      m.R = 0;
    }
    cls[c].push_back(m);
    cls_r[c] = std::max(cls_r[c], m.ks >> 1);
  }
  // decreasing R inside a class, region order among equal R (a stable counting sort: R <= MAX_R)
  for (int c = C_A1; c <= C_LARGE; c++) {
    std::vector<PatchMeta>& v = cls[c];
    if (v.size() < 2) continue;
    r_hist.assign(MAX_R + 2, 0);
    for (const PatchMeta& m : v) r_hist[m.R]++;
    int pos = 0;
    for (int R = MAX_R; R >= 0; R--) { const int cnt = r_hist[R]; r_hist[R] = pos; pos += cnt; }
    sort_tmp.resize(v.size());
    for (const PatchMeta& m : v) sort_tmp[r_hist[m.R]++] = m;
    v.swap(sort_tmp);
  }
  std::vector<PatchMeta>& large = cls[C_LARGE];
  long long scratch = 0;
  for (PatchMeta& m : large) {
    m.scratch_off = scratch;
    scratch += large_c_floats(m.R) + (long long)m.R * m.R + (long long)m.R * 2 * ps;
    scratch += scratch & 1;          // the float2 table at the head of the next region stays 8-byte aligned
  }
  // prefix sums of the per-region block counts of the two row-blocked phases
  const int nl = (int)large.size();
  std::vector<int> pre0(nl + 1, 0), pre1(nl + 1, 0), pre2(nl + 1, 0);
  int maxR = 0, maxPS = 0;
  for (int i = 0; i < nl; i++) {
    pre0[i + 1] = pre0[i] + ceil_div(large[i].R, L0_ROWS);
    maxPS = std::max(maxPS, (large[i].R + 2 * (large[i].ks >> 1) + 2) | 1);
    pre1[i + 1] = pre1[i] + ceil_div(large[i].R, L1_ROWS);
    pre2[i + 1] = pre2[i] + ceil_div(large[i].R, L2_ROWS);
    maxR = std::max(maxR, large[i].R);
  }
  size_t nm = 0, cls_off[NCLS];
  for (int c = 0; c < NCLS; c++) { cls_off[c] = nm; nm += cls[c].size(); }
  const size_t meta_bytes = nm * sizeof(PatchMeta), taps_bytes = taps_all.size() * 4, pre_bytes = (size_t)(nl + 1) * 4;
  MG_CUDA(ctx, ctx->h_stage2.ensure(meta_bytes + taps_bytes + 3 * pre_bytes + 64));
  uint8_t* hb = ctx->h_stage2.as<uint8_t>();
  for (int c = 0; c < NCLS; c++)
    if (!cls[c].empty()) memcpy(hb + cls_off[c] * sizeof(PatchMeta), cls[c].data(), cls[c].size() * sizeof(PatchMeta));
  if (!taps_all.empty()) memcpy(hb + meta_bytes, taps_all.data(), taps_bytes);
  memcpy(hb + meta_bytes + taps_bytes, pre1.data(), pre_bytes);
  memcpy(hb + meta_bytes + taps_bytes + pre_bytes, pre2.data(), pre_bytes);
  memcpy(hb + meta_bytes + taps_bytes + 2 * pre_bytes, pre0.data(), pre_bytes);
  // one upload: [metas | taps | pre1 | pre2 | pre0]
  MG_CUDA(ctx, ctx->smp_meta.ensure(meta_bytes + taps_bytes + 3 * pre_bytes + 64));
  MG_CUDA(ctx, ctx->smp_scratch.ensure((size_t)scratch * 4 + 16));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_meta.p, hb, meta_bytes + taps_bytes + 3 * pre_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const PatchMeta* dm = ctx->smp_meta.as<PatchMeta>();
  const float* dtaps = reinterpret_cast<const float*>(ctx->smp_meta.as<uint8_t>() + meta_bytes);
  const int* dpre1 = reinterpret_cast<const int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes + taps_bytes);
  const int* dpre2 = reinterpret_cast<const int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes + taps_bytes + pre_bytes);
  const int* dpre0 = reinterpret_cast<const int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes + taps_bytes + 2 * pre_bytes);
  static OnceFlags attr_set;
  if (attr_set.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_large_rowpass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (2 * MAX_PS + 642 + L2_ROWS * ((MAX_R + 602) | 1)) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_a<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, a_smem_floats(A2_R, 30) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_a<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, a_smem_floats(A2_R, 30) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_b<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b_smem_floats(B1_R, 30, 4) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_b<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b_smem_floats(B2_R, 30, 8) * 4));
    attr_set.set(ctx->device);
  }
  // algorithmic bytes: R*R*4 read + ps*ps written per region
  auto alg_bytes = [&](const std::vector<PatchMeta>& v) {
    double b = 0;
    for (const PatchMeta& m : v) b += (double)m.R * m.R * 4.0 + (double)ps * ps;
    return b;
  };
  if (!cls[C_SMALL].empty()) {
    MG_PROF(ctx, "k_sample_small", 0, alg_bytes(cls[C_SMALL]));
    k_sample_small<<<(unsigned)cls[C_SMALL].size(), 128, SMEM_SMALL, ctx->stream>>>(img->d, img->w, img->h, dm, dtaps, d_out, ps, d_outf, nullptr);
    MG_LAUNCHED(ctx);
  }
  const int a_maxR[2] = {A1_R, A2_R}, b_maxR[2] = {B1_R, B2_R};
  for (int c = C_A1; c <= C_A2; c++) {
    if (cls[c].empty()) continue;
    const int smem = a_smem_floats(std::min(a_maxR[c - C_A1], cls[c][0].R), cls_r[c]) * 4;
    MG_PROF(ctx, c == C_A1 ? "k_sample_a<R<=40>" : "k_sample_a<R<=65>", 0, alg_bytes(cls[c]));
    launch_sample_a(c == C_A1, (unsigned)cls[c].size(), smem, ctx->stream, (const float*)img->d, img->w, img->h, dm + cls_off[c], dtaps, d_out, ps, d_outf, (const int*)nullptr);
    MG_LAUNCHED(ctx);
  }
  for (int c = C_B1; c <= C_B2; c++) {
    if (cls[c].empty()) continue;
    const int smem = b_smem_floats(std::min(b_maxR[c - C_B1], cls[c][0].R), cls_r[c], c == C_B1 ? smp_b1_ts() : 8) * 4;
    MG_PROF(ctx, c == C_B1 ? "k_sample_b<R<=100>" : "k_sample_b<R<=160>", 0, alg_bytes(cls[c]));
    launch_sample_b(c == C_B1, (unsigned)cls[c].size(), smem, ctx->stream, (const float*)img->d, img->w, img->h, dm + cls_off[c], dtaps, d_out, (const int*)nullptr);
    MG_LAUNCHED(ctx);
  }
  if (nl > 0) {
    const PatchMeta* dl = dm + cls_off[C_LARGE];
    float* scr = ctx->smp_scratch.as<float>();
    MG_PROF(ctx, "k_large_starts", 2, (double)nl);
    k_large_starts<<<pre0[nl], L0_ROWS, 0, ctx->stream>>>(dl, nl, dpre0, scr, nullptr, img->w, img->h);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_resample", 0, alg_bytes(large));
    k_large_resample<<<pre1[nl], 128, 0, ctx->stream>>>(img->d, img->w, img->h, dl, nl, dpre1, scr, nullptr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_rowpass", 2, (double)nl);
    k_large_rowpass<<<pre2[nl], 256, (2 * MAX_PS + 642 + L2_ROWS * maxPS) * 4, ctx->stream>>>(dl, nl, dpre2, dtaps, scr, ps, nullptr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_colpass_final", 2, (double)nl);
    k_large_colpass_final<<<nl * ceil_div(ps, L3_OUT_ROWS), 256, SMEM_L3, ctx->stream>>>(dl, dtaps, scr, d_out, ps, d_outf, nullptr);
    MG_LAUNCHED(ctx);
  }
  return 0;
}

// =====================================================================================================================
// Device-side preparation (the per-view chain, chain.cu): the region list lives on the device and its length is only
// known there, so classification, the decreasing-R placement inside a class and the work lists of the large-window
// path are built on the device (k_smp_classify / k_smp_scan / k_smp_scatter) instead of the host loop of mg_sample_enqueue.  Launch grids are sized by
// the upper bounds of SmpStats (k_smp_stats over ALL keypoints of the view; later passes see subsets with the same
// scales); surplus CTAs leave at once.  patchSize 32, u8 output only.  The arithmetic that decides R, the class and
// the taps is the host path's, term by term.
// =====================================================================================================================
namespace {

enum { SC_SMALL = 0, SC_A1, SC_A2, SC_B1, SC_B2, SC_LARGE, SC_N };
constexpr int DEV_PS = 32;
constexpr int HB_SMALL = 162;                       // R bins of the five shared-memory classes (R <= 160)
constexpr int HB_TOTAL = 5 * HB_SMALL + MAX_R + 1;  // + the large class

struct TapTab { const int* off; const int* ks; const float* taps; };

// R, class and taps of one region (mg_sample_enqueue's loop body, ps = 32, u8 output)
__device__ __forceinline__ void classify_region(const modsgpu_region& k, double mrSize, const TapTab& tt, PatchMeta& m, int& c) {
  m.x = (float)k.x; m.y = (float)k.y;
  m.a11 = (float)k.a11; m.a12 = (float)k.a12; m.a21 = (float)k.a21; m.a22 = (float)k.a22;
  const float mrScale = (float)ceil(k.s * mrSize);
  int R0 = 2 * int(mrScale);
  if (R0 > MAX_R - 2) R0 = MAX_R - 2;      // the host refuses such views after the statistics read-back (maxR)
  if (R0 < 0) R0 = 0;
  m.scale = float(R0) / float(DEV_PS);
  m.scratch_off = 0; m.ks = 0; m.tap_off = 0; m.out_index = 0;
  c = SC_SMALL;
  if ((double)m.scale > 0.4) {
    m.R = R0 + 2;
    m.tap_off = tt.off[R0]; m.ks = tt.ks[R0];
    const bool blocked = m.ks >= 7 && m.ks <= 60;
    if (blocked && m.R <= A1_R) c = SC_A1;
    else if (blocked && m.R <= A2_R) c = SC_A2;
    else if (blocked && R0 >= 2 * DEV_PS && m.R <= B1_R) c = SC_B1;
    else if (blocked && R0 >= 2 * DEV_PS && m.R <= B2_R) c = SC_B2;
    else if (m.R <= SMALL_R && m.ks <= 31) c = SC_SMALL;
    else c = SC_LARGE;
  } else {
    m.R = 0;
  }
}
__device__ __forceinline__ int hist_bin(int c, int R) { return c < SC_LARGE ? c * HB_SMALL + R : 5 * HB_SMALL + R; }
__host__ __device__ inline long long large_region_floats(int R, int ps) {
  long long f = large_c_floats(R) + (long long)R * R + (long long)R * 2 * ps;
  return f + (f & 1);          // the float2 table at the head of the next region stays 8-byte aligned
}

// statistics of a region list (one CTA): per-class counts / largest window / widest blur, and the block and scratch
// totals of the large-window path
__global__ void __launch_bounds__(1024)
k_smp_stats(const DevRegion* __restrict__ regs, const int* __restrict__ cnt, double mrSize, TapTab tt, SmpStats* __restrict__ out) {
  __shared__ int s_cnt[SC_N], s_rmax[SC_N], s_kr[SC_N], s_i[8];
  __shared__ unsigned long long s_scratch;
  if (threadIdx.x < SC_N) { s_cnt[threadIdx.x] = 0; s_rmax[threadIdx.x] = 0; s_kr[threadIdx.x] = 0; }
  if (threadIdx.x < 8) s_i[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_scratch = 0ull;
  __syncthreads();
  const int n = *cnt;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    PatchMeta m; int c;
    classify_region(regs[i].det, mrSize, tt, m, c);
    atomicAdd(&s_cnt[c], 1);
    atomicMax(&s_rmax[c], m.R);
    atomicMax(&s_kr[c], m.ks >> 1);
    atomicMax(&s_i[5], 2 * int((float)ceil(regs[i].det.s * mrSize)) + 2);     // unclamped R
    atomicMax(&s_i[6], m.ks);
    if (c == SC_LARGE) {
      atomicAdd(&s_i[0], 1);
      atomicAdd(&s_i[1], (m.R + L0_ROWS - 1) / L0_ROWS);
      atomicAdd(&s_i[2], (m.R + L1_ROWS - 1) / L1_ROWS);
      atomicAdd(&s_i[3], (m.R + L2_ROWS - 1) / L2_ROWS);
      atomicMax(&s_i[4], (m.R + 2 * (m.ks >> 1) + 2) | 1);
      atomicAdd(&s_scratch, (unsigned long long)large_region_floats(m.R, DEV_PS));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < SC_N; c++) { out->cls_cnt[c] = s_cnt[c]; out->cls_rmax[c] = s_rmax[c]; out->cls_kr[c] = s_kr[c]; }
    out->nl = s_i[0]; out->pre0 = s_i[1]; out->pre1 = s_i[2]; out->pre2 = s_i[3]; out->maxPS = s_i[4];
    out->maxR = s_i[5]; out->maxks = s_i[6]; out->_pad = 0;
    out->scratch = (long long)s_scratch;
  }
}

struct PrepLayout { int cls_off[SC_N]; int rtop[SC_N]; int nl_cap; };   // class c's PatchMeta live at metas + cls_off[c] (upper-bound layout); rtop: its largest window

// The sampler's work lists for the regions [0, *cnt) in three small launches (a first version did everything in ONE
// 1024-thread CTA: 73 us, all of it the 128-byte-stride row reads and scattered 48-byte stores of 4.4k regions through a
// single SM's load/store path -- 27k warp instructions in 140k cycles):
//   k_smp_classify  grid-wide, one thread per region: PatchMeta + (class, R) bin, rank inside the bin by a global atomic
//   k_smp_scan      one CTA: bin counts -> start positions (class slabs in order of decreasing R), live class counts, and
//                   for the large-window class the per-bin bases of the scratch offsets / block-count prefixes (equal-R
//                   regions have equal sizes, so a region's values follow from its bin base and its rank)
//   k_smp_scatter   grid-wide: PatchMeta to its slot; large-window regions also get scratch_off and their pre0/1/2 entries
// The order inside a class only decides which CTAs start first; equal-R regions may land in any order.
struct PrepBins {                 // device scratch of the three kernels (ints unless noted)
  int hist[HB_TOTAL];             // counts (k_smp_classify), zero again after k_smp_scan
  int start[HB_TOTAL];            // slot of the first region of every bin
  int lb0[MAX_R + 1], lb1[MAX_R + 1], lb2[MAX_R + 1];   // large class: block-count prefixes at the first region of bin R
  long long lscr[MAX_R + 1];      //              scratch offset (floats) of the first region of bin R
  int lpos[MAX_R + 1];            //              index inside the large slab of the first region of bin R
};

__global__ void __launch_bounds__(256)
k_smp_classify(const DevRegion* __restrict__ regs, const int* __restrict__ cnt, double mrSize, TapTab tt, PrepBins* __restrict__ B,
               PatchMeta* __restrict__ mtmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *cnt) return;
  PatchMeta m; int c;
  classify_region(regs[i].det, mrSize, tt, m, c);
  const int bin = hist_bin(c, m.R);
  m.out_index = i;
  m.scratch_off = (long long)bin << 32 | (unsigned)atomicAdd(&B->hist[bin], 1);     // (bin, rank) ride in the unused field
  mtmp[i] = m;
}

__global__ void __launch_bounds__(256)
k_smp_scan(PrepBins* __restrict__ B, PrepLayout lay, int* __restrict__ pre1, int* __restrict__ pre2, int* __restrict__ pre0,
           int* __restrict__ cls_cnt_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= SC_N) return;
  // Warp c turns class c's bin counts into start positions, walking R downwards from the largest window the view can hold
  // (lay.rtop[c], from the statistics pass).  Lane l owns a STRIP of consecutive bins: it adds its strip up serially, one
  // warp scan combines the 32 strip totals, then it walks its strip again writing the bases.  (A version that scanned 32
  // bins per step with shuffles took 63 us: 65 steps x ~150 dependent instructions on one warp.)
  const int c = warp, rtop = lay.rtop[c];
  const int L = (rtop + 32) / 32;                       // bins per lane; lane l owns R = rtop - l*L ... rtop - l*L - (L-1)
  const int rhi = rtop - lane * L;
  int n = 0, t0 = 0, t1 = 0, t2 = 0;
  long long ts = 0;
  for (int k = 0; k < L; k++) {
    const int R = rhi - k;
    if (R < 0) break;
    const int v = B->hist[hist_bin(c, R)];
    n += v;
    if (c == SC_LARGE && v) {
      t0 += v * ((R + L0_ROWS - 1) / L0_ROWS); t1 += v * ((R + L1_ROWS - 1) / L1_ROWS); t2 += v * ((R + L2_ROWS - 1) / L2_ROWS);
      ts += (long long)v * large_region_floats(R, DEV_PS);
    }
  }
  int ni = n, i0 = t0, i1 = t1, i2 = t2;
  long long si = ts;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int un = __shfl_up_sync(0xffffffffu, ni, o), u0 = __shfl_up_sync(0xffffffffu, i0, o), u1 = __shfl_up_sync(0xffffffffu, i1, o),
              u2 = __shfl_up_sync(0xffffffffu, i2, o);
    const long long us = __shfl_up_sync(0xffffffffu, si, o);
    if (lane >= o) { ni += un; i0 += u0; i1 += u1; i2 += u2; si += us; }
  }
  const int total = __shfl_sync(0xffffffffu, ni, 31), tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31),
            tot2 = __shfl_sync(0xffffffffu, i2, 31);
  int pos = ni - n, p0 = i0 - t0, p1 = i1 - t1, p2 = i2 - t2;      // exclusive prefixes at the head of this lane's strip
  long long ps = si - ts;
  for (int k = 0; k < L; k++) {
    const int R = rhi - k;
    if (R < 0) break;
    const int bin = hist_bin(c, R);
    const int v = B->hist[bin];
    B->start[bin] = lay.cls_off[c] + pos;
    B->hist[bin] = 0;
    if (c == SC_LARGE) {
      B->lb0[R] = p0; B->lb1[R] = p1; B->lb2[R] = p2; B->lscr[R] = ps; B->lpos[R] = pos;
      if (v) {
        p0 += v * ((R + L0_ROWS - 1) / L0_ROWS); p1 += v * ((R + L1_ROWS - 1) / L1_ROWS); p2 += v * ((R + L2_ROWS - 1) / L2_ROWS);
        ps += (long long)v * large_region_floats(R, DEV_PS);
      }
    }
    pos += v;
  }
  if (lane == 0) {
    cls_cnt_out[c] = total;
    if (c == SC_LARGE) {
      const int nl = min(total, lay.nl_cap);
      pre0[nl] = tot0; pre1[nl] = tot1; pre2[nl] = tot2;
    }
  }
}

__global__ void __launch_bounds__(256)
k_smp_scatter(const PatchMeta* __restrict__ mtmp, const int* __restrict__ cnt, const PrepBins* __restrict__ B, PrepLayout lay,
              PatchMeta* __restrict__ metas, int* __restrict__ pre1, int* __restrict__ pre2, int* __restrict__ pre0,
              double* __restrict__ prof_bytes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = i < *cnt;
  double pb[SC_N] = {0, 0, 0, 0, 0, 0};
  if (live) {
    PatchMeta m = mtmp[i];
    const int bin = (int)(m.scratch_off >> 32), rank = (int)(m.scratch_off & 0xffffffffll);
    m.scratch_off = 0;
    const int c = bin < 5 * HB_SMALL ? bin / HB_SMALL : SC_LARGE;
    if (c == SC_LARGE) {
      const int R = m.R;
      const int li = B->lpos[R] + rank;                      // index inside the large slab
      m.scratch_off = B->lscr[R] + (long long)rank * large_region_floats(R, DEV_PS);
      if (li < lay.nl_cap) {
        pre0[li] = B->lb0[R] + rank * ((R + L0_ROWS - 1) / L0_ROWS);
        pre1[li] = B->lb1[R] + rank * ((R + L1_ROWS - 1) / L1_ROWS);
        pre2[li] = B->lb2[R] + rank * ((R + L2_ROWS - 1) / L2_ROWS);
      }
    }
    metas[B->start[bin] + rank] = m;
    // profiler only: the algorithmic bytes of this region (R*R*4 read + 32*32 written), per class
    if (prof_bytes != nullptr) pb[c] = (double)m.R * m.R * 4.0 + (double)(DEV_PS * DEV_PS);
  }
  if (prof_bytes != nullptr) {       // one atomic per warp and class
#pragma unroll
    for (int k = 0; k < SC_N; k++) {
      double v = pb[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v > 0) atomicAdd(prof_bytes + k, v);
    }
  }
}

}  // namespace

// The taps of every even window size R0 in [2, MAX_R - 2] for patchSize 32 (sigma = 1.5 * R0 / 32, helpers.cpp:726-731),
// built once per context: [off: MAX_R ints | ks: MAX_R ints | taps].  ~1.2 MB.
static int smp_taptab(modsgpu_ctx* ctx, TapTab& tt) {
  if (!ctx->smp_taptab.p) {
    std::vector<int> off(MAX_R, 0), ks(MAX_R, 0);
    std::vector<float> taps, t;
    for (int R0 = 2; R0 <= MAX_R - 2; R0 += 2) {
      const float scale = float(R0) / float(DEV_PS);
      if (!((double)scale > 0.4)) continue;
      ks[R0] = mg_gaussian_taps(1.5f * scale, t);
      off[R0] = (int)taps.size();
      taps.insert(taps.end(), t.begin(), t.end());
    }
    const size_t bytes = (size_t)2 * MAX_R * 4 + taps.size() * 4;
    MG_CUDA(ctx, ctx->smp_taptab.ensure(bytes));
    std::vector<uint8_t> blob(bytes);
    memcpy(blob.data(), off.data(), MAX_R * 4);
    memcpy(blob.data() + MAX_R * 4, ks.data(), MAX_R * 4);
    memcpy(blob.data() + 2 * MAX_R * 4, taps.data(), taps.size() * 4);
    MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_taptab.p, blob.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // `blob` is pageable and dies here
  }
  tt.off = ctx->smp_taptab.as<int>();
  tt.ks = tt.off + MAX_R;
  tt.taps = reinterpret_cast<const float*>(tt.ks + MAX_R);
  return 0;
}

// statistics of the regions [0, *cnt_dev) into *d_stats (device); no host synchronisation
int mg_sample_stats_enqueue(modsgpu_ctx* ctx, const DevRegion* regs, const int* cnt_dev, double mrSize, SmpStats* d_stats) {
  TapTab tt;
  if (int rc = smp_taptab(ctx, tt)) return rc;
  MG_PROF(ctx, "k_smp_stats", 2, 1.0);
  k_smp_stats<<<1, 1024, 0, ctx->stream>>>(regs, cnt_dev, mrSize, tt, d_stats);
  MG_LAUNCHED(ctx);
  return 0;
}

// The sampler over a DEVICE region list of *cnt_dev (<= n_ub) regions, 32x32 u8 patches in region order into d_out.
// `st` = the view's upper bounds (mg_sample_stats_enqueue over a superset of these regions).  No host synchronisation.
int mg_sample_enqueue_dev(modsgpu_ctx* ctx, const modsgpu_image* img, const DevRegion* regs, const int* cnt_dev, int n_ub,
                          const SmpStats& st, double mrSize, uint8_t* d_out) {
  if (n_ub <= 0) return 0;
  if (st.maxR > MAX_R) MG_FAIL(ctx, MODSGPU_EINVAL, "region too large for the sampler (R > 2048)");
  if (st.maxks > 600) MG_FAIL(ctx, MODSGPU_EINVAL, "sampler blur too wide");
  TapTab tt;
  if (int rc = smp_taptab(ctx, tt)) return rc;
  const int ps = DEV_PS;
  PrepLayout lay;
  size_t nm = 0;
  for (int c = 0; c < SC_N; c++) {
    lay.cls_off[c] = (int)nm; nm += (size_t)st.cls_cnt[c];
    lay.rtop[c] = std::min(c < SC_LARGE ? HB_SMALL - 1 : MAX_R, std::max(st.cls_rmax[c], 0));     // no region of a subset has a larger window
  }
  lay.nl_cap = st.nl;
  const int nl = st.nl;
  const size_t meta_bytes = (nm + 1) * sizeof(PatchMeta), pre_bytes = (size_t)(nl + 1) * 4;
  MG_CUDA(ctx, ctx->smp_meta.ensure(meta_bytes + 3 * pre_bytes + 64));
  MG_CUDA(ctx, ctx->smp_regs.ensure((size_t)(n_ub + 1) * sizeof(PatchMeta) + 16));
  MG_CUDA(ctx, ctx->smp_scratch.ensure((size_t)st.scratch * 4 + 16));
  PatchMeta* dm = ctx->smp_meta.as<PatchMeta>();
  int* dpre1 = reinterpret_cast<int*>(ctx->smp_meta.as<uint8_t>() + meta_bytes);
  int* dpre2 = dpre1 + (nl + 1);
  int* dpre0 = dpre2 + (nl + 1);
  int* dcnt = dpre0 + (nl + 1);
  double* prof_bytes = nullptr;      // device accumulators read by modsgpu_profile_report (api.cu)
  if (ctx->prof.on) {
    if (!ctx->smp_prof.p) { MG_CUDA(ctx, ctx->smp_prof.ensure(64)); MG_CUDA(ctx, cudaMemsetAsync(ctx->smp_prof.p, 0, 64, ctx->stream)); }
    prof_bytes = ctx->smp_prof.as<double>();
  }
  if (!ctx->smp_bins.p) {            // histogram scratch: zero once, k_smp_scan leaves it zero
    MG_CUDA(ctx, ctx->smp_bins.ensure(sizeof(PrepBins)));
    MG_CUDA(ctx, cudaMemsetAsync(ctx->smp_bins.p, 0, sizeof(PrepBins), ctx->stream));
  }
  PrepBins* bins = ctx->smp_bins.as<PrepBins>();
  PatchMeta* mtmp = ctx->smp_regs.as<PatchMeta>();
  const int nblk = ceil_div(n_ub, 256);
  MG_PROF(ctx, "k_smp_classify", 2, (double)n_ub);
  k_smp_classify<<<nblk, 256, 0, ctx->stream>>>(regs, cnt_dev, mrSize, tt, bins, mtmp);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_smp_scan", 2, (double)HB_TOTAL);
  k_smp_scan<<<1, 256, 0, ctx->stream>>>(bins, lay, dpre1, dpre2, dpre0, dcnt);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_smp_scatter", 2, (double)n_ub);
  k_smp_scatter<<<nblk, 256, 0, ctx->stream>>>(mtmp, cnt_dev, bins, lay, dm, dpre1, dpre2, dpre0, prof_bytes);
  MG_LAUNCHED(ctx);
  static OnceFlags attr_set;
  if (attr_set.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SMALL));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_large_rowpass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (2 * MAX_PS + 642 + L2_ROWS * ((MAX_R + 602) | 1)) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_a<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, a_smem_floats(A2_R, 30) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_a<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, a_smem_floats(A2_R, 30) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_b<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b_smem_floats(B1_R, 30, 4) * 4));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_sample_b<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b_smem_floats(B2_R, 30, 8) * 4));
    attr_set.set(ctx->device);
  }
  // the algorithmic bytes of these launches are only known on the device: k_smp_scatter accumulates them per class and
  // modsgpu_profile_report adds them to the kernels' records
  auto alg_bytes = [&](int) { return 0.0; };
  const float* dtaps = tt.taps;
  if (st.cls_cnt[SC_SMALL] > 0) {
    MG_PROF(ctx, "k_sample_small", 0, alg_bytes(SC_SMALL));
    k_sample_small<<<(unsigned)st.cls_cnt[SC_SMALL], 128, SMEM_SMALL, ctx->stream>>>(img->d, img->w, img->h, dm + lay.cls_off[SC_SMALL], dtaps,
                                                                                    d_out, ps, nullptr, dcnt + SC_SMALL);
    MG_LAUNCHED(ctx);
  }
  const int a_maxR[2] = {A1_R, A2_R}, b_maxR[2] = {B1_R, B2_R};
  for (int c = SC_A1; c <= SC_A2; c++) {
    if (st.cls_cnt[c] <= 0) continue;
    const int smem = a_smem_floats(std::min(a_maxR[c - SC_A1], st.cls_rmax[c]), st.cls_kr[c]) * 4;
    MG_PROF(ctx, c == SC_A1 ? "k_sample_a<R<=40>" : "k_sample_a<R<=65>", 0, alg_bytes(c));
    launch_sample_a(c == SC_A1, (unsigned)st.cls_cnt[c], smem, ctx->stream, (const float*)img->d, img->w, img->h, (const PatchMeta*)(dm + lay.cls_off[c]), dtaps, d_out, ps, (float*)nullptr, (const int*)(dcnt + c));
    MG_LAUNCHED(ctx);
  }
  for (int c = SC_B1; c <= SC_B2; c++) {
    if (st.cls_cnt[c] <= 0) continue;
    const int smem = b_smem_floats(std::min(b_maxR[c - SC_B1], st.cls_rmax[c]), st.cls_kr[c], c == SC_B1 ? smp_b1_ts() : 8) * 4;
    MG_PROF(ctx, c == SC_B1 ? "k_sample_b<R<=100>" : "k_sample_b<R<=160>", 0, alg_bytes(c));
    launch_sample_b(c == SC_B1, (unsigned)st.cls_cnt[c], smem, ctx->stream, (const float*)img->d, img->w, img->h, (const PatchMeta*)(dm + lay.cls_off[c]), dtaps, d_out, (const int*)(dcnt + c));
    MG_LAUNCHED(ctx);
  }
  if (nl > 0) {
    const PatchMeta* dl = dm + lay.cls_off[SC_LARGE];
    float* scr = ctx->smp_scratch.as<float>();
    const int* dnl = dcnt + SC_LARGE;
    MG_PROF(ctx, "k_large_starts", 2, (double)nl);
    k_large_starts<<<st.pre0, L0_ROWS, 0, ctx->stream>>>(dl, nl, dpre0, scr, dnl, img->w, img->h);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_resample", 0, alg_bytes(SC_LARGE));
    k_large_resample<<<st.pre1, 128, 0, ctx->stream>>>(img->d, img->w, img->h, dl, nl, dpre1, scr, dnl);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_rowpass", 2, (double)nl);
    k_large_rowpass<<<st.pre2, 256, (2 * MAX_PS + 642 + L2_ROWS * st.maxPS) * 4, ctx->stream>>>(dl, nl, dpre2, dtaps, scr, ps, dnl);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_large_colpass_final", 2, (double)nl);
    k_large_colpass_final<<<nl * ceil_div(ps, L3_OUT_ROWS), 256, SMEM_L3, ctx->stream>>>(dl, dtaps, scr, d_out, ps, nullptr, dnl);
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
  int rc = mg_sample_enqueue(ctx, img, regs, n, mrSize, patchSize, ctx->smp_out.as<uint8_t>(), nullptr);
  if (rc) return rc;
  if (n > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->smp_out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

// float patches (no u8 quantisation): what DescribeRegions (synth-detection.hpp:170-263) hands to the SIFT descriptor
extern "C" int modsgpu_extract_patches_f32(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                           double mrSize, int patchSize, float* out) {
  if (!ctx || !img || (n > 0 && (!regs || !out)) || n < 0) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  size_t bytes = (size_t)n * patchSize * patchSize * 4;
  MG_CUDA(ctx, ctx->smp_regs.ensure(bytes + 16));
  int rc = mg_sample_enqueue(ctx, img, regs, n, mrSize, patchSize, nullptr, ctx->smp_regs.as<float>());
  if (rc) return rc;
  if (n > 0) MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->smp_regs.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}
