// detect.cu -- Hessian scale-space detector on the device.
//
// Compiled with --fmad=false: every float expression below is evaluated exactly as written
// (no contraction); fused multiply-adds appear only as explicit fmaf().  That is what makes the
// keypoint list bit-identical to the CPU oracle (oracle/mods_oracle.cpp), which restates
// pyramid.cpp:196-529 and the arithmetic order of cv::GaussianBlur / cv::resize.
//
// Launch structure per image (nS = 3): per octave three dependent blurs on the main stream (k_blur3<ks>: blur + response of
// the blurred tile, the first one also the response of its source, the third one also the half-size base of the next
// octave), the fourth blur and the octave's NMS + localisation on a side stream; k_resolve / k_rank_export at the end.
// The sequence is captured into one CUDA graph per image buffer (mg_detect_graph).
//
// Data layout in HBM: one workspace holds, per octave o (w_o x h_o, dense row-major fp32),
// five blur levels L[0..4] and five responses R[0..4]; an int32 "octave map" (w_o x h_o)
// resolves the reference's sequential octaveMap de-duplication (pyramid.cpp:387-391)
// deterministically: the candidate with the smallest (level, r0, c0) visiting key wins.
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace {

constexpr int BT_X = 64, BT_Y = 32, BT_THREADS = 256;
constexpr int MAX_KS = 63;

struct Taps {
  float k[MAX_KS + 1];
  int ks;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Row pass of cv::GaussianBlur for one output pixel; p points at the tap-0 sample (x - r).
// `vec` = x < (w & ~3): the SIMD body of OpenCV's row filter (fused), else its scalar tail.
__device__ __forceinline__ float row_pass(const float* p, const float* k, int ks, bool vec, bool vec2) {
  const int r = ks >> 1;
  float s;
  if (ks == 5) {
    float p1 = p[r + 1] + p[r - 1], p2 = p[r + 2] + p[r - 2], x0 = p[r];
    if (vec2) {
      s = p1 * k[3];
      s = fmaf(x0, k[2], s);
      s = fmaf(p2, k[4], s);
    } else {
      s = x0 * k[2] + p1 * k[3];
      s = s + p2 * k[4];
    }
  } else if (ks < 5) {
    s = p[r] * k[r];
    for (int t = 1; t <= r; t++) s = s + (p[r + t] + p[r - t]) * k[r + t];
  } else if (vec) {
    s = 0.f;
    for (int t = 0; t < ks; t++) s = fmaf(p[t], k[t], s);
  } else {
    const int nf = (ks - 1) % 4;
    s = p[0] * k[0];
    for (int t = 1; t < ks; t++) {
      if (t >= ks - nf) s = fmaf(p[t], k[t], s);
      else s = s + p[t] * k[t];
    }
  }
  return s;
}

// Column pass; p points at the centre sample, `pitch` floats between rows. `vec` = x < (w & ~7).
__device__ __forceinline__ float col_pass(const float* p, int pitch, const float* k, int ks, bool vec) {
  const int r = ks >> 1;
  float s = p[0] * k[r];
  if (vec) {
    for (int t = 1; t <= r; t++) s = fmaf(p[-t * pitch] + p[t * pitch], k[r + t], s);
  } else {
    for (int t = 1; t <= r; t++) s = s + (p[-t * pitch] + p[t * pitch]) * k[r + t];
  }
  return s;
}

// pyramid.cpp:196-254
__device__ __forceinline__ float hessian_at(const float* c, int pitch, float norm2) {
  float v11 = c[-pitch - 1], v12 = c[-pitch], v13 = c[-pitch + 1];
  float v21 = c[-1], v22 = c[0], v23 = c[1];
  float v31 = c[pitch - 1], v32 = c[pitch], v33 = c[pitch + 1];
  float Lxx = (v21 - 2 * v22 + v23);
  float Lyy = (v12 - 2 * v22 + v32);
  float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
  return (Lxx * Lyy - Lxy * Lxy) * norm2;
}

// Fused separable Gaussian blur (+ optional Hessian response of the blurred image).
// One CTA -> a BT_X x BT_Y output tile.  Shared memory: A = source tile with (r+1) halo,
// B = row-filtered, C = blurred tile with 1-px halo (for the 3x3 Hessian stencil).
__global__ void __launch_bounds__(BT_THREADS)
k_blur_resp(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ resp,
            int w, int h, Taps taps, float norm2) {
  extern __shared__ float sm[];
  const int ks = taps.ks, r = ks >> 1;
  const int AW = BT_X + 2 + 2 * r, AH = BT_Y + 2 + 2 * r;
  const int BW = BT_X + 2, CW = BT_X + 2, CH = BT_Y + 2;
  float* A = sm;
  float* B = A + AW * AH;
  float* C = B + BW * AH;
  __shared__ float k[MAX_KS + 1];
  const int tid = threadIdx.x;
  if (tid < ks) k[tid] = taps.k[tid];
  const int x0 = blockIdx.x * BT_X, y0 = blockIdx.y * BT_Y;
  for (int i = tid; i < AW * AH; i += BT_THREADS) {
    int ly = i / AW, lx = i - ly * AW;
    int gy = clampi(y0 - 1 - r + ly, 0, h - 1), gx = clampi(x0 - 1 - r + lx, 0, w - 1);
    A[i] = src[(size_t)gy * w + gx];
  }
  __syncthreads();
  const int wv = w & ~3, wv2 = w & ~1, wc = w & ~7;
  for (int i = tid; i < BW * AH; i += BT_THREADS) {
    int ly = i / BW, lx = i - ly * BW;
    int gx = x0 - 1 + lx;
    float v = 0.f;
    if (gx >= 0 && gx < w) v = row_pass(A + ly * AW + lx, k, ks, gx < wv, gx < wv2);
    B[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < CW * CH; i += BT_THREADS) {
    int ly = i / CW, lx = i - ly * CW;
    int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
    float v = 0.f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      v = col_pass(B + (ly + r) * BW + lx, BW, k, ks, gx < wc);
      if (lx >= 1 && lx <= BT_X && ly >= 1 && ly <= BT_Y) dst[(size_t)gy * w + gx] = v;
    }
    C[i] = v;
  }
  if (!resp) return;
  __syncthreads();
  for (int i = tid; i < BT_X * BT_Y; i += BT_THREADS) {
    int ly = i / BT_X, lx = i - ly * BT_X;
    int gx = x0 + lx, gy = y0 + ly;
    if (gx >= w || gy >= h) continue;
    float v = 0.f;
    if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) v = hessian_at(C + (ly + 1) * CW + lx + 1, CW, norm2);
    resp[(size_t)gy * w + gx] = v;
  }
}

// ---- register-blocked version (ks >= 7, i.e. every blur of the default pyramid) ----------------------------
// Same tile and the same arithmetic, organised so that the FMA chains dominate the instruction stream:
//   load   one warp per tile row, clamped coordinates (replicate border) -> A (odd pitch)
//   row    a thread owns 4 adjacent columns x 4 rows (rows strided by the group count): one tap load and one
//          sample load per row and step feed 4 FMAs; taps rotate through registers; zero taps appended to the
//          tap array keep the chain s = fma(p[t], k[t], s) bit-exact (fma(d, 0, s) == s).  Columns in OpenCV's
//          scalar tail (x >= w & ~3) and blocks straddling it take the scalar form.
//   column a thread owns 4 vertically adjacent outputs of one column: the symmetric pairs T[y-t] + T[y+t] slide
//          through registers (2 new loads per step for 4 outputs)
//   hessian from the blurred tile with its 1-px halo, as before.
constexpr int BT_NT = 512;   // 2 x 4 row blocks / 5-row column blocks: ~4 output pixels per thread, short critical path

template <int CB, int RB>
__device__ __forceinline__ void blur_rowpass_block(const float* (&rows)[RB], const float* __restrict__ kp, int ks,
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

template <int RB>
__device__ __forceinline__ void blur_colpass_block(const float* __restrict__ Tc, int pitch, const float* __restrict__ k, int r,
                                                   bool vec, float (&out)[RB]) {
  float dn[RB], up[RB];
#pragma unroll
  for (int i = 0; i < RB; i++) { dn[i] = up[i] = Tc[i * pitch]; out[i] = dn[i] * k[r]; }
#pragma unroll 4
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

__host__ __device__ inline int blur2_pitch_a(int r) { return (BT_X + 2 + 2 * r + 3) | 1; }
constexpr int BLUR2_PB = (BT_X + 2 + 3) | 1;   // 69: row-filtered tile, 66 columns (+ slack for the last 4-block)

__global__ void __launch_bounds__(BT_NT)
k_blur_resp2(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ resp,
             int w, int h, Taps taps, float norm2) {
  extern __shared__ float sm[];
  const int ks = taps.ks, r = ks >> 1;
  const int AW = BT_X + 2 + 2 * r, AH = BT_Y + 2 + 2 * r, PA = blur2_pitch_a(r);
  constexpr int BW = BT_X + 2, CW = BT_X + 2, CH = BT_Y + 2, PB = BLUR2_PB;
  float* A = sm;                       // AH x PA
  float* B = A + AH * PA;              // (AH + 3) x PB   (slack rows for the last column block)
  float* Cc = B + (AH + 3) * PB;       // CH x CW
  __shared__ float k[MAX_KS + 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < ks + 3) k[tid] = tid < ks ? taps.k[tid] : 0.f;
  const int x0 = blockIdx.x * BT_X, y0 = blockIdx.y * BT_Y;
  for (int ly = warp; ly < AH; ly += BT_NT / 32) {
    const float* srow = src + (size_t)clampi(y0 - 1 - r + ly, 0, h - 1) * w;
    float* arow = A + ly * PA;
    for (int lx = lane; lx < AW + 3; lx += 32) arow[lx] = srow[clampi(x0 - 1 - r + lx, 0, w - 1)];
  }
  __syncthreads();
  // row pass -> B[ly][lx], lx = 0..65 <-> gx = x0 - 1 + lx
  {
    const int wv = w & ~3;
    constexpr int NCB = (BW + 3) / 4;                 // 17 blocks of 4 columns
    constexpr int RB = 2;
    const int NRG = (AH + RB - 1) / RB;
    for (int it = tid; it < NCB * NRG; it += BT_NT) {
      const int cb = it / NRG, rg = it - cb * NRG, lx0 = cb * 4, gx0 = x0 - 1 + lx0;
      int rowi[RB];
#pragma unroll
      for (int b = 0; b < RB; b++) rowi[b] = rg + b * NRG;
      if (gx0 >= 0 && gx0 + 3 < wv) {
        const float* rows[RB];
#pragma unroll
        for (int b = 0; b < RB; b++) rows[b] = A + min(rowi[b], AH - 1) * PA + lx0;
        float acc[RB][4];
        blur_rowpass_block<4, RB>(rows, k, ks, acc);
#pragma unroll
        for (int b = 0; b < RB; b++)
          if (rowi[b] < AH) {
            float* t = B + rowi[b] * PB + lx0;
#pragma unroll
            for (int c = 0; c < 4; c++) t[c] = acc[b][c];
          }
      } else {
#pragma unroll 1
        for (int b = 0; b < RB; b++) {
          if (rowi[b] >= AH) continue;
#pragma unroll 1
          for (int c = 0; c < 4; c++) {
            const int gx = gx0 + c;
            float v = 0.f;
            if (gx >= 0 && gx < w && lx0 + c < BW) v = row_pass(A + rowi[b] * PA + lx0 + c, k, ks, gx < wv, gx < (w & ~1));
            B[rowi[b] * PB + lx0 + c] = v;
          }
        }
      }
    }
  }
  __syncthreads();
  // column pass -> C (blurred tile with 1-px halo) and dst
  {
    const int wc = w & ~7;
    constexpr int RBC = 5, NRGC = (CH + RBC - 1) / RBC;   // 7 groups of 5 rows: 462 items, one round
    for (int it = tid; it < NRGC * CW; it += BT_NT) {
      const int rg = it / CW, lx = it - rg * CW, ly0 = rg * RBC;
      const int gx = x0 - 1 + lx;
      float o[RBC];
      blur_colpass_block<RBC>(B + (ly0 + r) * PB + lx, PB, k, r, gx < wc, o);
#pragma unroll
      for (int i = 0; i < RBC; i++) {
        const int ly = ly0 + i, gy = y0 - 1 + ly;
        if (ly >= CH) continue;
        float v = 0.f;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
          v = o[i];
          if (lx >= 1 && lx <= BT_X && ly >= 1 && ly <= BT_Y) dst[(size_t)gy * w + gx] = v;
        }
        Cc[ly * CW + lx] = v;
      }
    }
  }
  if (!resp) return;
  __syncthreads();
  for (int i = tid; i < BT_X * BT_Y; i += BT_NT) {
    const int ly = i >> 6, lx = i & 63;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gx >= w || gy >= h) continue;
    float v = 0.f;
    if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) v = hessian_at(Cc + (ly + 1) * CW + lx + 1, CW, norm2);
    resp[(size_t)gy * w + gx] = v;
  }
}

int blur2_smem_bytes(int ks) {
  const int r = ks >> 1, AH = BT_Y + 2 + 2 * r;
  return (AH * blur2_pitch_a(r) + (AH + 3) * BLUR2_PB + (BT_X + 2) * (BT_Y + 2)) * (int)sizeof(float);
}

// ---- ksize-specialised version (7 <= ks <= 23: every blur of a pyramid with 2..6 scales per octave) ---------------------
// Same tile, same arithmetic order, a third of the instructions and a shorter critical path per CTA (the detector is a
// chain of 19 dependent launches per image, so a CTA's latency IS the detector's latency):
//   load   every thread issues all of its (AH x AW) / 512 clamped global loads before the first shared-memory store
//          (the old loop exposed one L2 round trip per iteration)
//   row    a thread owns CB adjacent outputs of one row and streams the CB + ks - 1 samples once; the loops are fully
//          unrolled, so the taps are constant-bank operands of the FMAs (no tap loads, no register rotation) and each
//          chain still accumulates t = 0 .. ks-1 in order.  CB is chosen so that the whole tile is ONE round of 512 threads
//   column a thread owns 5 vertically adjacent outputs of one column: 5 + 2r loads, symmetric pairs in registers
//   fused  (a) Hessian response of the blurred tile (as before); (b) optionally the response of the SOURCE tile (first
//          level of octaves >= 1: replaces k_response); (c) optionally the half-size image of the blurred tile
//          (pyramid.cpp:476, replaces k_half): two launches fewer per octave on the critical path.
template <int KS>
struct B3 {
  static constexpr int R = KS / 2;
  static constexpr int AW = BT_X + 2 + 2 * R, AH = BT_Y + 2 + 2 * R;
  static constexpr int BW = BT_X + 2, CH = BT_Y + 2;
  static constexpr int pick_cb() {
    for (int cb = 6; cb <= 12; cb++)
      if (((BW + cb - 1) / cb) * AH <= BT_NT) return cb;
    return 0;
  }
  static constexpr int CB = pick_cb();
  static constexpr int NCB = (BW + CB - 1) / CB;
  static constexpr int PA = (NCB * CB + 2 * R) | 1;     // covers the reads of the last column block
  static constexpr int PB = (NCB * CB) | 1;
  static constexpr int NL = (AH * AW + BT_NT - 1) / BT_NT;
  static constexpr int RBC = 5, NRGC = (CH + RBC - 1) / RBC;
  static constexpr int SMEM = (AH * PA + (AH + 1) * PB + CH * BW) * (int)sizeof(float);
  static_assert(CB > 0 && NRGC * BW <= BT_NT, "tile does not fit one round");
};

template <int KS>
__global__ void __launch_bounds__(BT_NT)
k_blur3(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ resp, int w, int h,
        const __grid_constant__ Taps taps, float norm2,
        float* __restrict__ resp_src, float norm2_src, float* __restrict__ half_out, int ow, int oh) {
  using G = B3<KS>;
  constexpr int R = G::R, AW = G::AW, AH = G::AH, PA = G::PA, PB = G::PB, BW = G::BW, CH = G::CH, CB = G::CB;
  extern __shared__ float sm[];
  float* A = sm;                        // AH x PA   source tile, (R + 1)-px replicated halo
  float* B = A + AH * PA;               // (AH + 1) x PB   row-filtered
  float* Cc = B + (AH + 1) * PB;        // CH x BW   blurred tile, 1-px halo (0 outside the image)
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * BT_X, y0 = blockIdx.y * BT_Y;
  {
    float v[G::NL];
#pragma unroll
    for (int k = 0; k < G::NL; k++) {
      const int i = tid + k * BT_NT;
      if (i < AH * AW) {
        const int ly = i / AW, lx = i - ly * AW;
        const int gy = clampi(y0 - 1 - R + ly, 0, h - 1), gx = clampi(x0 - 1 - R + lx, 0, w - 1);
        v[k] = src[(size_t)gy * w + gx];
      }
    }
#pragma unroll
    for (int k = 0; k < G::NL; k++) {
      const int i = tid + k * BT_NT;
      if (i < AH * AW) {
        const int ly = i / AW, lx = i - ly * AW;
        A[ly * PA + lx] = v[k];
      }
    }
  }
  __syncthreads();
  // (b) response of the source image on the inner tile
  if (resp_src != nullptr) {
#pragma unroll
    for (int q = 0; q < BT_X * BT_Y / BT_NT; q++) {
      const int i = tid + q * BT_NT, ly = i >> 6, lx = i & 63;
      const int gx = x0 + lx, gy = y0 + ly;
      if (gx < w && gy < h) {
        float v = 0.f;
        if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) v = hessian_at(A + (ly + R + 1) * PA + lx + R + 1, PA, norm2_src);
        resp_src[(size_t)gy * w + gx] = v;
      }
    }
  }
  // row pass -> B[row][lx], lx = 0..65 <-> gx = x0 - 1 + lx; output lx reads A columns lx .. lx + ks - 1
  if (tid < G::NCB * AH) {
    const int cb = tid / AH, row = tid - cb * AH, lx0 = cb * CB, gx0 = x0 - 1 + lx0;
    const float* p = A + row * PA + lx0;
    float* b = B + row * PB + lx0;
    if (min(gx0 + CB - 1, w - 1) < (w & ~3)) {        // every in-image column of the block is in the vector body
      float acc[CB];
#pragma unroll
      for (int c = 0; c < CB; c++) acc[c] = 0.f;
#pragma unroll
      for (int j = 0; j < CB + KS - 1; j++) {
        const float d = p[j];
#pragma unroll
        for (int c = 0; c < CB; c++)
          if (j - c >= 0 && j - c < KS) acc[c] = fmaf(d, taps.k[j - c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < CB; c++) b[c] = acc[c];
    } else {
#pragma unroll 1
      for (int c = 0; c < CB; c++) {
        const int gx = gx0 + c;
        float v = 0.f;
        if (gx >= 0 && gx < w) v = row_pass(p + c, taps.k, KS, gx < (w & ~3), gx < (w & ~1));
        b[c] = v;
      }
    }
  }
  __syncthreads();
  // column pass -> C and dst
  if (tid < G::NRGC * BW) {
    constexpr int RBC = G::RBC;
    const int rg = tid / BW, lx = tid - rg * BW, ly0 = rg * RBC;
    const int gx = x0 - 1 + lx;
    const float* Tc = B + ly0 * PB + lx;               // C row ly <-> B rows ly .. ly + 2R (centre ly + R)
    float T[RBC + 2 * R], o[RBC];
#pragma unroll
    for (int i = 0; i < RBC + 2 * R; i++) T[i] = Tc[i * PB];
#pragma unroll
    for (int i = 0; i < RBC; i++) o[i] = T[i + R] * taps.k[R];
    if (gx < (w & ~7)) {
#pragma unroll
      for (int t = 1; t <= R; t++)
#pragma unroll
        for (int i = 0; i < RBC; i++) o[i] = fmaf(T[i + R - t] + T[i + R + t], taps.k[R + t], o[i]);
    } else {
#pragma unroll
      for (int t = 1; t <= R; t++)
#pragma unroll
        for (int i = 0; i < RBC; i++) o[i] = o[i] + (T[i + R - t] + T[i + R + t]) * taps.k[R + t];
    }
#pragma unroll
    for (int i = 0; i < RBC; i++) {
      const int ly = ly0 + i, gy = y0 - 1 + ly;
      if (ly < CH) {
        float v = 0.f;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
          v = o[i];
          if (lx >= 1 && lx <= BT_X && ly >= 1 && ly <= BT_Y) dst[(size_t)gy * w + gx] = v;
        }
        Cc[ly * BW + lx] = v;
      }
    }
  }
  if (resp == nullptr && half_out == nullptr) return;
  __syncthreads();
  // (a) response of the blurred tile: a thread owns 4 vertically adjacent pixels of one column
  if (resp != nullptr) {
    const int lx = tid & 63, ly0 = (tid >> 6) * 4;
    const int gx = x0 + lx;
    const float* c = Cc + ly0 * BW + lx;                // rows ly0 .. ly0 + 5 of the haloed tile, columns lx .. lx + 2
    float v[6][3];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) v[i][j] = c[i * BW + j];
    if (gx < w) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gy = y0 + ly0 + i;
        if (gy >= h) break;
        float out = 0.f;
        if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) {
          const float Lxx = (v[i + 1][0] - 2 * v[i + 1][1] + v[i + 1][2]);
          const float Lyy = (v[i][1] - 2 * v[i + 1][1] + v[i + 2][1]);
          const float Lxy = (v[i][2] - v[i][0] + v[i + 2][0] - v[i + 2][2]) / 4.0f;
          out = (Lxx * Lyy - Lxy * Lxy) * norm2;
        }
        resp[(size_t)gy * w + gx] = out;
      }
    }
  }
  // (c) half-size image of the blurred tile: lerp form a + (b - a) * 0.5, x then y (k_half)
  if (half_out != nullptr) {
    const int ox = (x0 >> 1) + (tid & 31), oy = (y0 >> 1) + (tid >> 5);
    if (ox < ow && oy < oh) {
      const int ya = min(2 * oy, h - 1) - y0 + 1, yb = min(2 * oy + 1, h - 1) - y0 + 1;
      const int xa = min(2 * ox, w - 1) - x0 + 1, xb = min(2 * ox + 1, w - 1) - x0 + 1;
      const float a = Cc[ya * BW + xa], b = Cc[ya * BW + xb];
      const float c = Cc[yb * BW + xa], d = Cc[yb * BW + xb];
      const float r0 = a + (b - a) * 0.5f;
      const float r1 = c + (d - c) * 0.5f;
      half_out[(size_t)oy * ow + ox] = r0 + (r1 - r0) * 0.5f;
    }
  }
}

// Hessian response of an image already in HBM (first level of octaves >= 1).
__global__ void k_response(const float* __restrict__ src, float* __restrict__ resp, int w, int h, float norm2) {
  int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
  if (gx >= w || gy >= h) return;
  float v = 0.f;
  if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) v = hessian_at(src + (size_t)gy * w + gx, w, norm2);
  resp[(size_t)gy * w + gx] = v;
}

// pyramid.cpp:476 cv::resize(.., 0.5, 0.5, INTER_LINEAR): lerp form a+(b-a)*0.5, x then y.
__global__ void k_half(const float* __restrict__ in, int w, int h, float* __restrict__ out, int ow, int oh) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  int y0 = min(2 * y, h - 1), y1 = min(2 * y + 1, h - 1);
  int xa = min(2 * x, w - 1), xb = min(2 * x + 1, w - 1);
  float a = in[(size_t)y0 * w + xa], b = in[(size_t)y0 * w + xb];
  float c = in[(size_t)y1 * w + xa], d = in[(size_t)y1 * w + xb];
  float r0 = a + (b - a) * 0.5f;
  float r1 = c + (d - c) * 0.5f;
  out[(size_t)y * ow + x] = r0 + (r1 - r0) * 0.5f;
}

// synth-detection.cpp:344-351: (B+G+R)/3.0 evaluated as convertTo(alpha = 1/3)
__global__ void k_gray_from_bgr(const uint8_t* __restrict__ bgr, float* __restrict__ gray, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float third = (float)(1.0 / 3.0);
  float s = ((float)bgr[3 * i] + (float)bgr[3 * i + 1]) + (float)bgr[3 * i + 2];
  gray[i] = s * third;
}

__global__ void k_fill_i32(int* p, long n, int v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- NMS + localisation -------------------------------------------------------------------

struct Cand {           // one localised keypoint awaiting octave-map resolution
  float x, y, s, response;
  int type, octave, level, r0, c0, r, c;
  unsigned order;       // global visiting order: octave_base + (level-1)*w*h + r0*w + c0
  int map_key;          // (level-1)*w*h + r0*w + c0  (per-octave)
  int map_off;          // offset of (r,c) inside the octave-map workspace
};

// helpers.cpp:309-368 solveLinear3x3 (pivoted Gauss, float)
__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }
__device__ void solveLinear3x3(float* A, float* b) {
  int i = 0, prow = 0;
  float vp = fabsf(A[0]);
  float tmp = fabsf(A[3]);
  if (tmp > vp) { prow = 3; i = 1; vp = tmp; }
  if (fabsf(A[6]) > vp) { prow = 6; i = 2; }
  if (prow != 0) {
    swapf(A[prow], A[0]); swapf(A[prow + 1], A[1]); swapf(A[prow + 2], A[2]);
    swapf(b[i], b[0]);
  }
  vp = A[3] / A[0]; A[4] -= vp * A[1]; A[5] -= vp * A[2]; b[1] -= vp * b[0];
  vp = A[6] / A[0]; A[7] -= vp * A[1]; A[8] -= vp * A[2]; b[2] -= vp * b[0];
  if (fabsf(A[4]) < fabsf(A[7])) {
    swapf(A[7], A[4]); swapf(A[8], A[5]); swapf(b[2], b[1]);
  }
  vp = A[7] / A[4]; A[8] -= vp * A[5]; b[2] -= vp * b[1];
  b[2] = (b[2]) / A[8];
  b[1] = (b[1] - A[5] * b[2]) / A[4];
  b[0] = (b[0] - A[2] * b[2] - A[1] * b[1]) / A[0];
}

struct LevelArgs {
  const float* low; const float* cur; const float* high; const float* blur;
  float curScale; int level;
};
struct NmsArgs {
  LevelArgs lv[8];
  int nlev;
  int w, h, border, octave;
  float pixelDistance;
  float posThr, negThr, finalThr;
  double edgeThr;
  int numberOfScales;
  unsigned octave_base;
  int map_base;
};

// pyramid.cpp:281-403 localizeKeypoint + :65-124 getPointType
__device__ void localize(const NmsArgs& a, const LevelArgs& L, int li, int r, int c, int* map, Cand* cands,
                         int* ncand, int cap) {
  const int w = a.w, cols = a.w, rows = a.h;
  const int r0 = r, c0 = c;
  float b[3] = {0.f, 0.f, 0.f};
  float val = 0.f;
  int nr = r, nc = c;
  for (int iter = 0; iter < 5; iter++) {
    r = nr; c = nc;
    const float* cur0 = L.cur + (size_t)(r - 1) * w; const float* cur1 = L.cur + (size_t)r * w; const float* cur2 = L.cur + (size_t)(r + 1) * w;
    const float* low0 = L.low + (size_t)(r - 1) * w; const float* low1 = L.low + (size_t)r * w; const float* low2 = L.low + (size_t)(r + 1) * w;
    const float* high0 = L.high + (size_t)(r - 1) * w; const float* high1 = L.high + (size_t)r * w; const float* high2 = L.high + (size_t)(r + 1) * w;
    float dxx = cur1[c - 1] - 2.0f * cur1[c] + cur1[c + 1];
    float dyy = cur0[c] - 2.0f * cur1[c] + cur2[c];
    float dss = low1[c] - 2.0f * cur1[c] + high1[c];
    float dxy = 0.25f * (cur2[c + 1] - cur2[c - 1] - cur0[c + 1] + cur0[c - 1]);
    if (0 == iter) {
      float edgeScore = (dxx + dyy) * (dxx + dyy) / (dxx * dyy - dxy * dxy);
      if ((double)edgeScore >= a.edgeThr || edgeScore < 0) return;
    }
    float dxs = 0.25f * (high1[c + 1] - high1[c - 1] - low1[c + 1] + low1[c - 1]);
    float dys = 0.25f * (high2[c] - high0[c] - low2[c] + low0[c]);
    float A[9] = {dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss};
    float dx = 0.5f * (cur1[c + 1] - cur1[c - 1]);
    float dy = 0.5f * (cur2[c] - cur0[c]);
    float ds = 0.5f * (high1[c] - low1[c]);
    b[0] = -dx; b[1] = -dy; b[2] = -ds;
    solveLinear3x3(A, b);
    if (isnan(b[0]) || isnan(b[1]) || isnan(b[2])) return;
    val = cur1[c] + 0.5f * (dx * b[0] + dy * b[1] + ds * b[2]);
    if ((double)b[0] > 0.6) { if (c < cols - 3) nc++; else return; }
    if ((double)b[1] > 0.6) { if (r < rows - 3) nr++; else return; }
    if ((double)b[0] < -0.6) { if (c > 3) nc--; else return; }
    if ((double)b[1] < -0.6) { if (r > 3) nr--; else return; }
    if (nr == r && nc == c) break;
  }
  if (fabsf(b[0]) > 1.5f || fabsf(b[1]) > 1.5f || fabsf(b[2]) > 1.5f || fabsf(val) < a.finalThr) return;
  float e = b[2] / (float)a.numberOfScales;
  float scale = L.curScale * (float)exp2((double)e);
  int type;
  if (val < 0) type = 2;
  else {
    const float* ptr = L.blur + (size_t)r * w + c;
    float Lxx = (ptr[-1] - 2 * ptr[0] + ptr[1]);
    type = (Lxx < 0) ? 0 : 1;
  }
  int slot = atomicAdd(ncand, 1);
  if (slot >= cap) return;
  Cand k;
  k.x = a.pixelDistance * ((float)c + b[0]);
  k.y = a.pixelDistance * ((float)r + b[1]);
  k.s = a.pixelDistance * scale;
  k.response = val;
  k.type = type; k.octave = a.octave; k.level = L.level;
  k.r0 = r0; k.c0 = c0; k.r = r; k.c = c;
  k.map_key = li * (a.w * a.h) + r0 * a.w + c0;
  k.order = a.octave_base + (unsigned)k.map_key;
  k.map_off = a.map_base + r * a.w + c;
  cands[slot] = k;
  atomicMin(map + k.map_off, k.map_key);
}

// pyramid.cpp:405-425 findLevelKeypoints (+ isMax/isMin :41-63), all NMS levels of one octave
// in one launch (blockIdx.z = level index).
__global__ void k_nms_localize(NmsArgs a, int* map, Cand* cands, int* ncand, int cap) {
  const int li = blockIdx.z;
  const LevelArgs& L = a.lv[li];
  const int c = a.border + blockIdx.x * blockDim.x + threadIdx.x;
  const int r = a.border + blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= a.w - a.border || r >= a.h - a.border) return;
  const int w = a.w;
  const float val = L.cur[(size_t)r * w + c];
  const bool pos = val > a.posThr, neg = val < a.negThr;
  if (!pos && !neg) return;
  // isMax / isMin over the 3 x 3 x 3 neighbourhood.  The tests are order independent (all 26 must pass), so the loads of a
  // plane are issued together: the early-exit loop of the first version was a chain of up to 27 dependent L2 round trips
  // per surviving thread and made every launch ~16 us whatever the level's size.
  {
    const float* q = L.cur + (size_t)(r - 1) * w + (c - 1);
    float v[9];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) v[j * 3 + i] = q[j * w + i];
    bool ok = true;
#pragma unroll
    for (int t = 0; t < 9; t++) ok = ok && !(pos ? (v[t] > val) : (v[t] < val));
    if (!ok) return;
  }
  {
    const float* ql = L.low + (size_t)(r - 1) * w + (c - 1);
    const float* qh = L.high + (size_t)(r - 1) * w + (c - 1);
    float v[18];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) { v[j * 3 + i] = ql[j * w + i]; v[9 + j * 3 + i] = qh[j * w + i]; }
    bool ok = true;
#pragma unroll
    for (int t = 0; t < 18; t++) ok = ok && !(pos ? (v[t] > val) : (v[t] < val));
    if (!ok) return;
  }
  localize(a, L, li, r, c, map, cands, ncand, cap);
}

// ---- in-pyramid Baumberg iteration (affine.cpp:26-158, doBaumberg = 1, AFF_BMBRG_SMM) ------------------------------
// One warp per surviving candidate.  The 19 x 19 window is resampled from prevBlur (the level below the response
// level, pyramid.cpp:402 / SURVEY Q12) with lane = window row (coordinates accumulated in the reference's order),
// the three second-moment sums run in raster order on three lanes, the 2x2 algebra on every lane redundantly.
struct OctTab { long long off[16]; int w[16], h[16]; };
struct AffPars { int maxIterations; float convergenceThreshold; int smmWindowSize; float initialSigma; };
constexpr int SMM_MAX = 25;

__device__ __forceinline__ float sample_image_d(const float* im, int w, int h, float WX, float WY) {
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
// helpers.cpp:461-503
__device__ void invSqrt_d(float& a, float& b, float& c, float& l1, float& l2) {
  double t, r;
  if (b != 0) {
    r = double(c - a) / (2 * b);
    if (r >= 0) t = 1.0 / (r + sqrt(1 + r * r));
    else t = -1.0 / (-r + sqrt(1 + r * r));
    r = 1.0 / sqrt(1 + t * t);
    t = t * r;
  } else { r = 1; t = 0; }
  double x, z, d;
  x = 1.0 / sqrt(r * r * a - 2 * r * t * b + t * t * c);
  z = 1.0 / sqrt(t * t * a + 2 * r * t * b + r * r * c);
  d = sqrt(x * z);
  x /= d; z /= d;
  if (x < z) { l1 = float(z); l2 = float(x); }
  else { l1 = float(x); l2 = float(z); }
  a = float(r * r * x + t * t * z);
  b = float(-r * t * x + t * r * z);
  c = float(t * t * x + r * r * z);
}

__global__ void __launch_bounds__(128)
k_baumberg(const float* __restrict__ pyr, OctTab ot, const Cand* __restrict__ cands, const int* __restrict__ ncand_p, int cap,
           unsigned long long* __restrict__ keys, const float* __restrict__ mask, AffPars ap, float* __restrict__ candA,
           int* __restrict__ nkept) {
  __shared__ float img_s[4][SMM_MAX * SMM_MAX];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + wl;
  const int n = min(*ncand_p, cap);
  if (i >= n) return;
  if (keys[i] == ~0ull) return;
  const Cand k = cands[i];
  const int w = ot.w[k.octave], h = ot.h[k.octave];
  const float* blur = pyr + ot.off[k.octave] + (size_t)w * h * (k.level - 1);   // prevBlur
  const float pixelDistance = (float)(1 << k.octave);
  float* img = img_s[wl];
  float eigen_ratio_act = 0.0f, eigen_ratio_bef = 0.0f;
  float u11 = 1.0f, u12 = 0.0f, u21 = 0.0f, u22 = 1.0f, l1 = 1.0f, l2 = 1.0f;
  const float lx = k.x / pixelDistance, ly = k.y / pixelDistance;
  const float ratio = k.s / (ap.initialSigma * pixelDistance);
  const int ws = ap.smmWindowSize, maskPixels = ws * ws, half = ws / 2;
  bool found = false;
  for (int l = 0; l < ap.maxIterations; l++) {
    const float a11 = u11 * ratio, a12 = u12 * ratio, a21 = u21 * ratio, a22 = u22 * ratio;
    for (int j = lane; j < ws; j += 32) {          // interpolate(blur, lx, ly, U*ratio, img), helpers.cpp:551-626
      float rx = lx - (float)half * a12, ry = ly - (float)half * a22;
      for (int t = 0; t < j; t++) { rx += a12; ry += a22; }
      float WX = rx - (float)half * a11, WY = ry - (float)half * a21;
      for (int q = 0; q < ws; q++) {
        img[j * ws + q] = sample_image_d(blur, w, h, WX, WY);
        WX += a11; WY += a21;
      }
    }
    __syncwarp();
    float acc = 0.f;                                 // lanes 0, 1, 2: sum of gx*gx*v, gx*gy*v, gy*gy*v in raster order
    if (lane < 3) {
      for (int r = 0; r < ws; ++r)
        for (int c = 0; c < ws; ++c) {
          float xgrad, ygrad;                        // computeGradient, helpers.cpp:779-797
          if (c == 0) xgrad = img[r * ws + c + 1] - img[r * ws + c];
          else if (c == ws - 1) xgrad = img[r * ws + c] - img[r * ws + c - 1];
          else xgrad = img[r * ws + c + 1] - img[r * ws + c - 1];
          if (r == 0) ygrad = img[(r + 1) * ws + c] - img[r * ws + c];
          else if (r == ws - 1) ygrad = img[r * ws + c] - img[(r - 1) * ws + c];
          else ygrad = img[(r + 1) * ws + c] - img[(r - 1) * ws + c];
          const float v = mask[r * ws + c];
          if (lane == 0) acc += xgrad * xgrad * v;
          else if (lane == 1) { const float gxy = xgrad * ygrad; acc += gxy * v; }
          else acc += ygrad * ygrad * v;
        }
    }
    float a = __shfl_sync(0xffffffffu, acc, 0), b = __shfl_sync(0xffffffffu, acc, 1), c = __shfl_sync(0xffffffffu, acc, 2);
    __syncwarp();
    a /= (float)maskPixels; b /= (float)maskPixels; c /= (float)maskPixels;
    invSqrt_d(a, b, c, l1, l2);
    if ((a != a) || (b != b) || (c != c)) break;
    eigen_ratio_bef = eigen_ratio_act;
    eigen_ratio_act = (float)(1.0 - (double)(l2 / l1));
    const float u11t = u11, u12t = u12;
    u11 = a * u11t + b * u21;
    u12 = a * u12t + b * u22;
    u21 = b * u11t + c * u21;
    u22 = b * u12t + c * u22;
    {                                                // getEigenvalues, helpers.cpp:504-515
      const float trace = u11 + u22;
      const float delta1 = (trace * trace - 4 * (u11 * u22 - u12 * u21));
      if (delta1 < 0) break;
      const float delta = sqrtf(delta1);
      l1 = (trace + delta) / 2.0f;
      l2 = (trace - delta) / 2.0f;
    }
    if ((l1 / l2 > 6) || (l2 / l1 > 6)) break;
    if (eigen_ratio_act < ap.convergenceThreshold && eigen_ratio_bef < ap.convergenceThreshold) { found = true; break; }
  }
  if (lane == 0) {
    if (found) {
      candA[4 * (size_t)i + 0] = u11; candA[4 * (size_t)i + 1] = u12; candA[4 * (size_t)i + 2] = u21; candA[4 * (size_t)i + 3] = u22;
      atomicAdd(nkept, 1);
    } else keys[i] = ~0ull;
  }
}

// Octave-map resolution + export order: a candidate survives iff it holds the minimum visiting key
// at its final (r,c); survivors are ranked by (|response| desc, visiting order asc) -- the order a
// stable sort by |response| gives the reference's push order (scale-space-detector.hpp:120-131).
__device__ __forceinline__ unsigned long long sort_key(const Cand& k) {
  unsigned ab = __float_as_uint(fabsf(k.response));
  return ((unsigned long long)(~ab) << 32) | (unsigned long long)k.order;
}
__global__ void k_resolve(const Cand* cands, const int* ncand_p, int cap, const int* map,
                          unsigned long long* keys, int* nkept) {
  int n = min(*ncand_p, cap);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Cand& k = cands[i];
  bool keep = map[k.map_off] == k.map_key;
  keys[i] = keep ? sort_key(k) : ~0ull;
  if (keep) atomicAdd(nkept, 1);
}
// rank of a candidate = number of smaller keys (keys are unique: the low word is the visiting order).  64 candidates per
// CTA, FOUR lanes per candidate (each compares a quarter of every 256-key tile): the scan over all n keys is the kernel's
// whole latency and was 57 us per image with one thread per candidate and every CTA of the cap-sized grid walking it.
__global__ void __launch_bounds__(256)
k_rank_export(const Cand* cands, const int* ncand_p, int cap, const unsigned long long* keys,
              modsgpu_keypoint* out, const float* candA, float* outA) {
  __shared__ unsigned long long tile[256];
  const int n = min(*ncand_p, cap);
  if ((int)blockIdx.x * 64 >= n) return;                 // whole CTA beyond the list
  const int i = blockIdx.x * 64 + (threadIdx.x >> 2), q = threadIdx.x & 3;
  const unsigned long long mine = i < n ? keys[i] : ~0ull;
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    const int j = base + threadIdx.x;
    tile[threadIdx.x] = j < n ? keys[j] : ~0ull;         // padding keys are never smaller than a live key
    __syncthreads();
#pragma unroll 16
    for (int t = 0; t < 64; t++) rank += tile[4 * t + q] < mine;
    __syncthreads();
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  if (q == 0 && i < n && mine != ~0ull) {
    const Cand& k = cands[i];
    modsgpu_keypoint o;
    o.x = k.x; o.y = k.y; o.s = k.s; o.response = k.response; o.type = k.type; o.octave = k.octave;
    o.level = k.level; o.r0 = k.r0; o.c0 = k.c0; o.r = k.r; o.c = k.c; o.seq = 0;
    out[rank] = o;
    if (candA) for (int e = 0; e < 4; e++) outA[4 * (size_t)rank + e] = candA[4 * (size_t)i + e];
  }
}

int blur_smem_bytes(int ks) {
  int r = ks >> 1;
  int AW = BT_X + 2 + 2 * r, AH = BT_Y + 2 + 2 * r;
  return (AW * AH + (BT_X + 2) * AH + (BT_X + 2) * (BT_Y + 2)) * (int)sizeof(float);
}

template <int KS>
static void launch_blur3(modsgpu_ctx* ctx, dim3 grid, const float* src, float* dst, float* resp, int w, int h, const Taps& taps,
                         float norm2, float* resp_src, float norm2_src, float* half_out, int ow, int oh) {
  static OnceFlags attr_set;
  if (attr_set.need(ctx->device)) {
    cudaFuncSetAttribute(k_blur3<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, B3<KS>::SMEM);
    attr_set.set(ctx->device);
  }
  k_blur3<KS><<<grid, BT_NT, B3<KS>::SMEM, ctx->stream>>>(src, dst, resp, w, h, taps, norm2, resp_src, norm2_src, half_out, ow, oh);
}

// `resp_src` (response of the SOURCE image, first level of octaves >= 1) and `half_out` (half-size copy of the blurred
// image) are fused into the blur when the ksize-specialised kernel applies; *fused tells the caller whether they were.
int launch_blur(modsgpu_ctx* ctx, const float* src, float* dst, float* resp, int w, int h, float sigma, float norm2,
                float* resp_src = nullptr, float norm2_src = 0.f, float* half_out = nullptr, int ow = 0, int oh = 0,
                bool* fused = nullptr) {
  std::vector<float> t;
  int ks = mg_gaussian_taps(sigma, t);
  if (ks > MAX_KS) MG_FAIL(ctx, MODSGPU_EINVAL, "gaussian kernel too wide for the pyramid blur (ksize > 63)");
  Taps taps;
  memset(&taps, 0, sizeof(taps));
  taps.ks = ks;
  for (int i = 0; i < ks; i++) taps.k[i] = t[i];
  static OnceFlags attr_set;
  if (attr_set.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_blur_resp, cudaFuncAttributeMaxDynamicSharedMemorySize, blur_smem_bytes(MAX_KS)));
    MG_CUDA(ctx, cudaFuncSetAttribute(k_blur_resp2, cudaFuncAttributeMaxDynamicSharedMemorySize, blur2_smem_bytes(MAX_KS)));
    attr_set.set(ctx->device);
  }
  static const bool no_blur3 = [] { const char* e = getenv("MODSGPU_NO_BLUR3"); return e && atoi(e) != 0; }();
  dim3 grid(ceil_div(w, BT_X), ceil_div(h, BT_Y));
  const bool spec = !no_blur3 && ks >= 7 && ks <= 23;
  if (fused) *fused = spec;
  double bytes = (double)w * h * 4.0 * (resp ? 3 : 2);
  if (spec && resp_src) bytes += (double)w * h * 4.0;
  if (spec && half_out) bytes += (double)ow * oh * 4.0;
  MG_PROF(ctx, resp ? "k_blur_resp" : "k_blur", 0, bytes);
  if (spec) {
#define B3_CASE(K) case K: launch_blur3<K>(ctx, grid, src, dst, resp, w, h, taps, norm2, resp_src, norm2_src, half_out, ow, oh); break;
    switch (ks) {
      B3_CASE(7) B3_CASE(9) B3_CASE(11) B3_CASE(13) B3_CASE(15) B3_CASE(17) B3_CASE(19) B3_CASE(21) B3_CASE(23)
    }
#undef B3_CASE
  } else if (ks >= 7) {
    k_blur_resp2<<<grid, BT_NT, blur2_smem_bytes(ks), ctx->stream>>>(src, dst, resp, w, h, taps, norm2);
  } else {
    k_blur_resp<<<grid, BT_THREADS, blur_smem_bytes(ks), ctx->stream>>>(src, dst, resp, w, h, taps, norm2);
  }
  MG_LAUNCHED(ctx);
  return 0;
}

}  // namespace

int mg_gaussian_taps(float sigmaf, std::vector<float>& taps) {
  double sigma = (double)sigmaf;
  int ks = (int)(2.0 * 3.0 * sigma + 1.0);
  if (ks % 2 == 0) ks++;
  if (ks < 1) ks = 1;
  int r = ks / 2;
  std::vector<double> kd(ks);
  double sum = 0;
  for (int i = 0; i < ks; i++) {
    double x = i - r;
    kd[i] = std::exp(-x * x / (2.0 * sigma * sigma));
    sum += kd[i];
  }
  taps.resize(ks);
  for (int i = 0; i < ks; i++) taps[i] = (float)(kd[i] / sum);
  return ks;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" void modsgpu_default_pyr_params(modsgpu_pyr_params* p) {
  p->numberOfScales = 3; p->initialSigma = 1.6f; p->threshold = 5.33f; p->edgeEigenValueRatio = 10.0; p->border = 5;
  p->detectorMode = MODSGPU_FIXED_TH; p->rel_threshold = -1.f; p->reg_number = -1; p->rel_reg_number = -1.f;
}

extern "C" int modsgpu_image_from_bgr8(modsgpu_ctx* ctx, const uint8_t* bgr, int w, int h, modsgpu_image** out) {
  if (!ctx || !bgr || w <= 0 || h <= 0 || !out) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  size_t n = (size_t)w * h;
  MG_CUDA(ctx, ctx->io_a.ensure(n * 3));
  modsgpu_image* img = new modsgpu_image();
  img->w = w; img->h = h;
  // every error path below hands the image (and its device buffer, if any) back
  auto body = [&]() -> int {
    MG_CUDA(ctx, mg_image_alloc(ctx, n * sizeof(float), &img->d));
    MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, bgr, n * 3, cudaMemcpyHostToDevice, ctx->stream));
    MG_PROF(ctx, "k_gray_from_bgr", 0, (double)n * 7.0);
    k_gray_from_bgr<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->io_a.as<uint8_t>(), img->d, (long)n);
    MG_LAUNCHED(ctx);
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  };
  if (const int rc = body()) { modsgpu_image_free(ctx, img); return rc; }
  *out = img;
  return 0;
}

extern "C" int modsgpu_image_from_gray32f(modsgpu_ctx* ctx, const float* gray, int w, int h, int stride, modsgpu_image** out) {
  if (!ctx || !gray || w <= 0 || h <= 0 || stride < w || !out) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  modsgpu_image* img = new modsgpu_image();
  img->w = w; img->h = h;
  auto body = [&]() -> int {
    MG_CUDA(ctx, mg_image_alloc(ctx, (size_t)w * h * sizeof(float), &img->d));
    MG_CUDA(ctx, cudaMemcpy2DAsync(img->d, (size_t)w * 4, gray, (size_t)stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, ctx->stream));
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  };
  if (const int rc = body()) { modsgpu_image_free(ctx, img); return rc; }
  *out = img;
  return 0;
}

extern "C" int modsgpu_image_download(modsgpu_ctx* ctx, const modsgpu_image* img, float* gray) {
  if (!ctx || !img || !gray) return MODSGPU_EINVAL;
  MG_CUDA(ctx, cudaMemcpyAsync(gray, img->d, (size_t)img->w * img->h * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  return 0;
}

extern "C" void modsgpu_image_size(const modsgpu_image* img, int* w, int* h) { *w = img->w; *h = img->h; }

cudaError_t mg_image_alloc(modsgpu_ctx* ctx, size_t bytes, float** out) {
  for (size_t i = 0; i < ctx->img_pool.size(); i++)
    if (ctx->img_pool[i].second == bytes) {
      *out = ctx->img_pool[i].first;
      ctx->img_pool.erase(ctx->img_pool.begin() + i);
      return cudaSuccess;
    }
  return cudaMalloc((void**)out, bytes);
}

extern "C" void modsgpu_image_free(modsgpu_ctx* ctx, modsgpu_image* img) {
  if (!img) return;
  if (img->d) {
    if (ctx && ctx->img_pool.size() < 16) ctx->img_pool.emplace_back(img->d, (size_t)img->w * img->h * sizeof(float));
    else cudaFree(img->d);
  }
  delete img;
}

extern "C" void modsgpu_free(void* p) { free(p); }

extern "C" int modsgpu_gaussian_blur(modsgpu_ctx* ctx, const float* in, float* out, int w, int h, float sigma) {
  if (!ctx || !in || !out || w <= 0 || h <= 0) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  size_t bytes = (size_t)w * h * 4;
  MG_CUDA(ctx, ctx->io_a.ensure(bytes));
  MG_CUDA(ctx, ctx->io_b.ensure(bytes));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = launch_blur(ctx, ctx->io_a.as<float>(), ctx->io_b.as<float>(), nullptr, w, h, sigma, 0.f);
  if (rc) return rc;
  MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

extern "C" int modsgpu_hessian_response(modsgpu_ctx* ctx, const float* in, float* out, int w, int h, float norm) {
  if (!ctx || !in || !out || w <= 0 || h <= 0) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  size_t bytes = (size_t)w * h * 4;
  MG_CUDA(ctx, ctx->io_a.ensure(bytes));
  MG_CUDA(ctx, ctx->io_b.ensure(bytes));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grid(ceil_div(w, 32), ceil_div(h, 8));
  k_response<<<grid, blk, 0, ctx->stream>>>(ctx->io_a.as<float>(), ctx->io_b.as<float>(), w, h, norm * norm);
  MG_LAUNCHED(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

static void half_size(int w, int h, int* ow, int* oh) {
  *ow = (int)std::nearbyint(w * 0.5);
  *oh = (int)std::nearbyint(h * 0.5);
}

extern "C" int modsgpu_half_image(modsgpu_ctx* ctx, const float* in, int w, int h, float* out) {
  if (!ctx || !in || !out || w <= 0 || h <= 0) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  int ow, oh;
  half_size(w, h, &ow, &oh);
  MG_CUDA(ctx, ctx->io_a.ensure((size_t)w * h * 4));
  MG_CUDA(ctx, ctx->io_b.ensure((size_t)ow * oh * 4 + 4));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, in, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grid(ceil_div(ow, 32), ceil_div(oh, 8));
  k_half<<<grid, blk, 0, ctx->stream>>>(ctx->io_a.as<float>(), w, h, ctx->io_b.as<float>(), ow, oh);
  MG_LAUNCHED(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_b.p, (size_t)ow * oh * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

// Device part of the detector: enqueues the whole pyramid on ctx->stream; the sorted keypoints end
// up in ctx->det_out and their count in ctx->det_misc[1].  No host synchronisation inside.
int mg_detect_enqueue(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p, int cap,
                      const modsgpu_affshape_params* aff) {
  const int nS = p->numberOfScales;
  if (nS < 1 || nS + 2 > 8) MG_FAIL(ctx, MODSGPU_EINVAL, "numberOfScales out of range");
  if (p->border < 2) MG_FAIL(ctx, MODSGPU_EINVAL, "border must be >= 2 (pyramid.cpp:407)");
  const int nlev = nS + 2;
  // octave geometry (pyramid.cpp:520-528)
  std::vector<int> ow, oh;
  {
    int W = img->w, H = img->h, minSize = 2 * p->border + 2;
    while (H > minSize && W > minSize) {
      ow.push_back(W); oh.push_back(H);
      int nW, nH;
      half_size(W, H, &nW, &nH);
      W = nW; H = nH;
    }
  }
  const int nOct = (int)ow.size();
  size_t pyr_floats = 0, map_ints = 0;
  std::vector<size_t> oct_off(nOct), map_off(nOct);
  for (int o = 0; o < nOct; o++) {
    oct_off[o] = pyr_floats; map_off[o] = map_ints;
    size_t px = (size_t)ow[o] * oh[o];
    pyr_floats += px * 2 * nlev;
    map_ints += px;
  }
  if (map_ints * nS > 0xfffffff0ull) MG_FAIL(ctx, MODSGPU_EINVAL, "image too large");
  MG_CUDA(ctx, ctx->det_pyr.ensure((pyr_floats + 16) * sizeof(float)));
  MG_CUDA(ctx, ctx->det_map.ensure((map_ints + 16) * sizeof(int)));
  MG_CUDA(ctx, ctx->det_cand.ensure((size_t)cap * (sizeof(Cand) + sizeof(unsigned long long))));
  MG_CUDA(ctx, ctx->det_out.ensure((size_t)cap * (sizeof(modsgpu_keypoint) + 16)));   // + A (4 floats) per keypoint
  if (aff) MG_CUDA(ctx, ctx->det_aff.ensure((size_t)cap * 16 + (size_t)SMM_MAX * SMM_MAX * 4 + 64));
  MG_CUDA(ctx, ctx->det_misc.ensure(64));
  float* pyr = ctx->det_pyr.as<float>();
  int* map = ctx->det_map.as<int>();
  Cand* cands = ctx->det_cand.as<Cand>();
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(cands + cap);
  int* counters = ctx->det_misc.as<int>();  // [0] candidates, [1] kept
  MG_CUDA(ctx, cudaMemsetAsync(counters, 0, 64, ctx->stream));
  if (nOct == 0) return 0;
  k_fill_i32<<<(unsigned)((map_ints + 255) / 256), 256, 0, ctx->stream>>>(map, (long)map_ints, 0x7fffffff);
  MG_LAUNCHED(ctx);

  // thresholds, pyramid.h:46-66 (DET_HESSIAN, FIXED_TH)
  const double edgeThr = (p->edgeEigenValueRatio + 1.0f) * (p->edgeEigenValueRatio + 1.0f) / p->edgeEigenValueRatio;
  // every mode but FIXED_TH keeps all extrema and truncates after the sort (pyramid.h:58-59)
  const bool fixedTh = p->detectorMode == MODSGPU_FIXED_TH;
  const float posThr = fixedTh ? (float)(0.8 * p->threshold) : 0.f, negThr = -posThr;
  const float finalThr = fixedTh ? p->threshold * p->threshold : 0.f;
  const float sigmaStep = std::pow(2.0f, 1.0f / (float)nS);

  // fork / join across two streams (MODSGPU_DET_NO_FORK=1 keeps the linear chain; the per-kernel profiler needs it linear too)
  static const bool no_fork = [] { const char* e = getenv("MODSGPU_DET_NO_FORK"); return e && atoi(e) != 0; }();
  cudaStream_t const main_stream = ctx->stream;
  bool fork = !no_fork && !ctx->prof.on && nS + 1 >= 2;
  if (fork && !ctx->stream2) {
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->det_fork_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->det_join_ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); fork = false; }
  }
  bool side_used = false;
  float pixelDistance = 1.0f;
  unsigned octave_base = 0;
  for (int o = 0; o < nOct; o++) {
    const int w = ow[o], h = oh[o];
    const size_t px = (size_t)w * h;
    float* Lv = pyr + oct_off[o];             // nlev blur levels
    float* Rv = Lv + px * nlev;               // nlev responses
    float curSigma = p->initialSigma;
    // first level of the octave + its response (pyramid.cpp:445-447; :514-518 for octave 0)
    if (o == 0) {
      float norm = curSigma * curSigma;
      if (p->initialSigma > 0.5f) {
        float sigma = std::sqrt(p->initialSigma * p->initialSigma - 0.5f * 0.5f);
        int rc = launch_blur(ctx, img->d, Lv, Rv, w, h, sigma, norm * norm);
        if (rc) return rc;
      } else {
        MG_CUDA(ctx, cudaMemcpyAsync(Lv, img->d, px * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        dim3 blk(32, 8), grid(ceil_div(w, 32), ceil_div(h, 8));
        k_response<<<grid, blk, 0, ctx->stream>>>(Lv, Rv, w, h, norm * norm);
        MG_LAUNCHED(ctx);
      }
    }
    // first level of octaves >= 1: its response is produced by the blur that reads it (level 1), see launch_blur
    bool first_resp_pending = o > 0;
    NmsArgs na;
    memset(&na, 0, sizeof(na));
    na.nlev = 0;
    for (int i = 1; i < nS + 2; i++) {
      float sigma = curSigma * std::sqrt(sigmaStep * sigmaStep - 1.0f);
      float s2 = curSigma * sigmaStep;
      float norm = s2 * s2;
      // The last level of an octave (i = nS + 1) and its non-maximum suppression feed nothing downstream but the candidate
      // list: they run on the context's side stream while the main stream halves level nS and starts the next octave
      // (in the captured graph: a fork -- 30 instead of 41 kernels on the critical path of a 1024x768 image).
      if (fork && i == nS + 1) {
        MG_CUDA(ctx, cudaEventRecord(ctx->det_fork_ev, main_stream));
        MG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->det_fork_ev, 0));
        ctx->stream = ctx->stream2;
      }
      const bool want_half = i == nS && o + 1 < nOct;
      bool fused = false;
      {
        const float n0 = p->initialSigma * p->initialSigma;
        int rc = launch_blur(ctx, Lv + px * (i - 1), Lv + px * i, Rv + px * i, w, h, sigma, norm * norm,
                             first_resp_pending ? Rv : nullptr, n0 * n0,
                             want_half ? pyr + oct_off[o + 1] : nullptr, want_half ? ow[o + 1] : 0, want_half ? oh[o + 1] : 0, &fused);
        if (rc) { ctx->stream = main_stream; return rc; }
      }
      if (first_resp_pending && !fused) {
        const float n0 = p->initialSigma * p->initialSigma;
        dim3 blk(32, 8), grid(ceil_div(w, 32), ceil_div(h, 8));
        MG_PROF(ctx, "k_response", 0, (double)px * 8.0);
        k_response<<<grid, blk, 0, ctx->stream>>>(Lv, Rv, w, h, n0 * n0);
        MG_LAUNCHED(ctx);
      }
      first_resp_pending = false;
      if (i >= 2) {
        LevelArgs& L = na.lv[na.nlev++];
        L.low = Rv + px * (i - 2); L.cur = Rv + px * (i - 1); L.high = Rv + px * i;
        L.blur = Lv + px * (i - 1);
        L.curScale = curSigma; L.level = i - 1;
      }
      if (want_half && !fused) {
        dim3 blk(32, 8), grid(ceil_div(ow[o + 1], 32), ceil_div(oh[o + 1], 8));
        MG_PROF(ctx, "k_half", 0, (double)ow[o + 1] * oh[o + 1] * 20.0);
        k_half<<<grid, blk, 0, ctx->stream>>>(Lv + px * i, w, h, pyr + oct_off[o + 1], ow[o + 1], oh[o + 1]);
        MG_LAUNCHED(ctx);
      }
      curSigma *= sigmaStep;
    }
    na.w = w; na.h = h; na.border = p->border; na.octave = o;
    na.pixelDistance = pixelDistance;
    na.posThr = posThr; na.negThr = negThr; na.finalThr = finalThr; na.edgeThr = edgeThr;
    na.numberOfScales = nS;
    na.octave_base = octave_base;
    na.map_base = (int)map_off[o];
    int iw = w - 2 * p->border, ih = h - 2 * p->border;
    if (iw > 0 && ih > 0 && na.nlev > 0) {
      dim3 blk(32, 8), grid(ceil_div(iw, 32), ceil_div(ih, 8), na.nlev);
      MG_PROF(ctx, "k_nms_localize", 0, (double)px * 4.0 * (na.nlev + 2));
      k_nms_localize<<<grid, blk, 0, ctx->stream>>>(na, map, cands, counters, cap);      // side stream when forked
      ctx->launches++;
      if (ctx->prof.open) mg_prof_end(ctx);
      if (cudaGetLastError() != cudaSuccess) { ctx->stream = main_stream; MG_FAIL(ctx, MODSGPU_ECUDA, "k_nms_localize launch failed"); }
    }
    if (fork && ctx->stream != main_stream) {
      side_used = true;
      ctx->stream = main_stream;
    }
    octave_base += (unsigned)(px * nS);
    pixelDistance *= 2.0f;
  }
  if (side_used) {                        // join: the candidate list is complete once the side stream has drained
    MG_CUDA(ctx, cudaEventRecord(ctx->det_join_ev, ctx->stream2));
    MG_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->det_join_ev, 0));
  }
  int nb = ceil_div(cap, 256);
  MG_PROF(ctx, "k_resolve", 2, (double)cap);
  k_resolve<<<nb, 256, 0, ctx->stream>>>(cands, counters, cap, map, keys, counters + 1);
  MG_LAUNCHED(ctx);
  float* candA = nullptr;
  float* outA = nullptr;
  if (aff && aff->doBaumberg) {
    if (aff->smmWindowSize < 3 || aff->smmWindowSize > SMM_MAX || nOct > 16) MG_FAIL(ctx, MODSGPU_EINVAL, "smmWindowSize must be in [3, 25]");
    // computeGaussMask (helpers.cpp:413-440)
    const int size = aff->smmWindowSize, halfSize = size >> 1;
    std::vector<float> mask((size_t)size * size), tmp(halfSize + 1);
    const float scale = float(halfSize) / 3.0f, scale2 = -2.0f * scale * scale;
    for (int i = 0; i <= halfSize; i++) tmp[i] = std::exp((float(i * i) / scale2));
    const int endSize = int(std::ceil(scale * 5.0f) - halfSize);
    for (int i = 1; i < endSize; i++) tmp[halfSize - i] += std::exp((float((i + halfSize) * (i + halfSize)) / scale2));
    for (int i = 0; i <= halfSize; i++)
      for (int j = 0; j <= halfSize; j++) {
        const float v = tmp[i] * tmp[j];
        mask[(i + halfSize) * size + (-j + halfSize)] = v; mask[(-i + halfSize) * size + (j + halfSize)] = v;
        mask[(i + halfSize) * size + (j + halfSize)] = v; mask[(-i + halfSize) * size + (-j + halfSize)] = v;
      }
    candA = ctx->det_aff.as<float>();
    float* d_mask = candA + (size_t)cap * 4;
    MG_CUDA(ctx, ctx->h_stage2.ensure(mask.size() * 4));
    memcpy(ctx->h_stage2.p, mask.data(), mask.size() * 4);
    MG_CUDA(ctx, cudaMemcpyAsync(d_mask, ctx->h_stage2.p, mask.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    OctTab ot;
    memset(&ot, 0, sizeof(ot));
    for (int o = 0; o < nOct; o++) { ot.off[o] = (long long)oct_off[o]; ot.w[o] = ow[o]; ot.h[o] = oh[o]; }
    AffPars ap = {aff->maxIterations, aff->convergenceThreshold, aff->smmWindowSize, aff->initialSigma};
    MG_CUDA(ctx, cudaMemsetAsync(counters + 1, 0, 4, ctx->stream));   // kept is recounted after the iteration
    MG_PROF(ctx, "k_baumberg", 2, (double)cap);
    k_baumberg<<<ceil_div(cap, 4), 128, 0, ctx->stream>>>(pyr, ot, cands, counters, cap, keys, d_mask, ap, candA, counters + 1);
    MG_LAUNCHED(ctx);
    outA = reinterpret_cast<float*>(ctx->det_out.as<modsgpu_keypoint>() + cap);
  }
  MG_PROF(ctx, "k_rank_export", 2, (double)cap);
  k_rank_export<<<ceil_div(cap, 64), 256, 0, ctx->stream>>>(cands, counters, cap, keys, ctx->det_out.as<modsgpu_keypoint>(), candA, outA);
  MG_LAUNCHED(ctx);
  return 0;
}

// prepareKeysForExport (scale-space-detector.hpp:125-198) on the |response|-descending list: how many keys survive
static int keys_to_export(const modsgpu_keypoint* k, int n, const modsgpu_pyr_params* p, bool doBaumberg) {
  if (n <= 0 || p->detectorMode == MODSGPU_FIXED_TH) return n;
  auto count_above = [&](double thr) {      // lower_bound with responseCompareInvOrder: first key with |r| <= thr
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) / 2; if (std::fabs((double)k[mid].response) > thr) lo = mid + 1; else hi = mid; }
    return lo;
  };
  int keep = n;
  switch (p->detectorMode) {
    case MODSGPU_RELATIVE_TH: {
      // effectiveThreshold is a float member, tempKey.response a double
      const float eff = (float)(std::fabs((double)k[0].response) * (double)p->rel_threshold);
      keep = count_above(std::fabs((double)eff));
      break;
    }
    case MODSGPU_FIXED_REG_NUMBER: {
      int nr = p->reg_number;
      if (doBaumberg) nr = (int)std::floor(3.0 * (double)nr);
      if (nr < n && nr >= 0) keep = nr;
      if (keep > p->reg_number) keep = p->reg_number;       // the closing clause of the function (:194-195)
      break;
    }
    case MODSGPU_RELATIVE_REG_NUMBER:
      keep = (int)std::floor((double)p->rel_reg_number * (double)n);
      break;
    case MODSGPU_NOT_LESS_THAN_REGIONS: {
      const int fixed = count_above((double)p->threshold);   // NB the un-squared threshold (reference quirk, :174)
      keep = fixed < p->reg_number ? std::min(p->reg_number, n) : std::min(fixed, n);
      break;
    }
    default: break;
  }
  return std::max(0, std::min(keep, n));
}

extern "C" int modsgpu_reg_number_for_view(int reg_number, double tilt, double zoom) {
  if (tilt > 2.0 || zoom < 0.5) return (int)std::floor(zoom * (double)reg_number / tilt);
  return reg_number;
}

// The detector's launch sequence for a given image buffer, geometry and parameter block is static (41 launches per
// 1024x768 image, each 11-17 us apart when issued one by one): the second request with the same key is captured into a
// CUDA graph and replayed from then on -- one driver call per image instead of 41, and back-to-back kernels on the device.
// The first request runs uncaptured (it sizes the workspaces and sets the function attributes); requests under the
// per-kernel profiler, with the Baumberg stage (it uploads its mask from the host) or with MODSGPU_NO_GRAPHS=1 stay plain.
static int mg_detect_graph(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p, int cap,
                           const modsgpu_affshape_params* aff) {
  static const bool no_graphs = [] { const char* e = getenv("MODSGPU_NO_GRAPHS"); return e && atoi(e) != 0; }();
  if (no_graphs || aff || ctx->prof.on) return mg_detect_enqueue(ctx, img, p, cap, aff);
  modsgpu_pyr_params pk;                 // field by field into a zeroed copy: the caller's padding bytes are not part of the key
  memset(&pk, 0, sizeof(pk));
  pk.numberOfScales = p->numberOfScales; pk.initialSigma = p->initialSigma; pk.threshold = p->threshold;
  pk.edgeEigenValueRatio = p->edgeEigenValueRatio; pk.border = p->border; pk.detectorMode = p->detectorMode;
  pk.rel_threshold = p->rel_threshold; pk.reg_number = p->reg_number; pk.rel_reg_number = p->rel_reg_number;
  std::string key(reinterpret_cast<const char*>(&pk), sizeof(pk));
  const void* ptrs[6] = {img->d, ctx->det_pyr.p, ctx->det_map.p, ctx->det_cand.p, ctx->det_out.p, ctx->det_misc.p};
  const int dims[3] = {img->w, img->h, cap};
  key.append(reinterpret_cast<const char*>(ptrs), sizeof(ptrs));
  key.append(reinterpret_cast<const char*>(dims), sizeof(dims));
  modsgpu_ctx::DetGraph& g = ctx->det_graphs[key];
  g.uses++;
  if (g.exec) {
    MG_CUDA(ctx, cudaGraphLaunch(g.exec, ctx->stream));
    ctx->launches += g.launches;
    return 0;
  }
  if (g.uses < 2) return mg_detect_enqueue(ctx, img, p, cap, aff);   // NB the workspace pointers enter the key AFTER this call sized them
  if (ctx->det_graphs.size() > 64) {   // a caller cycling through many image buffers: do not hoard executables
    for (auto& e : ctx->det_graphs) if (e.second.exec) cudaGraphExecDestroy(e.second.exec);
    ctx->det_graphs.clear();
    return mg_detect_enqueue(ctx, img, p, cap, aff);
  }
  const long long l0 = ctx->launches;
  MG_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = mg_detect_enqueue(ctx, img, p, cap, aff);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  const long long nl = ctx->launches - l0;
  ctx->launches = l0;
  if (rc || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->det_graphs.erase(key);
    if (rc) return rc;
    return mg_detect_enqueue(ctx, img, p, cap, aff);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess || !exec) { cudaGetLastError(); ctx->det_graphs.erase(key); return mg_detect_enqueue(ctx, img, p, cap, aff); }
  modsgpu_ctx::DetGraph& g2 = ctx->det_graphs[key];
  g2.exec = exec; g2.launches = nl;
  MG_CUDA(ctx, cudaGraphLaunch(exec, ctx->stream));
  ctx->launches += nl;
  return 0;
}

// the per-view chain (chain.cu) runs the detector without the keypoint read-back
int mg_detect_device(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p, int cap) {
  return mg_detect_graph(ctx, img, p, cap, nullptr);
}
int mg_keys_to_export(const modsgpu_keypoint* k, int n, const modsgpu_pyr_params* p) { return keys_to_export(k, n, p, false); }

static int detect_impl(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p,
                       const modsgpu_affshape_params* aff, modsgpu_keypoint** out, float** A, int* n) {
  if (!ctx || !img || !p || !out || !n) return MODSGPU_EINVAL;
  if (p->detectorMode < MODSGPU_FIXED_TH || p->detectorMode > MODSGPU_NOT_LESS_THAN_REGIONS)
    MG_FAIL(ctx, MODSGPU_EINVAL, "unknown detectorMode");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  int cap = 1 << 16;
  for (;;) {
    int rc = mg_detect_graph(ctx, img, p, cap, aff);
    if (rc) return rc;
    MG_CUDA(ctx, ctx->h_stage.ensure(64));
    int* hc = ctx->h_stage.as<int>();
    MG_CUDA(ctx, cudaMemcpyAsync(hc, ctx->det_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    if (hc[0] > cap) { cap = hc[0] + hc[0] / 8; continue; }  // candidate list overflowed: redo with room
    int kept = hc[1];
    modsgpu_keypoint* res = (modsgpu_keypoint*)malloc(sizeof(modsgpu_keypoint) * (size_t)std::max(kept, 1));
    if (kept > 0)
      MG_CUDA(ctx, cudaMemcpyAsync(res, ctx->det_out.p, sizeof(modsgpu_keypoint) * (size_t)kept, cudaMemcpyDeviceToHost, ctx->stream));
    float* resA = nullptr;
    if (A) {
      resA = (float*)malloc(16 * (size_t)std::max(kept, 1));
      if (kept > 0)
        MG_CUDA(ctx, cudaMemcpyAsync(resA, ctx->det_out.as<modsgpu_keypoint>() + cap, 16 * (size_t)kept, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (mg_end(ctx)) { free(res); free(resA); return MODSGPU_ECUDA; }
    kept = keys_to_export(res, kept, p, aff != nullptr && aff->doBaumberg);
    *out = res; *n = kept;
    if (A) *A = resA;
    return 0;
  }
}

extern "C" int modsgpu_detect(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p,
                              modsgpu_keypoint** out, int* n) {
  return detect_impl(ctx, img, p, nullptr, out, nullptr, n);
}

extern "C" void modsgpu_default_affshape_params(modsgpu_affshape_params* a) {
  a->maxIterations = 16; a->convergenceThreshold = 0.05f; a->smmWindowSize = 19; a->initialSigma = 1.6f; a->doBaumberg = 1;
}

extern "C" int modsgpu_detect_affine(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p,
                                     const modsgpu_affshape_params* aff, modsgpu_keypoint** out, float** A, int* n) {
  if (!aff || !A) return MODSGPU_EINVAL;
  return detect_impl(ctx, img, p, aff, out, A, n);
}
