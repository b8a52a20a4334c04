// cnn.cu -- AffNet / OriNet / HardNet++ as fused sm_100a kernels (seam S2).
// Replaces the three ZeroMQ/PyTorch daemons (build/affnet_server.py:45-84, orinet_server.py:45-82,
// desc_server.py:58-92).  BatchNorm is folded into the convolutions by tools/export_weights.py.
//
// Data layout in HBM (all activations fp16, accumulation fp32):
//   An activation tensor with C channels over S x S maps is stored as C/8 "planes"; a plane is a flat
//   array of 16-byte slots (8 channels of one pixel).  Pixels of all patches of a chunk share one
//   linear slot index  g = patch*(S+1)^2 + (y+1)*(S+1) + x  : every row carries ONE trailing pad
//   slot and every patch ONE leading pad row, so a 3x3 tap (dy,dx) is the constant slot shift
//   dy*(S+1)+dx and all out-of-image reads land on pad slots (zero, never written).
//   => a 128-pixel M tile of the implicit GEMM is 128 consecutive slots, its im2col operand for one
//   tap is the SAME shared-memory tile read at a shifted start address -- expressed directly in the
//   tcgen05 shared-memory descriptor (K-major, no swizzle: rows 16 B apart).  No im2col copy exists.
//   Stride-2 layers read four parity planes (space-to-depth) written by the previous layer's epilogue,
//   which turns them into unit-stride shifts as well.
//
// Kernels:
//   k_conv1     CUDA cores: per-patch mean/std normalisation + conv1 (1->C1) + bias + ReLU
//   k_conv_umma tcgen05 implicit GEMM for conv2..conv6: warp-specialised (bulk-copy producer, single
//               thread MMA issuer, 4 epilogue warps), weights resident in smem, A tiles double buffered
//               by cp.async.bulk + mbarrier, accumulators double buffered in TMEM
//   k_head_gemm HardNet 8x8 conv (K = 8192 GEMM) + BN + L2 norm + uint8 quantisation epilogue
//   k_head_aff / k_head_ori  the tiny 8x8 heads (+tanh) on CUDA cores
#include "common.cuh"
#include "umma.cuh"
#include "npz.h"
#include <map>
#include <mutex>

using namespace umma;

namespace {

constexpr int FS = 64;  // zero slots in front of every plane (covers the negative tap shifts)
enum { OUT_NORMAL = 0, OUT_PARITY = 1, OUT_GEMM = 2 };

struct ConvW { __half* w = nullptr; float* b = nullptr; };

}  // namespace

struct NetWeights {
  int net = 0, C1 = 0, out_dim = 0;
  float* c1_w = nullptr; float* c1_b = nullptr;
  __half* c1_split = nullptr;    // conv1 weights for k_conv1_mma: [4 K-planes][C1][8] fp16, K = wh | wh | wl | 0
  ConvW conv[5];                 // conv2..conv6 in UMMA block order
  __half* head_w16 = nullptr;    // HardNet head, UMMA block order
  float* head_w32 = nullptr;     // AffNet / OriNet head, k_head_small's shared-memory layout
  float* head_b = nullptr;
  int cap = 0;                   // patches per chunk the activation buffers hold
  bool fused12 = false;          // conv1 weights resident in c_conv1[net] on this device -> k_conv12 path
  __half* act[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t slots[6] = {0, 0, 0, 0, 0, 0};
  const NetWeights* origin = nullptr;   // clone of a sibling context: the weights belong to *origin, only act[] is owned
};

namespace {

__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// {max(lo,0), max(hi,0)} rounded to nearest-even fp16, lo in the low half: one F2FP with the .relu modifier
// (same values as __float2half_rn(fmaxf(x, 0.f)) for every finite x)
__device__ __forceinline__ uint32_t relu_pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// =================================================================================================
// conv1 (+ input normalisation) on CUDA cores
//
// Persistent, warp-specialised: CTAs (three per SM, 4 conv warps + 1 prep warp: 15 warps leave 128 registers per thread)
// walk over patches; warp 4 ("prep") loads the next patch (one 32-pixel row per lane), reduces its mean / unbiased std
// with shuffles and writes the normalised pixels into the other half of a double-buffered padded tile, while warps 0-3
// run the FMA chains of the current patch (8 independent chains per thread, taps outermost).  The 72 weights + 8 biases of a
// thread's 8-channel group are loaded into registers once per CTA.  (The first version ran one CTA per patch: global load
// -> block reduction -> fp64 section on one thread -> 3 barriers were exposed in front of every patch and the issue slots
// were 56 % busy, profiles/r01_ncu_full_k_conv1_v5.txt.)
// =================================================================================================
// Patch count of a launch whose exact size only the device knows (the region lists between the net passes are compacted
// on the device, csrc/chain.cu): the host passes its upper bound `np`, the kernel clamps it to the live count of this
// chunk.  cnt_dev == nullptr: np is exact.
__device__ __forceinline__ int live_patches(int np, const int* __restrict__ cnt_dev, int cnt_base) {
  if (cnt_dev == nullptr) return np;
  const int live = *cnt_dev - cnt_base;
  return live < 0 ? 0 : (live < np ? live : np);
}

constexpr int C1_PW_ = 40;
constexpr int C1_PW = 40;   // padded row: pixel x lives at column x + 4 (16-byte aligned float4 stores), halo at 3 and 36
template <int C1>
__global__ void __launch_bounds__(160, 3)
k_conv1(const uint8_t* __restrict__ patches, int np, const float* __restrict__ w, const float* __restrict__ b,
        __half* __restrict__ out, size_t out_slots, const int* __restrict__ cnt_dev, int cnt_base) {
  np = live_patches(np, cnt_dev, cnt_base);
  __shared__ __align__(16) float P[2][34][C1_PW];
  __shared__ __align__(16) float bs[C1];
  __shared__ uint64_t bars[4];            // full[2], empty[2]
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  for (int i = tid; i < 2 * 34 * C1_PW; i += 160) (&P[0][0][0])[i] = 0.f;
  if (tid < C1) bs[tid] = b[tid];
  if (tid == 0) {
    mbar_init(full + 0, 1); mbar_init(full + 1, 1);
    mbar_init(empty + 0, 4); mbar_init(empty + 1, 4);
    fence_barrier_init();
  }
  __syncthreads();
  const int n_my = (np - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 4) {
    // ---- prep warp: lane = row of the patch
    const uint4* src = reinterpret_cast<const uint4*>(patches);
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (n_my > 0) { a0 = __ldg(src + (size_t)blockIdx.x * 64 + lane * 2); a1 = __ldg(src + (size_t)blockIdx.x * 64 + lane * 2 + 1); }
    for (int it = 0; it < n_my; it++) {
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      if (it + 1 < n_my) {
        const size_t pn = (size_t)blockIdx.x + (size_t)(it + 1) * gridDim.x;
        n0 = __ldg(src + pn * 64 + lane * 2); n1 = __ldg(src + pn * 64 + lane * 2 + 1);
      }
      const uint32_t wd[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      unsigned s1 = 0, s2 = 0;
#pragma unroll
      for (int k = 0; k < 8; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) { const unsigned v = (wd[k] >> (8 * j)) & 255u; s1 += v; s2 += v * v; }
      for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      // torch.mean / torch.std (unbiased) of the 1024 pixels; (x-mean)/(std+1e-7) (desc_server.py:83-87)
      const double mean_d = (double)s1 / 1024.0;
      double var = ((double)s2 - (double)s1 * mean_d) / 1023.0;
      if (var < 0) var = 0;
      const float mean = (float)mean_d, sd = (float)sqrt(var) + 1e-7f;
      const int bf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(empty + bf, ph ^ 1);
      float* row = &P[bf][lane + 1][4];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float4 f;
        f.x = ((float)(wd[k] & 255u) - mean) / sd;
        f.y = ((float)((wd[k] >> 8) & 255u) - mean) / sd;
        f.z = ((float)((wd[k] >> 16) & 255u) - mean) / sd;
        f.w = ((float)(wd[k] >> 24) - mean) / sd;
        *reinterpret_cast<float4*>(row + 4 * k) = f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full + bf);
      a0 = n0; a1 = n1;
    }
  } else {
    // ---- conv warps: each thread keeps the 72 weights + 8 biases of ONE group of 8 output channels in registers and
    // walks over pixels: 9 shared-memory loads + 72 FMAs per output slot
    constexpr int C8 = C1 / 8, TPC = 128 / C8;
    const int c8 = tid / TPC, t0 = tid - c8 * TPC;
    float wr[8][9];
#pragma unroll
    for (int e = 0; e < 8; e++)
#pragma unroll
      for (int t = 0; t < 9; t++) wr[e][t] = __ldg(w + (c8 * 8 + e) * 9 + t);
    for (int it = 0; it < n_my; it++) {
      const int patch = (int)blockIdx.x + it * (int)gridDim.x;
      const int bf = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + bf, ph);
      __half* obase = out + ((size_t)c8 * out_slots + (size_t)FS + (size_t)patch * 1089) * 8;
#pragma unroll 1
      for (int p = t0; p < 1024; p += TPC) {
        const int y = p >> 5, x = p & 31;
        float v[9];
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
          for (int dx = 0; dx < 3; dx++) v[dy * 3 + dx] = P[bf][y + dy][x + dx + 3];
        const float4 b0 = *reinterpret_cast<const float4*>(bs + c8 * 8), b1 = *reinterpret_cast<const float4*>(bs + c8 * 8 + 4);
        float acc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        // every accumulator sees its taps in the order t = 0..8 (the contract with the oracle); the 8 chains are independent
#pragma unroll
        for (int t = 0; t < 9; t++)
#pragma unroll
          for (int e = 0; e < 8; e++) acc[e] = fmaf(wr[e][t], v[t], acc[e]);
        uint32_t h[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; e2++) h[e2] = relu_pack_h2(acc[2 * e2], acc[2 * e2 + 1]);
        *reinterpret_cast<uint4*>(obase + (size_t)((y + 1) * 33 + x) * 8) = make_uint4(h[0], h[1], h[2], h[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + bf);
    }
  }
}

// =================================================================================================
// conv1 (+ input normalisation) on the tensor pipe
//
// The first layer has ONE input channel: 9 multiply-adds per output value, 147k (C1 = 16) / 295k (C1 = 32) per patch, which
// on CUDA cores costs ~7k / ~14k warp instructions per patch (k_conv1, FMA pipe 43 % busy).  As a GEMM it is
// [1089 slots x K] * [K x C1] with K = 9 taps -- far too thin for fp16 operands to be exact enough on their own, so both
// operands are split: x = xh + xl, w = wh + wl (fp16 each) and K = 32 carries  xh*wh (9+1) | xl*wh (9+1) | xh*wl (9+1) | 0 (2);
// the dropped xl*wl term is 2^-22 relative, i.e. the result is fp32-accurate (products of fp16 pairs are exact in the fp32
// accumulator).  Per 128-slot M tile: 128 worker threads write one im2col row each (9 shared-memory loads, the hi / lo
// splits, 64 bytes out), one thread of a fifth warp issues TWO tcgen05.mma (K = 2 x 16), and the same 128 workers drain the
// previous tile's accumulator (bias + ReLU + fp16, conv1's map in the layout conv2 reads).  ~2.2k warp instructions per patch instead of 7-14k; what remains is
// the map's HBM write.  Results differ from k_conv1's in the last fp32 bits only (same values after the fp16 rounding of
// the map except where the sum sits on a rounding boundary).  Opt-in: see conv1_mma_enabled() for why it is not the default.
// =================================================================================================
template <int C1>
struct Conv1MmaCfg {
  static constexpr int PT = 33, PP = PT * PT, NT = (PP + 127) / 128;
  static constexpr int A_BYTES = 4 * 128 * 16;                 // [4 K-planes][128 rows][8 halves]
  static constexpr int W_BYTES = 4 * C1 * 16;                  // [4 K-planes][C1 rows][8 halves]
  // The chain  row stores -> mbarrier -> tcgen05.mma -> commit -> mbarrier -> tcgen05.ld  of ONE tile is ~1 us long however
  // small the MMA is, so the pipeline has to be deep: TS accumulators in TMEM (a worker drains tile k - (TS - 1) after it
  // built tile k) and TS + 1 operand stages.  With two accumulators and three stages the kernel was no faster than the
  // CUDA-core version, whatever the warp roles.
  static constexpr int TS = C1 == 16 ? 6 : 4, LAG = TS - 1;
  static constexpr int NSTAGE = TS + 1;
  static constexpr int TMEM_COLS = 128;          // >= TS * C1, a power of two
  static constexpr int SMEM_DYN = NSTAGE * A_BYTES;
};

template <int C1>
__global__ void __launch_bounds__(160, 3)
k_conv1_mma(const uint8_t* __restrict__ patches, int np, const __half* __restrict__ wsplit, const float* __restrict__ b,
            __half* __restrict__ out, size_t out_slots, const int* __restrict__ cnt_dev, int cnt_base) {
  using Cfg = Conv1MmaCfg<C1>;
  np = live_patches(np, cnt_dev, cnt_base);
  extern __shared__ __align__(1024) uint8_t a_s[];            // NSTAGE operand stages
  __shared__ __align__(128) uint8_t w_s[Cfg::W_BYTES];
  __shared__ __align__(16) float P[34][C1_PW_];               // normalised patch, pixel (y, x) at P[y + 1][x + 4], zero halo
  __shared__ __align__(16) float bs[C1];
  __shared__ unsigned red[2][4];
  __shared__ uint64_t bars[2 * Cfg::NSTAGE + 2 * Cfg::TS];
  __shared__ uint32_t tmem_slot;
  uint64_t* a_full = bars;                    // [NSTAGE] 4 worker warps
  uint64_t* a_empty = a_full + Cfg::NSTAGE;   // [NSTAGE] MMAs retired
  uint64_t* t_full = a_empty + Cfg::NSTAGE;   // [TS]
  uint64_t* t_empty = t_full + Cfg::TS;       // [TS] 4 worker warps
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  for (int i = tid; i < 34 * C1_PW_; i += 160) (&P[0][0])[i] = 0.f;
  for (int i = tid; i < Cfg::W_BYTES / 16; i += 160) reinterpret_cast<uint4*>(w_s)[i] = __ldg(reinterpret_cast<const uint4*>(wsplit) + i);
  if (tid < C1) bs[tid] = b[tid];
  fence_proxy_async();
  if (tid == 0) {
    for (int i = 0; i < Cfg::NSTAGE; i++) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < Cfg::TS; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int n_my = (int)blockIdx.x < np ? (np - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < 4) {
    // ---- workers (4 warps = the 128 rows of an M tile = the 128 TMEM lanes): per patch the normalisation, then per tile
    //      build row `tid` of tile k and drain the accumulator of tile k-1 -- no warp sits in a barrier wait while others work
    //      (a first version with dedicated epilogue warps spent three quarters of its issued instructions spinning)
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    auto drain = [&](int tc) {          // accumulator of running tile tc -> bias + ReLU + fp16 -> HBM
      const int ts = tc % Cfg::TS, tph = (tc / Cfg::TS) & 1;
      const size_t patch = (size_t)blockIdx.x + (size_t)(tc / Cfg::NT) * gridDim.x;
      const int t = tc % Cfg::NT;
      const int idx = t * 128 + tid;
      const int yy = idx / Cfg::PT, x = idx - yy * Cfg::PT;
      const bool valid = idx < Cfg::PP && yy >= 1 && x < 32;
      mbar_wait(t_full + ts, tph);
      fence_after_sync();
#pragma unroll
      for (int cc = 0; cc < C1 / 16; cc++) {
        float v[16];
        tmem_ld16(lane_base + ts * C1 + cc * 16, v);
        if (valid) {
          uint32_t h[8];
#pragma unroll
          for (int e = 0; e < 8; e++) h[e] = relu_pack_h2(v[2 * e] + bs[cc * 16 + 2 * e], v[2 * e + 1] + bs[cc * 16 + 2 * e + 1]);
          __half* o = out + ((size_t)(cc * 2) * out_slots + (size_t)FS + patch * Cfg::PP + idx) * 8;
          *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(o + out_slots * 8) = make_uint4(h[4], h[5], h[6], h[7]);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + ts);
    };
    int tc = 0;
    uint2 px = make_uint2(0, 0);
    if (n_my > 0) px = __ldg(reinterpret_cast<const uint2*>(patches + (size_t)blockIdx.x * 1024) + tid);
    for (int it = 0; it < n_my; it++) {
      const size_t patch = (size_t)blockIdx.x + (size_t)it * gridDim.x;
      const uint32_t wd[2] = {px.x, px.y};                    // 8 pixels of row tid / 4
      if (it + 1 < n_my) px = __ldg(reinterpret_cast<const uint2*>(patches + (patch + gridDim.x) * 1024) + tid);   // next patch in flight
      unsigned s1 = 0, s2 = 0;
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) { const unsigned v = (wd[k] >> (8 * j)) & 255u; s1 += v; s2 += v * v; }
      for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
      // every worker is done with the previous patch's P (and red[]) before either is overwritten
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      unsigned t1 = 0, t2 = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) { t1 += red[0][k]; t2 += red[1][k]; }
      // torch.mean / torch.std (unbiased) of the 1024 pixels; (x-mean)/(std+1e-7) (desc_server.py:83-87), as k_conv1
      const double mean_d = (double)t1 / 1024.0;
      double var = ((double)t2 - (double)t1 * mean_d) / 1023.0;
      if (var < 0) var = 0;
      const float mean = (float)mean_d, sd = (float)sqrt(var) + 1e-7f;
      {
        const int y = tid >> 2, x0 = (tid & 3) * 8;
        float* row = &P[y + 1][4 + x0];
#pragma unroll
        for (int k = 0; k < 2; k++) {
          float4 f;
          f.x = ((float)(wd[k] & 255u) - mean) / sd;
          f.y = ((float)((wd[k] >> 8) & 255u) - mean) / sd;
          f.z = ((float)((wd[k] >> 16) & 255u) - mean) / sd;
          f.w = ((float)(wd[k] >> 24) - mean) / sd;
          *reinterpret_cast<float4*>(row + 4 * k) = f;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int t = 0; t < Cfg::NT; t++, tc++) {
        const int idx = t * 128 + tid;
        const int yy = idx / Cfg::PT, x = idx - yy * Cfg::PT, y = yy - 1;
        const bool valid = idx < Cfg::PP && yy >= 1 && x < 32;
        __half2 k2[16];
        if (valid) {
          float v[9];
#pragma unroll
          for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) v[dy * 3 + dx] = P[y + dy][x + dx + 3];
          // K order (pairs, so that every conversion is a packed F2FP): xh0..7 | xh8 0 | xl0..7 | xl8 0 | xh0..7 | xh8 0 | 0 0
          __half2 H[5], L[5];
#pragma unroll
          for (int k = 0; k < 5; k++) {
            const float a0 = v[2 * k], a1 = k < 4 ? v[2 * k + 1] : 0.f;
            H[k] = __floats2half2_rn(a0, a1);
            const float2 back = __half22float2(H[k]);
            L[k] = __floats2half2_rn(a0 - back.x, a1 - back.y);
          }
#pragma unroll
          for (int k = 0; k < 5; k++) { k2[k] = H[k]; k2[5 + k] = L[k]; k2[10 + k] = H[k]; }
          k2[15] = __float2half2_rn(0.f);
        } else {
#pragma unroll
          for (int k = 0; k < 16; k++) k2[k] = __float2half2_rn(0.f);
        }
        const int s = tc % Cfg::NSTAGE, ph = (tc / Cfg::NSTAGE) & 1;
        mbar_wait(a_empty + s, ph ^ 1);
        uint8_t* dst = a_s + s * Cfg::A_BYTES + tid * 16;
#pragma unroll
        for (int pl = 0; pl < 4; pl++)
          *reinterpret_cast<uint4*>(dst + pl * 2048) = make_uint4(*reinterpret_cast<uint32_t*>(&k2[4 * pl]), *reinterpret_cast<uint32_t*>(&k2[4 * pl + 1]),
                                                                  *reinterpret_cast<uint32_t*>(&k2[4 * pl + 2]), *reinterpret_cast<uint32_t*>(&k2[4 * pl + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + s);
        if (tc >= Cfg::LAG) drain(tc - Cfg::LAG);
      }
    }
    for (int k = tc > Cfg::LAG ? tc - Cfg::LAG : 0; k < tc; k++) drain(k);
  } else {
    // ---- MMA issuer: two K = 16 steps per tile
    constexpr uint32_t idesc = instr_desc_f16(C1);
    const uint64_t a_desc0 = smem_desc(smem_u32(a_s), 2048, 128);
    const uint64_t w_desc0 = smem_desc(smem_u32(w_s), C1 * 16, 128);
    int tc = 0;
    for (int it = 0; it < n_my; it++)
#pragma unroll 1
      for (int t = 0; t < Cfg::NT; t++, tc++) {
        const int ts = tc % Cfg::TS, tph = (tc / Cfg::TS) & 1;
        const int s = tc % Cfg::NSTAGE, ph = (tc / Cfg::NSTAGE) & 1;
        while (!mbar_try_wait(t_empty + ts, tph ^ 1)) __nanosleep(20);
        while (!mbar_try_wait(a_full + s, ph)) __nanosleep(20);
        fence_after_sync();
        const uint64_t a_tile = a_desc0 + (uint64_t)((s * Cfg::A_BYTES) >> 4);
        const uint32_t d_tmem = tmem_base + ts * C1;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ks++)
            mma_f16(d_tmem, a_tile + (uint64_t)(2 * ks * 128), w_desc0 + (uint64_t)(2 * ks * C1), idesc, ks != 0);
          mma_commit(a_empty + s);
          mma_commit(t_full + ts);
        }
        __syncwarp();
      }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =================================================================================================
// conv2..conv6: tcgen05 implicit GEMM
// =================================================================================================
template <int CIN, int COUT_T, int NSPLIT, int S, int NGRP, int OUT_MODE>
struct ConvCfg {
  static constexpr int PT = S + 1, PP = PT * PT;
  static constexpr int C8 = CIN / 8, KSTEPS = CIN / 16;
  static constexpr int COUT = COUT_T * NSPLIT;
  static constexpr int HALO_LO = PT + 1, HALO_HI = (NGRP == 1) ? PT + 1 : 0;
  static constexpr int TP = 128 + HALO_LO + HALO_HI;
  static constexpr int NPL = NGRP * C8;
  static constexpr int A_BYTES = NPL * TP * 16;
  static constexpr int W_BYTES = 9 * KSTEPS * 2 * COUT_T * 16;
  static constexpr int TMEM_COLS = (2 * COUT_T <= 32) ? 32 : (2 * COUT_T <= 64 ? 64 : (2 * COUT_T <= 128 ? 128 : 256));
  // A-tile ring: as deep as shared memory allows (small tiles are latency- not bandwidth-limited).  Layers whose weights
  // and >= 3 A stages fit in half an SM's shared memory run TWO CTAs per SM: two independent load/MMA/epilogue
  // pipelines interleave on the tensor pipe and hide each other's per-tile latencies.
  static constexpr int NST_HALF = (113 * 1024 - W_BYTES - 256) / A_BYTES;
  static constexpr bool TWO_CTAS = NST_HALF >= 3;
  static constexpr int NST_FIT = TWO_CTAS ? NST_HALF : (232448 - W_BYTES - 256) / A_BYTES;
  static constexpr int NST = NST_FIT < 2 ? 2 : (NST_FIT > 8 ? 8 : NST_FIT);
  static constexpr int SMEM_BYTES = W_BYTES + NST * A_BYTES + 256;
  static_assert(CIN % 16 == 0 && COUT_T % 16 == 0, "UMMA shape");
};

template <int CIN, int COUT_T, int NSPLIT, int S, int NGRP, int OUT_MODE>
__global__ void __launch_bounds__(192, 1)
k_conv_umma(const __half* __restrict__ in, size_t in_slots, const __half* __restrict__ wts,
            const float* __restrict__ bias, __half* __restrict__ out, size_t out_slots, int np, int ntiles,
            int patch_base, int contig, int nst, const int* __restrict__ cnt_dev, int cnt_base) {
  using Cfg = ConvCfg<CIN, COUT_T, NSPLIT, S, NGRP, OUT_MODE>;
  if (cnt_dev != nullptr) {
    np = live_patches(np, cnt_dev, cnt_base);
    ntiles = (np * Cfg::PP + 127) / 128;
  }
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_s = smem;
  // M tiles of this CTA.  contig: one balanced run of consecutive tiles (the halo rows a tile shares with its predecessor
  // were fetched by this very SM one tile earlier); strided: tile = blockIdx.x + i * gridDim.x.  Measured equal on B200
  // (bench 241.3 vs 241.1 pairs/s; DRAM reads are at the algorithmic size either way: the halo always comes from L2).
  int tile_first, tile_end, tile_step;
  if (contig) {
    const int base = ntiles / (int)gridDim.x, rem = ntiles % (int)gridDim.x, b = (int)blockIdx.x;
    tile_first = b * base + (b < rem ? b : rem);
    tile_end = tile_first + base + (b < rem ? 1 : 0);
    tile_step = 1;
  } else {
    tile_first = blockIdx.x; tile_end = ntiles; tile_step = gridDim.x;
  }
  uint8_t* a_s = smem + Cfg::W_BYTES;
  // nst <= Cfg::NST stages of the A ring are in use (the launch sizes the dynamic shared memory for nst: a shallower ring
  // leaves room for CTAs of other streams' kernels on the same SM)
  const int NST = nst;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::W_BYTES + NST * Cfg::A_BYTES);
  uint64_t* w_full = bars;                  // weights landed
  uint64_t* a_full = bars + 1;              // [NST]
  uint64_t* a_empty = bars + 1 + NST;       // [NST]
  uint64_t* t_full = bars + 1 + 2 * NST;    // [2]
  uint64_t* t_empty = bars + 3 + 2 * NST;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * NST);

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps descriptors,
  // barrier addresses and loop counters in uniform registers (otherwise every tcgen05.mma / bulk copy is wrapped
  // in an R2UR + ELECT + branch "waterfall", measured at ~100-146 cycles per MMA instead of the 8-32 cycle floor)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nsp = blockIdx.y;  // which COUT_T slice of the output channels

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < NST; i++) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer (whole warp walks the loop, one elected lane issues): weights once, then one A tile (NPL planes) per M tile
    if (elect_one()) {
      mbar_expect_tx(w_full, Cfg::W_BYTES);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(wts) + (size_t)nsp * Cfg::W_BYTES;
      for (int off = 0; off < Cfg::W_BYTES; off += 32768) {
        int bytes = Cfg::W_BYTES - off < 32768 ? Cfg::W_BYTES - off : 32768;
        bulk_g2s(w_s + off, wsrc + off, bytes, w_full);
      }
    }
    __syncwarp();
    int s = 0, ph = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_step) {
      mbar_wait(a_empty + s, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(a_full + s, Cfg::A_BYTES);
        const size_t slot0 = (size_t)FS + (size_t)tile * 128 - Cfg::HALO_LO;
        uint8_t* dst = a_s + s * Cfg::A_BYTES;
#pragma unroll 1
        for (int pl = 0; pl < Cfg::NPL; pl++)
          bulk_g2s(dst + pl * Cfg::TP * 16, in + ((size_t)pl * in_slots + slot0) * 8, Cfg::TP * 16, a_full + s);
      }
      __syncwarp();
      if (++s == NST) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the warp waits together, one elected lane issues the 9 x KSTEPS MMAs of the tile back to back
    constexpr uint32_t idesc = instr_desc_f16(COUT_T);
    mbar_wait(w_full, 0);
    const uint32_t w_addr = smem_u32(w_s);
    const uint64_t a_desc0 = smem_desc(smem_u32(a_s), Cfg::TP * 16, 128);
    const uint64_t b_desc0 = smem_desc(w_addr, COUT_T * 16, 128);
    int it = 0, s = 0, ph = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_step, it++) {
      const int ts = it & 1, tph = (it >> 1) & 1;
      mbar_wait(t_empty + ts, tph ^ 1);
      mbar_wait(a_full + s, ph);
      fence_after_sync();
      // descriptors differ from tile/tap/k-step only in the 16-byte start-address field (bits 0-13): plain adds
      const uint64_t a_tile = a_desc0 + (uint64_t)((s * Cfg::A_BYTES) >> 4);
      const uint32_t d_tmem = tmem_base + ts * COUT_T;
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 9; t++) {
          const int dy = t / 3, dx = t % 3;
          int grp, shift;
          if (NGRP == 1) { grp = 0; shift = (dy - 1) * Cfg::PT + (dx - 1); }
          else {
            grp = ((dy == 1) ? 0 : 2) + ((dx == 1) ? 0 : 1);
            shift = ((dy == 0) ? -Cfg::PT : 0) + ((dx == 0) ? -1 : 0);
          }
#pragma unroll
          for (int ks = 0; ks < Cfg::KSTEPS; ks++) {
            const int a_off16 = (grp * Cfg::C8 + 2 * ks) * Cfg::TP + Cfg::HALO_LO + shift;     // >= 0
            const int b_off16 = (t * Cfg::KSTEPS + ks) * 2 * COUT_T;
            mma_f16(d_tmem, a_tile + (uint64_t)a_off16, b_desc0 + (uint64_t)b_off16, idesc, (t | ks) != 0);
          }
        }
        mma_commit(a_empty + s);   // smem tile may be refilled once these MMAs retire
        mma_commit(t_full + ts);   // accumulator ready for the epilogue
      }
      __syncwarp();
      if (++s == NST) { s = 0; ph ^= 1; }
    }
  } else {
    // ---- epilogue warps 2..5: TMEM -> bias + ReLU -> fp16 -> next layer's layout
    const int q = warp & 3;               // TMEM lane quarter this warp may touch
    const int m = q * 32 + lane;          // row of the M tile
    float br[COUT_T];
#pragma unroll
    for (int e = 0; e < COUT_T; e++) br[e] = __ldg(bias + nsp * COUT_T + e);
    int it = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_step, it++) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int g = tile * 128 + m;
      const int patch = g / Cfg::PP;
      const int idx = g - patch * Cfg::PP;
      const int yy = idx / Cfg::PT, x = idx - yy * Cfg::PT, y = yy - 1;
      const bool valid = patch < np && yy >= 1 && x < S;
      size_t oslot; int oplane0;
      if (OUT_MODE == OUT_NORMAL) { oslot = (size_t)FS + g; oplane0 = 0; }
      else if (OUT_MODE == OUT_PARITY) {
        constexpr int PT2 = S / 2 + 1, PP2 = PT2 * PT2;
        oslot = (size_t)FS + (size_t)patch * PP2 + ((y >> 1) + 1) * PT2 + (x >> 1);
        oplane0 = (((y & 1) << 1) | (x & 1)) * (Cfg::COUT / 8);
      } else {
        oslot = (size_t)(patch_base + patch);
        oplane0 = (y * S + x) * (Cfg::COUT / 8);
      }
      oplane0 += nsp * (COUT_T / 8);
      mbar_wait(t_full + s, ph);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * COUT_T;
#pragma unroll
      for (int cc = 0; cc < COUT_T / 16; cc++) {
        float v[16];
        tmem_ld16(taddr + cc * 16, v);
        if (valid) {
          uint32_t h[8];
#pragma unroll
          for (int e = 0; e < 8; e++) h[e] = relu_pack_h2(v[2 * e] + br[cc * 16 + 2 * e], v[2 * e + 1] + br[cc * 16 + 2 * e + 1]);
          __half* o = out + ((size_t)(oplane0 + cc * 2) * out_slots + oslot) * 8;
          *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(o + out_slots * 8) = make_uint4(h[4], h[5], h[6], h[7]);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + s);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =================================================================================================
// conv2 .. conv4 (C1 = 16: AffNet / OriNet) or conv2 .. conv3 (C1 = 32: HardNet++) in ONE kernel: the activation maps
// between the fused layers never leave the SM.
//
// Unfused, every layer round-trips its map through HBM (HardNet conv2 alone: 64 KB in + 64 KB out per patch) and every
// launch pays its own prologue, weight load and tail wave.  Here a CTA walks over whole PATCHES:
//   conv2  (C1 -> C1 @ 32x32)    A operand streamed from conv1's map in HBM through a ring of patch-aligned M tiles
//                                (128 slots + halo per plane, bulk copies), exactly as k_conv_umma does; the epilogue
//                                writes bias + ReLU + fp16 into the PARITY planes of a shared-memory map A2
//   conv3  (C1 -> 2C1, stride 2) A operand = A2 in shared memory (four parity groups, unit-stride tap shifts);
//                                epilogue -> shared-memory map A3 (C1 = 16) or HBM (C1 = 32, last fused layer)
//   conv4  (2C1 -> 2C1 @ 16x16)  A operand = A3; epilogue -> HBM in the parity layout conv5 reads          (C1 = 16)
// Roles as in k_conv_umma: warp 0 bulk-copy producer, warp 1 issues every MMA of every layer, warps 2-5 are the epilogue
// (TMEM lane quarters).  Accumulators are double buffered in TMEM across ALL tiles of all layers, so the MMAs of tile k+1
// overlap the epilogue of tile k; only at a layer boundary must the MMA warp wait for the map it is about to read
// (mbarrier a2_full / a3_full: 4 epilogue warps x tiles arrivals, after a generic->async proxy fence).  The epilogue of a
// later layer is ordered after the MMAs that read the map it overwrites by the accumulator barriers themselves
// (tcgen05.commit covers every MMA issued before it), so no "empty" barrier exists for A2 / A3.
// Arithmetic is the unfused path's, tap by tap and k-step by k-step, with the same fp16 rounding points: outputs are
// bit-identical to k_conv_umma's (tests/test_gpu_parity.py::test_fused_trunk_bit_identical).
// =================================================================================================
template <int C1, int NL>
struct TrunkCfg {
  static constexpr int PT1 = 33, PP1 = PT1 * PT1, NT1 = (PP1 + 127) / 128, HALO1 = PT1 + 1, TP1 = 128 + 2 * HALO1;
  static constexpr int PT2 = 17, PP2 = PT2 * PT2, NT2 = (PP2 + 127) / 128, HALO2 = PT2 + 1;
  static constexpr int CO0 = C1, CO1 = 2 * C1, CO2 = 2 * C1;                 // output channels of conv2 / conv3 / conv4
  static constexpr int KS0 = C1 / 16, KS1 = C1 / 16, KS2 = 2 * C1 / 16;      // k-steps (input channels / 16)
  static constexpr int W0_BYTES = 9 * KS0 * 2 * CO0 * 16, W1_BYTES = 9 * KS1 * 2 * CO1 * 16;
  static constexpr int W2_BYTES = NL == 3 ? 9 * KS2 * 2 * CO2 * 16 : 0;
  static constexpr int NPL1 = C1 / 8, A1_BYTES = NPL1 * TP1 * 16;            // one ring stage of conv2's input
  static constexpr int PS2 = HALO2 + PP2, NPL2 = 4 * (C1 / 8);               // A2: parity planes, plane stride in slots
  static constexpr int A2_BYTES = NPL2 * PS2 * 16 + (NT2 * 128 - PP2) * 16;  // + the overrun of the last tile's garbage rows
  static constexpr int PS3 = HALO2 + PP2 + HALO2, NPL3 = 2 * C1 / 8;         // A3: normal planes with both halos
  static constexpr int A3_BYTES = NL == 3 ? NPL3 * PS3 * 16 + (NT2 * 128 - PP2) * 16 : 0;
  static constexpr int BIAS_BYTES = (CO0 + CO1 + CO2) * 4;
  static constexpr int NST = C1 == 16 ? 3 : 4;
  static constexpr int OFF_W0 = 0, OFF_W1 = OFF_W0 + W0_BYTES, OFF_W2 = OFF_W1 + W1_BYTES, OFF_RING = OFF_W2 + W2_BYTES;
  static constexpr int OFF_A2 = OFF_RING + NST * A1_BYTES, OFF_A3 = OFF_A2 + A2_BYTES, OFF_BIAS = OFF_A3 + A3_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int NMAX = 2 * C1;                                        // widest accumulator
  static constexpr int TMEM_COLS = 2 * NMAX <= 64 ? 64 : 128;
  static_assert(OFF_RING % 16 == 0 && OFF_A2 % 16 == 0 && OFF_A3 % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
};

template <int C1, int NL>
__global__ void __launch_bounds__(192, (C1 == 16 ? 2 : 1))
k_trunk(const __half* __restrict__ in, size_t in_slots, const __half* __restrict__ w0, const float* __restrict__ b0,
        const __half* __restrict__ w1, const float* __restrict__ b1, const __half* __restrict__ w2, const float* __restrict__ b2,
        __half* __restrict__ out, size_t out_slots, int np, const int* __restrict__ cnt_dev, int cnt_base) {
  using Cfg = TrunkCfg<C1, NL>;
  np = live_patches(np, cnt_dev, cnt_base);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint8_t* a2_s = smem + Cfg::OFF_A2;
  uint8_t* a3_s = smem + Cfg::OFF_A3;
  float* bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* w_full = bars;                      // the three weight blocks landed
  uint64_t* a_full = bars + 1;                  // [NST] ring stage landed
  uint64_t* a_empty = a_full + Cfg::NST;        // [NST] MMAs reading the stage retired
  uint64_t* t_full = a_empty + Cfg::NST;        // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;               // [2] accumulator drained (4 epilogue warps)
  uint64_t* a2_full = t_empty + 2;              // conv2's map complete in shared memory (4 warps x NT1 tiles per patch)
  uint64_t* a3_full = a2_full + 1;              // conv3's map complete (4 warps x NT2 tiles per patch)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a3_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int n_my = (int)blockIdx.x < np ? (np - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // the resident maps start as zeros: their pad slots are never written afterwards
  for (int i = threadIdx.x; i < (Cfg::A2_BYTES + Cfg::A3_BYTES) / 16; i += 192) reinterpret_cast<uint4*>(a2_s)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < Cfg::CO0 + Cfg::CO1 + (NL == 3 ? Cfg::CO2 : 0); i += 192)
    bias_s[i] = i < Cfg::CO0 ? __ldg(b0 + i) : (i < Cfg::CO0 + Cfg::CO1 ? __ldg(b1 + i - Cfg::CO0) : __ldg(b2 + i - Cfg::CO0 - Cfg::CO1));
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < Cfg::NST; i++) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    mbar_init(a2_full, 4 * Cfg::NT1);
    mbar_init(a3_full, 4 * Cfg::NT2);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer: weights once, then conv2's input tiles of every patch of this CTA (patch-aligned M tiles)
    if (elect_one()) {
      mbar_expect_tx(w_full, Cfg::W0_BYTES + Cfg::W1_BYTES + Cfg::W2_BYTES);
      bulk_g2s(smem + Cfg::OFF_W0, w0, Cfg::W0_BYTES, w_full);
      for (int off = 0; off < Cfg::W1_BYTES; off += 32768)
        bulk_g2s(smem + Cfg::OFF_W1 + off, reinterpret_cast<const uint8_t*>(w1) + off, min(32768, Cfg::W1_BYTES - off), w_full);
      if (NL == 3)
        for (int off = 0; off < Cfg::W2_BYTES; off += 32768)
          bulk_g2s(smem + Cfg::OFF_W2 + off, reinterpret_cast<const uint8_t*>(w2) + off, min(32768, Cfg::W2_BYTES - off), w_full);
    }
    __syncwarp();
    int s = 0, ph = 0;
    for (int it = 0; it < n_my; it++) {
      const size_t patch = (size_t)blockIdx.x + (size_t)it * gridDim.x;
      for (int t = 0; t < Cfg::NT1; t++) {
        mbar_wait(a_empty + s, ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(a_full + s, Cfg::A1_BYTES);
          const size_t slot0 = (size_t)FS + patch * Cfg::PP1 + (size_t)t * 128 - Cfg::HALO1;
          uint8_t* dst = ring + s * Cfg::A1_BYTES;
#pragma unroll 1
          for (int pl = 0; pl < Cfg::NPL1; pl++)
            bulk_g2s(dst + pl * Cfg::TP1 * 16, in + ((size_t)pl * in_slots + slot0) * 8, Cfg::TP1 * 16, a_full + s);
        }
        __syncwarp();
        if (++s == Cfg::NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: per patch conv2's NT1 tiles from the ring, then conv3's (and conv4's) NT2 tiles from the resident maps
    constexpr uint32_t idesc0 = instr_desc_f16(Cfg::CO0), idesc1 = instr_desc_f16(Cfg::CO1), idesc2 = instr_desc_f16(Cfg::CO2);
    mbar_wait(w_full, 0);
    const uint64_t ring_desc = smem_desc(smem_u32(ring), Cfg::TP1 * 16, 128);
    const uint64_t a2_desc = smem_desc(smem_u32(a2_s), Cfg::PS2 * 16, 128);
    const uint64_t a3_desc = smem_desc(smem_u32(a3_s), Cfg::PS3 * 16, 128);
    const uint64_t w0_desc = smem_desc(smem_u32(smem + Cfg::OFF_W0), Cfg::CO0 * 16, 128);
    const uint64_t w1_desc = smem_desc(smem_u32(smem + Cfg::OFF_W1), Cfg::CO1 * 16, 128);
    const uint64_t w2_desc = smem_desc(smem_u32(smem + Cfg::OFF_W2), Cfg::CO2 * 16, 128);
    int s = 0, ph = 0, tc = 0;      // ring stage / phase, running accumulator-tile counter
    for (int it = 0; it < n_my; it++) {
      // conv2
      for (int t = 0; t < Cfg::NT1; t++, tc++) {
        const int ts = tc & 1, tph = (tc >> 1) & 1;
        mbar_wait(t_empty + ts, tph ^ 1);
        mbar_wait(a_full + s, ph);
        fence_after_sync();
        const uint64_t a_tile = ring_desc + (uint64_t)((s * Cfg::A1_BYTES) >> 4);
        const uint32_t d_tmem = tmem_base + ts * Cfg::NMAX;
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const int shift = (tap / 3 - 1) * Cfg::PT1 + (tap % 3 - 1);
#pragma unroll
            for (int ks = 0; ks < Cfg::KS0; ks++)
              mma_f16(d_tmem, a_tile + (uint64_t)((2 * ks) * Cfg::TP1 + Cfg::HALO1 + shift),
                      w0_desc + (uint64_t)((tap * Cfg::KS0 + ks) * 2 * Cfg::CO0), idesc0, (tap | ks) != 0);
          }
          mma_commit(a_empty + s);
          mma_commit(t_full + ts);
        }
        __syncwarp();
        if (++s == Cfg::NST) { s = 0; ph ^= 1; }
      }
      // conv3 (stride 2): four parity groups of A2
      mbar_wait(a2_full, it & 1);
      fence_after_sync();
      for (int t = 0; t < Cfg::NT2; t++, tc++) {
        const int ts = tc & 1, tph = (tc >> 1) & 1;
        mbar_wait(t_empty + ts, tph ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + ts * Cfg::NMAX;
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const int dy = tap / 3, dx = tap % 3;
            const int grp = ((dy == 1) ? 0 : 2) + ((dx == 1) ? 0 : 1);
            const int shift = ((dy == 0) ? -Cfg::PT2 : 0) + ((dx == 0) ? -1 : 0);
#pragma unroll
            for (int ks = 0; ks < Cfg::KS1; ks++)
              mma_f16(d_tmem, a2_desc + (uint64_t)((grp * (C1 / 8) + 2 * ks) * Cfg::PS2 + Cfg::HALO2 + shift + t * 128),
                      w1_desc + (uint64_t)((tap * Cfg::KS1 + ks) * 2 * Cfg::CO1), idesc1, (tap | ks) != 0);
          }
          mma_commit(t_full + ts);
        }
        __syncwarp();
      }
      if (NL == 3) {
        // conv4: A3 with both halos
        mbar_wait(a3_full, it & 1);
        fence_after_sync();
        for (int t = 0; t < Cfg::NT2; t++, tc++) {
          const int ts = tc & 1, tph = (tc >> 1) & 1;
          mbar_wait(t_empty + ts, tph ^ 1);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + ts * Cfg::NMAX;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
              const int shift = (tap / 3 - 1) * Cfg::PT2 + (tap % 3 - 1);
#pragma unroll
              for (int ks = 0; ks < Cfg::KS2; ks++)
                mma_f16(d_tmem, a3_desc + (uint64_t)((2 * ks) * Cfg::PS3 + Cfg::HALO2 + shift + t * 128),
                        w2_desc + (uint64_t)((tap * Cfg::KS2 + ks) * 2 * Cfg::CO2), idesc2, (tap | ks) != 0);
            }
            mma_commit(t_full + ts);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue warps 2..5: the same tile sequence; TMEM -> bias + ReLU -> fp16 -> the next layer's operand
    const int q = warp & 3, m = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int tc = 0;
    for (int it = 0; it < n_my; it++) {
      const size_t patch = (size_t)blockIdx.x + (size_t)it * gridDim.x;
      // conv2 -> A2 (parity planes in shared memory)
      for (int t = 0; t < Cfg::NT1; t++, tc++) {
        const int ts = tc & 1, tph = (tc >> 1) & 1;
        const int idx = t * 128 + m;
        const int yy = idx / Cfg::PT1, x = idx - yy * Cfg::PT1, y = yy - 1;
        const bool valid = idx < Cfg::PP1 && yy >= 1 && x < 32;
        const int oslot = ((y >> 1) + 1) * Cfg::PT2 + (x >> 1);
        const int oplane0 = (((y & 1) << 1) | (x & 1)) * (Cfg::CO0 / 8);
        mbar_wait(t_full + ts, tph);
        fence_after_sync();
#pragma unroll
        for (int cc = 0; cc < Cfg::CO0 / 16; cc++) {
          float v[16];
          tmem_ld16(lane_base + ts * Cfg::NMAX + cc * 16, v);
          if (valid) {
            uint32_t h[8];
#pragma unroll
            for (int e = 0; e < 8; e++) h[e] = relu_pack_h2(v[2 * e] + bias_s[cc * 16 + 2 * e], v[2 * e + 1] + bias_s[cc * 16 + 2 * e + 1]);
            uint8_t* o = a2_s + ((size_t)(oplane0 + cc * 2) * Cfg::PS2 + Cfg::HALO2 + oslot) * 16;
            *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(o + Cfg::PS2 * 16) = make_uint4(h[4], h[5], h[6], h[7]);
          }
        }
        fence_before_sync();
        fence_proxy_async();          // the stores above are read by tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(t_empty + ts); mbar_arrive(a2_full); }
      }
      // conv3 -> A3 (shared memory, C1 = 16) or HBM (C1 = 32)
      for (int t = 0; t < Cfg::NT2; t++, tc++) {
        const int ts = tc & 1, tph = (tc >> 1) & 1;
        const int idx = t * 128 + m;
        const int yy = idx / Cfg::PT2, x = idx - yy * Cfg::PT2;
        const bool valid = idx < Cfg::PP2 && yy >= 1 && x < 16;
        mbar_wait(t_full + ts, tph);
        fence_after_sync();
#pragma unroll
        for (int cc = 0; cc < Cfg::CO1 / 16; cc++) {
          float v[16];
          tmem_ld16(lane_base + ts * Cfg::NMAX + cc * 16, v);
          if (valid) {
            uint32_t h[8];
#pragma unroll
            for (int e = 0; e < 8; e++)
              h[e] = relu_pack_h2(v[2 * e] + bias_s[Cfg::CO0 + cc * 16 + 2 * e], v[2 * e + 1] + bias_s[Cfg::CO0 + cc * 16 + 2 * e + 1]);
            if (NL == 3) {
              uint8_t* o = a3_s + ((size_t)(cc * 2) * Cfg::PS3 + Cfg::HALO2 + idx) * 16;
              *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(o + Cfg::PS3 * 16) = make_uint4(h[4], h[5], h[6], h[7]);
            } else {
              __half* o = out + ((size_t)(cc * 2) * out_slots + (size_t)FS + patch * Cfg::PP2 + idx) * 8;
              *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(o + out_slots * 8) = make_uint4(h[4], h[5], h[6], h[7]);
            }
          }
        }
        fence_before_sync();
        if (NL == 3) fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(t_empty + ts); if (NL == 3) mbar_arrive(a3_full); }
      }
      if (NL == 3) {
        // conv4 -> HBM, parity layout of the 8x8 stride-2 layer that follows (k_conv_umma OUT_PARITY with S = 16)
        for (int t = 0; t < Cfg::NT2; t++, tc++) {
          const int ts = tc & 1, tph = (tc >> 1) & 1;
          const int idx = t * 128 + m;
          const int yy = idx / Cfg::PT2, x = idx - yy * Cfg::PT2, y = yy - 1;
          const bool valid = idx < Cfg::PP2 && yy >= 1 && x < 16;
          const size_t oslot = (size_t)FS + patch * 81 + ((y >> 1) + 1) * 9 + (x >> 1);
          const int oplane0 = (((y & 1) << 1) | (x & 1)) * (Cfg::CO2 / 8);
          mbar_wait(t_full + ts, tph);
          fence_after_sync();
#pragma unroll
          for (int cc = 0; cc < Cfg::CO2 / 16; cc++) {
            float v[16];
            tmem_ld16(lane_base + ts * Cfg::NMAX + cc * 16, v);
            if (valid) {
              uint32_t h[8];
#pragma unroll
              for (int e = 0; e < 8; e++)
                h[e] = relu_pack_h2(v[2 * e] + bias_s[Cfg::CO0 + Cfg::CO1 + cc * 16 + 2 * e], v[2 * e + 1] + bias_s[Cfg::CO0 + Cfg::CO1 + cc * 16 + 2 * e + 1]);
              __half* o = out + ((size_t)(oplane0 + cc * 2) * out_slots + oslot) * 8;
              *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(o + out_slots * 8) = make_uint4(h[4], h[5], h[6], h[7]);
            }
          }
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + ts);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =================================================================================================
// conv1 + conv2 fused: the 1 -> C1 first layer never leaves the SM.
//
//   loader (1 warp)      : bulk copies of the normalised pixels a tile touches, 8 tiles ahead
//   producers (8 warps)  : conv1 on CUDA cores (fp32, the very FMA chain of k_conv1) for the tile's 196 activation
//                          slots -> bias + ReLU -> fp16 -> A tile in shared memory, in the [C1/8 planes][196 rows][8]
//                          layout k_conv_umma would have loaded from HBM
//   MMA warp             : D = sum over 9 taps (row-shifted views of the A tile) * W2      conv2 on tcgen05
//   epilogue (4 warps)   : D + bias -> ReLU -> fp16 -> HBM in the next layer's layout
//
// Only HBM traffic left: 4 KB of normalised pixels in, the conv2 map out (conv1's 32-64 KB map per patch is gone).
// A first version ran conv1 on the tensor pipe as well (im2col rows, K = 32 with split hi/lo operands): with N = 16-32
// every MMA re-reads 4 KB of A from shared memory, the pipe saturates the shared-memory port and the generic
// LDS/STS of producers and epilogues starve (970-1740 busy cycles per tile in the accumulator->A-tile epilogue,
// profiles/r01_conv12_stalls.txt); CUDA-core conv1 removes 28 KB of that traffic per tile.
// =================================================================================================
// per patch: mean / unbiased std of the 1024 pixels, then v = (px - mean) / (std + 1e-7) (desc_server.py:83-87, as k_conv1)
__global__ void __launch_bounds__(256)
k_patch_prep(const uint8_t* __restrict__ patches, int np, float* __restrict__ norm) {
  __shared__ unsigned red[2][8];
  __shared__ float s_mean, s_sd;
  const int patch = blockIdx.x, tid = threadIdx.x;
  const uchar4 px = reinterpret_cast<const uchar4*>(patches + (size_t)patch * 1024)[tid];
  unsigned s1 = px.x + px.y + px.z + px.w;
  unsigned s2 = px.x * px.x + px.y * px.y + px.z * px.z + px.w * px.w;
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    unsigned a = 0, q = 0;
    for (int i = 0; i < 8; i++) { a += red[0][i]; q += red[1][i]; }
    const double mean = (double)a / 1024.0;
    double var = ((double)q - (double)a * mean) / 1023.0;
    if (var < 0) var = 0;
    s_mean = (float)mean;
    s_sd = (float)sqrt(var) + 1e-7f;
  }
  __syncthreads();
  const float mean = s_mean, sd = s_sd;
  reinterpret_cast<float4*>(norm + (size_t)patch * 1024)[tid] =
      make_float4(((float)px.x - mean) / sd, ((float)px.y - mean) / sd, ((float)px.z - mean) / sd, ((float)px.w - mean) / sd);
}

// conv1 weights of the three nets for k_conv12: [C1][9] weights then [C1] biases, read as FFMA constant operands
// (uniform across the warp: every lane applies the same weight to its own pixel)
__constant__ float c_conv1[3][32 * 9 + 32];

template <int C1>
struct Conv12Cfg {
  static constexpr int PT = 33, PP = PT * PT, HALO = PT + 1, TP = 128 + 2 * HALO;   // S = 32: 196 rows of conv1 per tile
  static constexpr int C8 = C1 / 8, KSTEPS = C1 / 16;
  static constexpr int A_STAGES = 4, A_BYTES = C8 * TP * 16;
  static constexpr int W2_BYTES = 9 * KSTEPS * 2 * C1 * 16;
  // ring of raw pixel ranges, filled by bulk copies several tiles ahead: a tile's 196 slots touch at most 288
  // consecutive pixels of the [np][1024] array (9 rows of 32; patches are adjacent in memory)
  static constexpr int RAW_STAGES = 8, RAW_BYTES = 288 * 4;
  static constexpr int OFF_W2 = 0, OFF_A = OFF_W2 + W2_BYTES, OFF_RAW = OFF_A + A_STAGES * A_BYTES;
  static constexpr int OFF_BAR = OFF_RAW + RAW_STAGES * RAW_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
  static constexpr int TMEM_COLS = 2 * C1 <= 32 ? 32 : 64;               // D[2 stages][C1]
  static constexpr int NTHREADS = 22 * 32;       // epilogue 0-3, MMA 4, producer group 0 = 5-12, group 1 = 13-20, loader 21
};

template <int C1, int NET, int OUT_MODE>
__global__ void __launch_bounds__(704, 1)
k_conv12(const float* __restrict__ norm, const __half* __restrict__ w2, const float* __restrict__ b2,
         __half* __restrict__ out, size_t out_slots, int np, int ntiles, long long* __restrict__ dbg) {
  using Cfg = Conv12Cfg<C1>;
  constexpr int PT = Cfg::PT, PP = Cfg::PP, HALO = Cfg::HALO, TP = Cfg::TP, S = 32, NA = Cfg::A_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w2_s = smem + Cfg::OFF_W2;
  uint8_t* a_s = smem + Cfg::OFF_A;
  uint8_t* raw_s = smem + Cfg::OFF_RAW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* w_full = bars;                          // W2 landed (bulk copy)
  uint64_t* a_full = bars + 1;                      // [NA] 8 producer warps
  uint64_t* a_empty = a_full + NA;                  // [NA] MMAs of the tile retired
  uint64_t* d_full = a_empty + NA;                  // [2]  MMAs of the tile retired
  uint64_t* d_empty = d_full + 2;                   // [2]  4 epilogue warps
  uint64_t* raw_full = d_empty + 2;                 // [RAW_STAGES] bulk copy landed
  uint64_t* raw_empty = raw_full + Cfg::RAW_STAGES; // [RAW_STAGES] 8 producer warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + Cfg::RAW_STAGES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  // stall accounting (debug launches only): cycles each role spends in each mbarrier wait, CTA 0
  long long stall[3] = {0, 0, 0};
  const bool prof = dbg != nullptr && blockIdx.x == 0;
  const long long t_start = prof ? clock64() : 0;
#define TWAIT(slot, bar, par) do { if (prof) { const long long t0_ = clock64(); mbar_wait(bar, par); stall[slot] += clock64() - t0_; } else mbar_wait(bar, par); } while (0)
  const int n_my = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < NA; i++) { mbar_init(a_full + i, 8); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
    for (int i = 0; i < Cfg::RAW_STAGES; i++) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, 8); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // first pixel (index into norm) of the range a tile's rows can touch, and one past the last
  auto raw_range = [&](int tile, int& lo, int& hi) {
    const int g0 = tile * 128 - HALO, g1 = tile * 128 + 127 + HALO;
    lo = 0;
    if (g0 >= 0) { const int p = g0 / PP, yy = (g0 - p * PP) / PT; lo = p >= np ? np * 1024 : p * 1024 + max(yy - 2, 0) * 32; }
    const int p = g1 / PP, yy = (g1 - p * PP) / PT;
    hi = p >= np ? np * 1024 : p * 1024 + min(yy + 1, 32) * 32;
  };
  if (warp == 21) {
    // ---- loader: W2 once, then the raw pixel range of every tile, RAW_STAGES tiles ahead of the producers
    if (elect_one()) {
      mbar_expect_tx(w_full, Cfg::W2_BYTES);
      bulk_g2s(w2_s, w2, Cfg::W2_BYTES, w_full);
    }
    __syncwarp();
    for (int it = 0; it < n_my; it++) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it % Cfg::RAW_STAGES, ph = (it / Cfg::RAW_STAGES) & 1;
      TWAIT(0, raw_empty + s, ph ^ 1);
      int lo, hi;
      raw_range(tile, lo, hi);
      if (elect_one()) {
        mbar_expect_tx(raw_full + s, (uint32_t)(hi - lo) * 4u);
        if (hi > lo) bulk_g2s(raw_s + s * Cfg::RAW_BYTES, norm + lo, (uint32_t)(hi - lo) * 4u, raw_full + s);
      }
      __syncwarp();
    }
  } else if (warp >= 5) {
    // ---- producers: two groups of 8 warps take alternate tiles; thread = one activation slot of the tile (row r <-> slot
    //      128*tile - HALO + r), all C1 output channels: 9 shared-memory loads, 9*C1 FFMAs with constant-bank weights
    const int grp = warp >= 13 ? 1 : 0;
    const int r = threadIdx.x - (grp ? 13 : 5) * 32;
    for (int it = grp; it < n_my; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it % NA, ph = (it / NA) & 1;
      const int rs = it % Cfg::RAW_STAGES, rph = (it / Cfg::RAW_STAGES) & 1;
      int lo, hi;
      raw_range(tile, lo, hi);
      // branch-free: taps outside the patch read a clamped address and are zeroed by the select
      const int g = tile * 128 - HALO + r;
      const int gc = max(g, 0);
      const int patch = gc / PP;
      const int idx = gc - patch * PP;
      const int yy = idx / PT, x = idx - yy * PT, y = yy - 1;
      const bool valid = r < TP && g >= 0 && patch < np && yy >= 1 && x < S;
      const int base = valid ? patch * 1024 - lo : 0;
      const float* raw = reinterpret_cast<const float*>(raw_s + rs * Cfg::RAW_BYTES);
      TWAIT(0, raw_full + rs, rph);
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const int py = y + t / 3 - 1, px = x + t % 3 - 1;
        const bool inb = valid && py >= 0 && py < S && px >= 0 && px < S;
        const float f = raw[inb ? base + py * 32 + px : 0];
        v[t] = inb ? f : 0.f;
      }
      float acc[C1];
#pragma unroll
      for (int e = 0; e < C1; e++) acc[e] = c_conv1[NET][C1 * 9 + e];
#pragma unroll
      for (int t = 0; t < 9; t++)
#pragma unroll
        for (int e = 0; e < C1; e++) acc[e] = fmaf(c_conv1[NET][e * 9 + t], v[t], acc[e]);
      uint32_t hw[C1 / 2];
#pragma unroll
      for (int e = 0; e < C1 / 2; e++) {
        const __half2 h2 = __floats2half2_rn(valid ? fmaxf(acc[2 * e], 0.f) : 0.f, valid ? fmaxf(acc[2 * e + 1], 0.f) : 0.f);
        hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(raw_empty + rs);      // this warp's reads of the raw stage are done
      TWAIT(1, a_empty + s, ph ^ 1);
      const long long tf0 = prof ? clock64() : 0;
      if (r < TP) {
#pragma unroll
        for (int c8 = 0; c8 < Cfg::C8; c8++)
          *reinterpret_cast<uint4*>(a_s + s * Cfg::A_BYTES + (c8 * TP + r) * 16) =
              make_uint4(hw[4 * c8], hw[4 * c8 + 1], hw[4 * c8 + 2], hw[4 * c8 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + s);
      if (prof) stall[2] += clock64() - tf0;
    }
  } else if (warp == 4) {
    // ---- MMA issuer: 9 x KSTEPS MMAs per tile, back to back from one elected lane
    constexpr uint32_t idesc = instr_desc_f16(C1);
    const uint64_t a_desc0 = smem_desc(smem_u32(a_s), TP * 16, 128);
    const uint64_t w2_desc0 = smem_desc(smem_u32(w2_s), C1 * 16, 128);
    mbar_wait(w_full, 0);
    for (int it = 0; it < n_my; it++) {
      const int s = it % NA, ph = (it / NA) & 1;
      const int ts = it & 1, tph = (it >> 1) & 1;
      TWAIT(0, d_empty + ts, tph ^ 1);
      TWAIT(1, a_full + s, ph);
      fence_after_sync();
      const uint64_t a = a_desc0 + (uint64_t)((s * Cfg::A_BYTES) >> 4);
      const uint32_t d = tmem_base + ts * C1;
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 9; t++) {
          const int shift = (t / 3 - 1) * PT + (t % 3 - 1);
#pragma unroll
          for (int ks = 0; ks < Cfg::KSTEPS; ks++)
            mma_f16(d, a + (uint64_t)((2 * ks) * TP + HALO + shift), w2_desc0 + (uint64_t)((t * Cfg::KSTEPS + ks) * 2 * C1), idesc,
                    (t | ks) != 0);
        }
        mma_commit(a_empty + s);
        mma_commit(d_full + ts);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue (warps 0..3): conv2 accumulators -> bias + ReLU -> fp16 -> next layer's layout (as k_conv_umma)
    const int q = warp & 3, m = q * 32 + lane;
    float bias2[C1];
#pragma unroll
    for (int e = 0; e < C1; e++) bias2[e] = __ldg(b2 + e);
    for (int it = 0; it < n_my; it++) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it & 1, ph = (it >> 1) & 1;
      const int g = tile * 128 + m;
      const int patch = g / PP;
      const int idx = g - patch * PP;
      const int yy = idx / PT, x = idx - yy * PT, y = yy - 1;
      const bool valid = patch < np && yy >= 1 && x < S;
      size_t oslot; int oplane0;
      if (OUT_MODE == OUT_NORMAL) { oslot = (size_t)FS + g; oplane0 = 0; }
      else {
        constexpr int PT2 = S / 2 + 1, PP2 = PT2 * PT2;
        oslot = (size_t)FS + (size_t)patch * PP2 + ((y >> 1) + 1) * PT2 + (x >> 1);
        oplane0 = (((y & 1) << 1) | (x & 1)) * (C1 / 8);
      }
      TWAIT(0, d_full + s, ph);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * C1;
#pragma unroll
      for (int cc = 0; cc < C1 / 16; cc++) {
        float v[16];
        tmem_ld16(taddr + cc * 16, v);
        if (valid) {
          __align__(16) __half hh[16];
#pragma unroll
          for (int e = 0; e < 16; e++) hh[e] = __float2half_rn(fmaxf(v[e] + bias2[cc * 16 + e], 0.f));
          __half* o = out + ((size_t)(oplane0 + cc * 2) * out_slots + oslot) * 8;
          *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hh);
          *reinterpret_cast<uint4*>(o + out_slots * 8) = *reinterpret_cast<const uint4*>(hh + 8);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty + s);
    }
  }
#undef TWAIT
  if (prof && lane == 0) {      // per warp: total cycles of the role loop and its stall counters
    long long* o = dbg + warp * 4;
    o[0] = clock64() - t_start; o[1] = stall[0]; o[2] = stall[1]; o[3] = stall[2];
    if (warp == 0) dbg[22 * 4] = n_my;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =================================================================================================
// HardNet head: 8x8 valid conv == GEMM  [patches x 8192] * [8192 x 128]  + BN + L2 norm + quantise
// =================================================================================================
constexpr int HG_STAGES = 3, HG_STAGE_BYTES = 65536;
constexpr int HG_SMEM = HG_STAGES * HG_STAGE_BYTES + 128;
constexpr int HG_KSPLIT = 8, HG_NKB = 64;   // 64 K blocks of 128, 8 per CTA

// Split-K: CTA (m, ks) accumulates K blocks [ks*8, ks*8+8) of M tile m and writes its fp32 partial tile to
// part[ks][m*128 + row][128]; k_head_finish adds the partials in a fixed order (deterministic).
__global__ void __launch_bounds__(192, 1)
k_head_gemm(const __half* __restrict__ act, size_t slots, const __half* __restrict__ wts,
            float* __restrict__ part, int m_pad, const int* __restrict__ cnt_dev) {
  if (cnt_dev != nullptr && (int)blockIdx.x * 128 >= *cnt_dev) return;     // whole M tile beyond the live patches
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HG_STAGES * HG_STAGE_BYTES);
  uint64_t* full = bars;                 // [3]
  uint64_t* empty = bars + HG_STAGES;    // [3]
  uint64_t* t_full = bars + 2 * HG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HG_STAGES + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform roles (see k_conv_umma)
  const int m0 = blockIdx.x * 128, ksp = blockIdx.y;
  constexpr int NKB = HG_NKB / HG_KSPLIT;
  const int kb0 = ksp * NKB;
  if (threadIdx.x == 0) {
    for (int i = 0; i < HG_STAGES; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    for (int i = 0; i < NKB; i++) {
      const int kb = kb0 + i, s = i % HG_STAGES, ph = (i / HG_STAGES) & 1;
      mbar_wait(empty + s, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full + s, HG_STAGE_BYTES);
        uint8_t* dst = smem + s * HG_STAGE_BYTES;
#pragma unroll 1
        for (int pl = 0; pl < 16; pl++)
          bulk_g2s(dst + pl * 2048, act + ((size_t)(kb * 16 + pl) * slots + m0) * 8, 2048, full + s);
        bulk_g2s(dst + 32768, wts + (size_t)kb * 16384, 32768, full + s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc_f16(128);
    const uint64_t a_desc0 = smem_desc(smem_u32(smem), 2048, 128), b_desc0 = smem_desc(smem_u32(smem) + 32768, 2048, 128);
    for (int i = 0; i < NKB; i++) {
      const int s = i % HG_STAGES, ph = (i / HG_STAGES) & 1;
      mbar_wait(full + s, ph);
      fence_after_sync();
      const uint64_t so = (uint64_t)((s * HG_STAGE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          mma_f16(tmem_base, a_desc0 + so + (uint64_t)(j * 256), b_desc0 + so + (uint64_t)(j * 256), idesc, (i | j) != 0);
        mma_commit(empty + s);
        if (i == NKB - 1) mma_commit(t_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3, m = q * 32 + lane;
    mbar_wait(t_full, 0);
    fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* dst = reinterpret_cast<float4*>(part + ((size_t)ksp * m_pad + m0 + m) * 128);
#pragma unroll 1
    for (int cc = 0; cc < 8; cc++) {
      float v[16];
      tmem_ld16(taddr + cc * 16, v);
#pragma unroll
      for (int e = 0; e < 4; e++) dst[cc * 4 + e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// one warp per patch: sum of the K-split partials + folded BN bias -> L2Norm (desc_server.py:49-52) ->
// uint8(clip(210*(d+0.45),0,255)) as float (desc_server.py:42)
__global__ void k_head_finish(const float* __restrict__ part, int m_pad, const float* __restrict__ bias,
                              float* __restrict__ out, int np, const int* __restrict__ cnt_dev) {
  np = live_patches(np, cnt_dev, 0);
  const int patch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (patch >= np) return;
  float4 acc = *reinterpret_cast<const float4*>(bias + lane * 4);
  for (int ks = 0; ks < HG_KSPLIT; ks++) {
    const float4 p = *reinterpret_cast<const float4*>(part + ((size_t)ks * m_pad + patch) * 128 + lane * 4);
    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  }
  float ss = acc.x * acc.x;
  ss = fmaf(acc.y, acc.y, ss); ss = fmaf(acc.z, acc.z, ss); ss = fmaf(acc.w, acc.w, ss);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float norm = sqrtf(ss + 1e-10f);
  float v[4] = {acc.x, acc.y, acc.z, acc.w}, o4[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float d = v[e] / norm;
    double qd = 210.0 * ((double)d + 0.45);
    qd = qd < 0.0 ? 0.0 : (qd > 255.0 ? 255.0 : qd);
    o4[e] = (float)(int)qd;
  }
  *reinterpret_cast<float4*>(out + (size_t)patch * 128 + lane * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
}

// =================================================================================================
// AffNet / OriNet heads on CUDA cores (input: conv6 output, S = 8, 64 channels, NORMAL layout)
// =================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void load8(const __half* p, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// AffNet: conv 8x8 (64->3, bias) -> tanh -> +1 on outputs 0 and 2 (affnet_server.py:64-66,:80-84)
// OriNet: conv 8x8 pad 1 (64->2, bias) -> 3x3 map -> tanh -> mean (orinet_server.py:64-70)
// Weights live in shared memory as [o][e/4][item][4] (item = c8*64 + pix, e = channel within the 8-channel plane) so that
// the 32 lanes of a warp, which walk consecutive items, read consecutive float4s.  A warp works on HEAD_PB patches at
// once (OriNet): every pair of weight LDS.128 feeds 8 FMAs of each of the HEAD_PB patches (one patch per warp was bound by the
// shared-memory loads: 8 TFLOP/s); the accumulation order per (patch, position, output, lane) is unchanged.
// OriNet: 4 patches per warp, 4 warps per CTA; AffNet (12k MACs per patch, bound by its 8 KB activation read): 1 patch
// per warp, 8 warps per CTA
template <int NOUT, bool ORI, int HEAD_PB, int NW>
__global__ void __launch_bounds__(NW * 32)
k_head_small(const __half* __restrict__ act, size_t slots, const float* __restrict__ w, const float* __restrict__ b,
             float* __restrict__ out, int np, const int* __restrict__ cnt_dev, int cnt_base) {
  np = live_patches(np, cnt_dev, cnt_base);
  if (np <= 0) return;
  extern __shared__ float ws[];   // NOUT * 8 * 512
  // `w` already has the shared-memory layout (modsgpu_load_weights): a straight 16-byte copy
  for (int i = threadIdx.x; i < NOUT * 1024; i += blockDim.x) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NPOS = ORI ? 9 : 1;
  for (int p0 = (blockIdx.x * NW + warp) * HEAD_PB; p0 < np; p0 += gridDim.x * NW * HEAD_PB) {
    float acc[HEAD_PB][NPOS][NOUT];
#pragma unroll
    for (int q = 0; q < HEAD_PB; q++)
#pragma unroll
      for (int p = 0; p < NPOS; p++)
#pragma unroll
        for (int o = 0; o < NOUT; o++) acc[q][p][o] = 0.f;
#pragma unroll (ORI ? 1 : 8)
    for (int k = 0; k < 16; k++) {
      const int it = lane + 32 * k;
      const int c8 = it >> 6, pix = it & 63, y = pix >> 3, x = pix & 7;   // a warp reads 32 consecutive pixels of one plane
      float a[HEAD_PB][8];
#pragma unroll
      for (int q = 0; q < HEAD_PB; q++) {
        const int patch = min(p0 + q, np - 1);   // the tail group re-reads the last patch; its results are not written
        load8(act + ((size_t)c8 * slots + FS + (size_t)patch * 81 + (y + 1) * 9 + x) * 8, a[q]);
      }
#pragma unroll
      for (int pos = 0; pos < NPOS; pos++) {
        int wit = it;
        if (ORI) {
          const int ky = y - pos / 3 + 1, kx = x - pos % 3 + 1;
          if (ky < 0 || ky > 7 || kx < 0 || kx > 7) continue;
          wit = c8 * 64 + ky * 8 + kx;
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + ((o * 2 + 0) * 512 + wit) * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + ((o * 2 + 1) * 512 + wit) * 4);
#pragma unroll
          for (int q = 0; q < HEAD_PB; q++) {
            float s0 = acc[q][pos][o];
            s0 = fmaf(a[q][0], w0.x, s0); s0 = fmaf(a[q][1], w0.y, s0); s0 = fmaf(a[q][2], w0.z, s0); s0 = fmaf(a[q][3], w0.w, s0);
            s0 = fmaf(a[q][4], w1.x, s0); s0 = fmaf(a[q][5], w1.y, s0); s0 = fmaf(a[q][6], w1.z, s0); s0 = fmaf(a[q][7], w1.w, s0);
            acc[q][pos][o] = s0;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < HEAD_PB; q++) {
      float res[NOUT];
#pragma unroll
      for (int o = 0; o < NOUT; o++) res[o] = 0.f;
#pragma unroll
      for (int p = 0; p < NPOS; p++)
#pragma unroll
        for (int o = 0; o < NOUT; o++) res[o] += tanhf(warp_sum(acc[q][p][o]) + b[o]);
      const int patch = p0 + q;
      if (lane == 0 && patch < np) {
        if (ORI) {
          out[patch * 2 + 0] = res[0] / 9.f;
          out[patch * 2 + 1] = res[1] / 9.f;
        } else {
          out[patch * 3 + 0] = res[0] + 1.f;
          out[patch * 3 + 1] = res[1];
          out[patch * 3 + 2] = res[NOUT - 1] + 1.f;
        }
      }
    }
  }
}

// =================================================================================================
// debug probe: one 128 x N x 64 GEMM through the same descriptor conventions (tests only)
// =================================================================================================
__global__ void __launch_bounds__(128, 1)
k_umma_probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int swap_lbo_sbo) {
  // A: [8 planes][128 rows][8], B: [8 planes][32 rows][8]  (K = 64, N = 32)
  __shared__ __align__(1024) uint8_t sa[8 * 128 * 16];
  __shared__ __align__(1024) uint8_t sb[8 * 32 * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 128; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(A)[i];
  for (int i = threadIdx.x; i < 8 * 32; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(B)[i];
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(&tslot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = instr_desc_f16(32);
    for (int ks = 0; ks < 4; ks++) {
      uint32_t a = smem_u32(sa) + (2 * ks) * 128 * 16, b = smem_u32(sb) + (2 * ks) * 32 * 16;
      uint64_t da = swap_lbo_sbo ? smem_desc(a, 128, 128 * 16) : smem_desc(a, 128 * 16, 128);
      uint64_t db = swap_lbo_sbo ? smem_desc(b, 128, 32 * 16) : smem_desc(b, 32 * 16, 128);
      mma_f16(tb, da, db, idesc, ks != 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const int m = threadIdx.x;
  for (int cc = 0; cc < 2; cc++) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + cc * 16, v);
    for (int e = 0; e < 16; e++) D[m * 32 + cc * 16 + e] = v[e];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tb);
}

// =================================================================================================
// pacing probe (tests / profiling only): `reps` back-to-back tcgen05.mma (M = 128, K = 16, runtime N) from one
// thread, operands = zero-filled shared memory, descriptor fields given by the caller.  Returns the clock64
// span from the first issue to the commit's mbarrier flip.  Used to measure how the operand layout (row
// alignment of the A start address, LBO/SBO, swizzle mode) paces the MMA pipe.
// =================================================================================================
__global__ void __launch_bounds__(128, 1)
k_umma_pace(int n, int a_off_bytes, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int layout, int reps, int n_acc, int a_step_bytes,
            long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];        // 96 KB A region + 32 KB B region, zeroed
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (131072 / 16); i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tslot;
  if (__shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) == 0) {     // warp-uniform branch: operands stay in uniform registers
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem) + 4096 + a_off_bytes, b0 = smem_u32(smem) + 98304;
    const uint64_t lay = (uint64_t)(layout & 7) << 61;
    const uint64_t db = smem_desc(b0, b_lbo, b_sbo) | lay;
    const uint64_t da = smem_desc(a0, a_lbo, a_sbo) | lay;
    const uint32_t astep = (uint32_t)a_step_bytes >> 4;
    const uint32_t tstep = (n_acc > 1) ? (uint32_t)n : 0u;
    const long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 8
      for (int r = 0; r < reps; r++)
        mma_f16(tb + (r & 1) * tstep, da + (uint64_t)((r & 7) * astep), db, idesc, 1);
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

// =================================================================================================
// host side
// =================================================================================================
size_t plane_slots(int cap, int S) { return (size_t)FS + (size_t)round_up(cap * (S + 1) * (S + 1), 128) + 128; }

// [Cout][3][3][Cin] fp32 -> per output slice nsp: [tap][k16][2][COUT_T][8] fp16
std::vector<__half> pack_conv(const NpzArray& w, int cin, int cout, int nsplit) {
  const int cout_t = cout / nsplit, ksteps = cin / 16;
  std::vector<__half> r((size_t)9 * cin * cout);
  size_t o = 0;
  for (int ns = 0; ns < nsplit; ns++)
    for (int t = 0; t < 9; t++)
      for (int ks = 0; ks < ksteps; ks++)
        for (int j = 0; j < 2; j++)
          for (int n = 0; n < cout_t; n++)
            for (int e = 0; e < 8; e++)
              r[o++] = __float2half_rn(w.data[((size_t)(ns * cout_t + n) * 9 + t) * cin + ks * 16 + j * 8 + e]);
  return r;
}

template <class T>
cudaError_t upload(T** dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMalloc((void**)dst, bytes);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
}

int cnn_chunk_cap() {
  const char* e = getenv("MODSGPU_CNN_CHUNK");
  // 4096 patches per launch: the per-launch prologue (weights into smem, TMEM alloc) and the tail wave are
  // amortised over ~30k M tiles; activations of a chunk (<= 1.1 GB for HardNet++) stream through HBM
  // (measured on B200: CNN kernel time per pair 6.0 ms @512, 4.4 @1024, 3.6 @2048, 3.2 @4096)
  int v = e ? atoi(e) : 4608;
  if (v < 128) v = 128;
  return round_up(v, 128);
}

template <int CIN, int COUT_T, int NSPLIT, int S, int NGRP, int OUT_MODE>
int launch_conv(modsgpu_ctx* ctx, const __half* in, size_t in_slots, const ConvW& w, __half* out, size_t out_slots, int np,
                int patch_base, const int* cnt_dev) {
  using Cfg = ConvCfg<CIN, COUT_T, NSPLIT, S, NGRP, OUT_MODE>;
  auto kern = k_conv_umma<CIN, COUT_T, NSPLIT, S, NGRP, OUT_MODE>;
  static OnceFlags attr;
  if (attr.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr.set(ctx->device);
  }
  const int ntiles = ceil_div(np * Cfg::PP, 128);
  static char pname[64] = {0};
  if (!pname[0]) snprintf(pname, sizeof(pname), "k_conv_umma<%d,%d,S%d,%s>", CIN, Cfg::COUT, S, NGRP == 4 ? "s2" : "s1");
  constexpr int S_IN = NGRP == 4 ? 2 * S : S;   // stride-2 layers read the parity planes of a 2S x 2S map
  MG_PROF2(ctx, pname, 1, 2.0 * np * S * S * 9.0 * CIN * Cfg::COUT, 2.0 * np * ((double)S_IN * S_IN * CIN + (double)S * S * Cfg::COUT));
  static const bool one_cta = [] { const char* e = getenv("MODSGPU_CONV_ONE_CTA"); return e && atoi(e) != 0; }();   // A/B switch
  int gx = std::min(ntiles, std::max(1, ctx->num_sms * ((Cfg::TWO_CTAS && !one_cta) ? 2 : 1) / NSPLIT));
  dim3 grid(gx, NSPLIT);
  static const int contig = [] { const char* e = getenv("MODSGPU_CONV_STRIDED_TILES"); return (e && atoi(e) != 0) ? 0 : 1; }();   // A/B switch
  // depth of the A ring: MODSGPU_CONV_STAGES caps it (default: as deep as the layer's configuration allows)
  static const int stage_cap = [] { const char* e = getenv("MODSGPU_CONV_STAGES"); return e ? std::max(2, atoi(e)) : 64; }();
  const int nst = std::min(Cfg::NST, stage_cap);
  const int smem_bytes = Cfg::W_BYTES + nst * Cfg::A_BYTES + 256;
  kern<<<grid, 192, smem_bytes, ctx->stream>>>(in, in_slots, w.w, w.b, out, out_slots, np, ntiles, patch_base, contig, nst,
                                               cnt_dev, patch_base);
  MG_LAUNCHED(ctx);
  return 0;
}


// conv2..conv4 (C1 = 16) / conv2..conv3 (C1 = 32) in one launch (k_trunk).  MODSGPU_NO_FUSED_TRUNK=1 keeps the layer-by-layer
// path (the parity test runs both and compares them bit for bit).
// k_conv1_mma is OPT-IN (MODSGPU_CONV1_MMA=1, read per call).  Measured on B200 in three shapes (dedicated epilogue warps;
// workers that build and drain; 4-6 accumulators deep): 0.26 / 0.18 ms per pair for C1 = 16 / 32 against 0.20 / 0.17 ms of
// k_conv1 -- whatever the structure, because BOTH kernels sit on the same bound: conv1's map is 32 / 64 KB per patch of pure
// HBM WRITE (1.02 GB per pair), and a write-only stream saturates near 2.5-2.9 TB/s on this part, not at the 6.5 TB/s of
// the read + write copy peak (ncu: 195 MB written in 78 us, L2 hit rate 7 %).  The FMA kernel stays the product path.
bool conv1_mma_enabled() {
  const char* e = getenv("MODSGPU_CONV1_MMA");
  return e && atoi(e) != 0;
}
bool fused_trunk_enabled() {      // read per call: the parity test switches paths inside one process
  const char* e = getenv("MODSGPU_NO_FUSED_TRUNK");
  return !(e && atoi(e) != 0);
}
template <int C1, int NL>
int launch_trunk(modsgpu_ctx* ctx, const __half* in, size_t in_slots, const ConvW* w /* conv2.. */, __half* out, size_t out_slots,
                 int np, int patch_base, const int* cnt_dev) {
  using Cfg = TrunkCfg<C1, NL>;
  auto kern = k_trunk<C1, NL>;
  static OnceFlags attr;
  if (attr.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr.set(ctx->device);
  }
  static char pname[64] = {0};
  if (!pname[0]) snprintf(pname, sizeof(pname), "k_trunk<%d,conv2-%d>", C1, NL + 1);
  const double macs = 1024.0 * 9 * C1 * C1 + 256.0 * 9 * C1 * 2 * C1 + (NL == 3 ? 256.0 * 9 * 2 * C1 * 2 * C1 : 0.0);
  // algorithmic bytes: conv1's map in, the last fused layer's map out
  MG_PROF2(ctx, pname, 1, 2.0 * np * macs, 2.0 * np * (1024.0 * C1 + 256.0 * 2 * C1));
  // MODSGPU_TRUNK_CTAS=1: one CTA per SM for the 16-channel trunk as well (A/B switch: leaves half of the shared memory
  // to the kernels of other streams)
  static const int cta_cap = [] { const char* e = getenv("MODSGPU_TRUNK_CTAS"); return e ? std::max(1, atoi(e)) : 2; }();
  const int ctas = ctx->num_sms * (C1 == 16 ? std::min(2, cta_cap) : 1);
  kern<<<std::min(np, ctas), 192, Cfg::SMEM_BYTES, ctx->stream>>>(in, in_slots, w[0].w, w[0].b, w[1].w, w[1].b, NL == 3 ? w[2].w : nullptr,
                                                                  NL == 3 ? w[2].b : nullptr, out, out_slots, np, cnt_dev, patch_base);
  MG_LAUNCHED(ctx);
  return 0;
}

// conv1 + conv2 in one launch (k_conv12).  EXPERIMENTAL, opt-in with MODSGPU_FUSED_CONV12=1: parity-green but slower on
// B200 than k_conv1 + k_conv_umma (profiles/r01_conv12_stalls.txt has the three variants and their stall tables).
bool fused_conv12_enabled() {
  static const bool on = [] { const char* e = getenv("MODSGPU_FUSED_CONV12"); return e && atoi(e) != 0; }();
  return on;
}
template <int C1, int NET>
int launch_conv12(modsgpu_ctx* ctx, const uint8_t* patches, const ConvW& w2, __half* out, size_t out_slots, int np) {
  using Cfg = Conv12Cfg<C1>;
  auto kern = k_conv12<C1, NET, OUT_PARITY>;
  static OnceFlags attr;
  if (attr.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr.set(ctx->device);
  }
  MG_CUDA(ctx, ctx->cnn_stats.ensure((size_t)np * 4096));
  MG_PROF(ctx, "k_patch_prep", 0, (double)np * (1024.0 + 4096.0));
  k_patch_prep<<<np, 256, 0, ctx->stream>>>(patches, np, ctx->cnn_stats.as<float>());
  MG_LAUNCHED(ctx);
  const int ntiles = ceil_div(np * Cfg::PP, 128);
  static char pname[64] = {0};
  if (!pname[0]) snprintf(pname, sizeof(pname), "k_conv12<1,%d,%d,S32>", C1, C1);
  // algorithmic bytes: the normalised patch in, the conv2 map out (conv1's map stays on the SM)
  MG_PROF2(ctx, pname, 1, 2.0 * np * 1024 * 9.0 * (C1 + (double)C1 * C1), (double)np * (4096.0 + 1024.0 * C1 * 2));
  static const bool debug = [] { const char* e = getenv("MODSGPU_CONV12_DEBUG"); return e && atoi(e) != 0; }();
  long long* dbg = nullptr;
  if (debug) { MG_CUDA(ctx, ctx->io_c.ensure(23 * 4 * 8)); dbg = ctx->io_c.as<long long>(); }
  kern<<<std::min(ntiles, ctx->num_sms), Cfg::NTHREADS, Cfg::SMEM_BYTES, ctx->stream>>>(
      ctx->cnn_stats.as<float>(), w2.w, w2.b, out, out_slots, np, ntiles, dbg);
  MG_LAUNCHED(ctx);
  if (debug) {      // stall table of CTA 0 (cycles): role loop total, then its wait counters
    long long h[23 * 4];
    MG_CUDA(ctx, cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    const long long n = h[22 * 4] > 0 ? h[22 * 4] : 1;
    fprintf(stderr, "k_conv12<%d> np %d tiles/CTA %lld (cycles per tile: loop, wait0, wait1, store+fence+arrive)\n", C1, np, n);
    for (int w : {0, 3, 4, 5, 11, 13, 21})
      fprintf(stderr, "  warp %2d %-4s %8.0f %8.0f %8.0f %8.0f\n", w, w < 4 ? "epi" : w == 4 ? "mma" : w == 21 ? "load" : "prod",
              (double)h[w * 4] / n, (double)h[w * 4 + 1] / n, (double)h[w * 4 + 2] / n, (double)h[w * 4 + 3] / n);
  }
  return 0;
}

}  // namespace

static void free_net(NetWeights* nw) {
  if (!nw) return;
  if (!nw->origin) {
    cudaFree(nw->c1_w); cudaFree(nw->c1_b); cudaFree(nw->c1_split);
    for (auto& c : nw->conv) { cudaFree(c.w); cudaFree(c.b); }
    cudaFree(nw->head_w16); cudaFree(nw->head_w32); cudaFree(nw->head_b);
  }
  for (auto a : nw->act) cudaFree(a);
  delete nw;
}

// The nets of a sibling context (modsgpu_ctx_sibling): the weights of `src` shared, the activation buffers its own --
// the two contexts run their forward passes at the same time.
int mg_nets_share(modsgpu_ctx* sib, const modsgpu_ctx* src) {
  for (int i = 0; i < 3; i++) {
    const NetWeights* o = src->nets[i];
    NetWeights* c = sib->nets[i];
    if (c && c->origin == o && o && c->c1_w == o->c1_w) continue;      // still the clone of the loaded weights
    if (c) { free_net(c); sib->nets[i] = nullptr; }
    if (!o) continue;
    c = new NetWeights(*o);
    c->origin = o;
    for (int k = 0; k < 6; k++) {
      c->act[k] = nullptr;
      if (!o->act[k]) continue;
      const int C1 = o->C1;
      const int planes[6] = {C1 / 8, 4 * C1 / 8, 2 * C1 / 8, 4 * 2 * C1 / 8, 4 * C1 / 8, 4 * C1 / 8};
      const size_t bytes = o->slots[k] * planes[k] * 16;
      if (cudaMalloc((void**)&c->act[k], bytes) != cudaSuccess || cudaMemset(c->act[k], 0, bytes) != cudaSuccess) {
        free_net(c);
        sib->err = "sibling context: activation buffers could not be allocated";
        return MODSGPU_ECUDA;
      }
    }
    sib->nets[i] = c;
  }
  return 0;
}

void mg_free_nets(modsgpu_ctx* ctx) {
  for (int i = 0; i < 3; i++) { free_net(ctx->nets[i]); ctx->nets[i] = nullptr; }
}

extern "C" int modsgpu_net_out_dim(modsgpu_net net) { return net == MODSGPU_AFFNET ? 3 : (net == MODSGPU_ORINET ? 2 : 128); }

extern "C" int modsgpu_load_weights(modsgpu_ctx* ctx, modsgpu_net net, const char* path) {
  if (!ctx || !path || (int)net < 0 || (int)net > 2) return MODSGPU_EINVAL;
  MG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::map<std::string, NpzArray> z;
  std::string err;
  if (!npz_load(path, z, err)) MG_FAIL(ctx, MODSGPU_EIO, "weights: " + err);
  const int C1 = net == MODSGPU_HARDNET ? 32 : 16;
  const int cins[5] = {C1, C1, 2 * C1, 2 * C1, 4 * C1}, couts[5] = {C1, 2 * C1, 2 * C1, 4 * C1, 4 * C1};
  const int nsplit[5] = {1, 1, 1, net == MODSGPU_HARDNET ? 2 : 1, net == MODSGPU_HARDNET ? 2 : 1};
  auto need = [&](const char* k, std::vector<int> shape) -> const NpzArray* {
    auto it = z.find(k);
    if (it == z.end() || it->second.shape != shape) return nullptr;
    return &it->second;
  };
  if (ctx->nets[net]) { free_net(ctx->nets[net]); ctx->nets[net] = nullptr; }
  NetWeights* nw = new NetWeights();
  nw->net = net; nw->C1 = C1; nw->out_dim = modsgpu_net_out_dim(net);
  const NpzArray* w1 = need("c1_w", {C1, 3, 3, 1});
  const NpzArray* b1 = need("c1_b", {C1});
  if (!w1 || !b1) { delete nw; MG_FAIL(ctx, MODSGPU_EIO, "weights: c1_w/c1_b missing or wrong shape"); }
  MG_CUDA(ctx, upload(&nw->c1_w, w1->data.data(), w1->data.size() * 4));
  MG_CUDA(ctx, upload(&nw->c1_b, b1->data.data(), b1->data.size() * 4));
  {
    // split conv1 weights for the tensor-pipe kernel: B[k][n] in the K order of the im2col rows (three groups of 10:
    // taps 0..8 + one zero): k = 0..9 wh (pairs with xh), 10..19 wh (pairs with xl), 20..29 wl (pairs with xh), 30..31 zero
    std::vector<__half> ws((size_t)4 * C1 * 8);
    for (int n = 0; n < C1; n++)
      for (int k = 0; k < 32; k++) {
        float v = 0.f;
        const int grp = k / 10, tap = k % 10;
        if (grp < 3 && tap < 9) {
          const float w = w1->data[(size_t)n * 9 + tap];
          const __half wh = __float2half_rn(w);
          v = grp < 2 ? __half2float(wh) : w - __half2float(wh);
        }
        ws[((size_t)(k >> 3) * C1 + n) * 8 + (k & 7)] = __float2half_rn(v);
      }
    MG_CUDA(ctx, upload(&nw->c1_split, ws.data(), ws.size() * 2));
  }
  {
    // k_conv12 reads conv1's weights from the constant bank, one slot per (device, net).  The first context that loads a
    // net on a device claims the slot; a context that loads DIFFERENT weights for it keeps the two-kernel path.
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::vector<float>> resident;
    std::vector<float> blob(w1->data);
    blob.insert(blob.end(), b1->data.begin(), b1->data.end());
    std::lock_guard<std::mutex> lk(mu);
    auto it = resident.find({ctx->device, (int)net});
    if (it == resident.end()) {
      MG_CUDA(ctx, cudaMemcpyToSymbol(c_conv1, blob.data(), blob.size() * 4, (size_t)net * (32 * 9 + 32) * 4));
      resident[{ctx->device, (int)net}] = blob;
      nw->fused12 = true;
    } else {
      nw->fused12 = it->second == blob;
    }
    nw->fused12 = nw->fused12 && fused_conv12_enabled();
  }
  for (int l = 0; l < 5; l++) {
    std::string wn = "c" + std::to_string(l + 2) + "_w", bn = "c" + std::to_string(l + 2) + "_b";
    const NpzArray* w = need(wn.c_str(), {couts[l], 3, 3, cins[l]});
    const NpzArray* b = need(bn.c_str(), {couts[l]});
    if (!w || !b) { delete nw; MG_FAIL(ctx, MODSGPU_EIO, "weights: " + wn + " missing or wrong shape"); }
    std::vector<__half> pk = pack_conv(*w, cins[l], couts[l], nsplit[l]);
    MG_CUDA(ctx, upload(&nw->conv[l].w, pk.data(), pk.size() * 2));
    MG_CUDA(ctx, upload(&nw->conv[l].b, b->data.data(), b->data.size() * 4));
  }
  const NpzArray* hw = need("h_w", {nw->out_dim, 8, 8, 4 * C1});
  const NpzArray* hb = need("h_b", {nw->out_dim});
  if (!hw || !hb) { delete nw; MG_FAIL(ctx, MODSGPU_EIO, "weights: h_w/h_b missing or wrong shape"); }
  MG_CUDA(ctx, upload(&nw->head_b, hb->data.data(), hb->data.size() * 4));
  if (net == MODSGPU_HARDNET) {
    // K = (y*8+x)*128 + c ; blocks of k16: [2][128 n][8]
    std::vector<__half> pk((size_t)128 * 8192);
    size_t o = 0;
    for (int kk = 0; kk < 512; kk++)
      for (int j = 0; j < 2; j++)
        for (int n = 0; n < 128; n++)
          for (int e = 0; e < 8; e++) pk[o++] = __float2half_rn(hw->data[(size_t)n * 8192 + kk * 16 + j * 8 + e]);
    MG_CUDA(ctx, upload(&nw->head_w16, pk.data(), pk.size() * 2));
  } else {
    // k_head_small's shared-memory image, prepared once: [o][e/4][item = (c/8)*64 + pix][e%4] with e = c % 8
    const int nout = nw->out_dim;
    std::vector<float> img((size_t)nout * 4096);
    for (int idx = 0; idx < nout * 4096; idx++) {
      const int o = idx >> 12, rem = idx & 4095, pix = rem >> 6, c = rem & 63;
      const int e = c & 7, item = (c >> 3) * 64 + pix;
      img[((size_t)(o * 2 + (e >> 2)) * 512 + item) * 4 + (e & 3)] = hw->data[idx];
    }
    MG_CUDA(ctx, upload(&nw->head_w32, img.data(), img.size() * 4));
  }
  // activation buffers, zeroed once: pad slots are never written afterwards
  const int cap = cnn_chunk_cap();
  nw->cap = cap;
  const int S_of[6] = {32, 16, 16, 8, 8, 8};
  const int planes[6] = {C1 / 8, 4 * C1 / 8, 2 * C1 / 8, 4 * 2 * C1 / 8, 4 * C1 / 8, 4 * C1 / 8};
  for (int i = 0; i < 6; i++) {
    size_t slots = plane_slots(cap, S_of[i]);
    int npl = planes[i];
    if (i == 5 && net == MODSGPU_HARDNET) continue;   // conv6 writes straight into the head GEMM operand (ctx->cnn_act0)
    nw->slots[i] = slots;
    size_t bytes = slots * npl * 16;
    MG_CUDA(ctx, cudaMalloc((void**)&nw->act[i], bytes));
    MG_CUDA(ctx, cudaMemset(nw->act[i], 0, bytes));
  }
  ctx->nets[net] = nw;
  return 0;
}

// Enqueue the forward pass of `net` on n device-resident 32x32 u8 patches; d_out: n x out_dim floats.
// cnt_dev (optional, device): the live patch count when n is only the host's upper bound (chain.cu).
int mg_net_forward_enqueue(modsgpu_ctx* ctx, modsgpu_net net, const uint8_t* d_patches, int n, float* d_out, const int* cnt_dev) {
  NetWeights* nw = ctx->nets[net];
  if (!nw) MG_FAIL(ctx, MODSGPU_ESTATE, "modsgpu_load_weights has not been called for this net");
  // HardNet: conv6 of every chunk lands in one [8192/8 planes][m_pad patches] operand; the 8x8 head then runs
  // once over all patches as a split-K GEMM
  const int m_pad = round_up(n, 128);
  __half* act6all = nullptr;
  float* part = nullptr;
  if (net == MODSGPU_HARDNET) {
    MG_CUDA(ctx, ctx->cnn_act0.ensure((size_t)1024 * m_pad * 16));
    MG_CUDA(ctx, ctx->cnn_act1.ensure((size_t)HG_KSPLIT * m_pad * 128 * 4));
    act6all = ctx->cnn_act0.as<__half>();
    part = ctx->cnn_act1.as<float>();
  }
  // balanced chunks: 4300 patches run as 2 x 2176 rather than 4096 + 204
  const int nchunks = ceil_div(n, nw->cap);
  const int per = std::min(nw->cap, round_up(ceil_div(n, nchunks), 128));
  for (int p0 = 0; p0 < n; p0 += per) {
    const int np = std::min(per, n - p0);
    const uint8_t* pin = d_patches + (size_t)p0 * 1024;
    float* pout = d_out + (size_t)p0 * nw->out_dim;
    int rc = 0;
    if (net == MODSGPU_HARDNET) {
      if (nw->fused12) {
        if ((rc = launch_conv12<32, 2>(ctx, pin, nw->conv[0], nw->act[1], nw->slots[1], np))) return rc;
      } else {
        MG_PROF2(ctx, "k_conv1<32>", 1, 2.0 * np * 1024 * 9 * 32, (double)np * (1024.0 + 1024.0 * 32 * 2));
        if (conv1_mma_enabled()) {
          static OnceFlags c1attr;
          if (c1attr.need(ctx->device)) {
            MG_CUDA(ctx, cudaFuncSetAttribute(k_conv1_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1MmaCfg<32>::SMEM_DYN));
            c1attr.set(ctx->device);
          }
          k_conv1_mma<32><<<std::min(np, 3 * ctx->num_sms), 160, Conv1MmaCfg<32>::SMEM_DYN, ctx->stream>>>(pin, np, nw->c1_split, nw->c1_b, nw->act[0], nw->slots[0], cnt_dev, p0);
        }
        else
          k_conv1<32><<<std::min(np, 3 * ctx->num_sms), 160, 0, ctx->stream>>>(pin, np, nw->c1_w, nw->c1_b, nw->act[0], nw->slots[0], cnt_dev, p0);
        MG_LAUNCHED(ctx);
        if (fused_trunk_enabled()) {
          if ((rc = launch_trunk<32, 2>(ctx, nw->act[0], nw->slots[0], nw->conv, nw->act[2], nw->slots[2], np, p0, cnt_dev))) return rc;
        } else {
          if ((rc = launch_conv<32, 32, 1, 32, 1, OUT_PARITY>(ctx, nw->act[0], nw->slots[0], nw->conv[0], nw->act[1], nw->slots[1], np, p0, cnt_dev))) return rc;
        }
      }
      if (nw->fused12 || !fused_trunk_enabled())
        if ((rc = launch_conv<32, 64, 1, 16, 4, OUT_NORMAL>(ctx, nw->act[1], nw->slots[1], nw->conv[1], nw->act[2], nw->slots[2], np, p0, cnt_dev))) return rc;
      if ((rc = launch_conv<64, 64, 1, 16, 1, OUT_PARITY>(ctx, nw->act[2], nw->slots[2], nw->conv[2], nw->act[3], nw->slots[3], np, p0, cnt_dev))) return rc;
      if ((rc = launch_conv<64, 64, 2, 8, 4, OUT_NORMAL>(ctx, nw->act[3], nw->slots[3], nw->conv[3], nw->act[4], nw->slots[4], np, p0, cnt_dev))) return rc;
      if ((rc = launch_conv<128, 64, 2, 8, 1, OUT_GEMM>(ctx, nw->act[4], nw->slots[4], nw->conv[4], act6all, (size_t)m_pad, np, p0, cnt_dev))) return rc;
    } else {
      if (nw->fused12) {
        if ((rc = (net == MODSGPU_AFFNET ? launch_conv12<16, 0>(ctx, pin, nw->conv[0], nw->act[1], nw->slots[1], np)
                                         : launch_conv12<16, 1>(ctx, pin, nw->conv[0], nw->act[1], nw->slots[1], np)))) return rc;
      } else {
        MG_PROF2(ctx, "k_conv1<16>", 1, 2.0 * np * 1024 * 9 * 16, (double)np * (1024.0 + 1024.0 * 16 * 2));
        if (conv1_mma_enabled()) {
          static OnceFlags c1attr;
          if (c1attr.need(ctx->device)) {
            MG_CUDA(ctx, cudaFuncSetAttribute(k_conv1_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1MmaCfg<16>::SMEM_DYN));
            c1attr.set(ctx->device);
          }
          k_conv1_mma<16><<<std::min(np, 3 * ctx->num_sms), 160, Conv1MmaCfg<16>::SMEM_DYN, ctx->stream>>>(pin, np, nw->c1_split, nw->c1_b, nw->act[0], nw->slots[0], cnt_dev, p0);
        }
        else
          k_conv1<16><<<std::min(np, 3 * ctx->num_sms), 160, 0, ctx->stream>>>(pin, np, nw->c1_w, nw->c1_b, nw->act[0], nw->slots[0], cnt_dev, p0);
        MG_LAUNCHED(ctx);
        if (fused_trunk_enabled()) {
          if ((rc = launch_trunk<16, 3>(ctx, nw->act[0], nw->slots[0], nw->conv, nw->act[3], nw->slots[3], np, p0, cnt_dev))) return rc;
        } else {
          if ((rc = launch_conv<16, 16, 1, 32, 1, OUT_PARITY>(ctx, nw->act[0], nw->slots[0], nw->conv[0], nw->act[1], nw->slots[1], np, p0, cnt_dev))) return rc;
        }
      }
      if (nw->fused12 || !fused_trunk_enabled()) {
        if ((rc = launch_conv<16, 32, 1, 16, 4, OUT_NORMAL>(ctx, nw->act[1], nw->slots[1], nw->conv[1], nw->act[2], nw->slots[2], np, p0, cnt_dev))) return rc;
        if ((rc = launch_conv<32, 32, 1, 16, 1, OUT_PARITY>(ctx, nw->act[2], nw->slots[2], nw->conv[2], nw->act[3], nw->slots[3], np, p0, cnt_dev))) return rc;
      }
      if ((rc = launch_conv<32, 64, 1, 8, 4, OUT_NORMAL>(ctx, nw->act[3], nw->slots[3], nw->conv[3], nw->act[4], nw->slots[4], np, p0, cnt_dev))) return rc;
      if ((rc = launch_conv<64, 64, 1, 8, 1, OUT_NORMAL>(ctx, nw->act[4], nw->slots[4], nw->conv[4], nw->act[5], nw->slots[5], np, p0, cnt_dev))) return rc;
      static OnceFlags hattr;
      if (hattr.need(ctx->device)) {
        MG_CUDA(ctx, cudaFuncSetAttribute(k_head_small<3, false, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 4096 * 4));
        MG_CUDA(ctx, cudaFuncSetAttribute(k_head_small<2, true, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 4));
        hattr.set(ctx->device);
      }
      MG_PROF(ctx, net == MODSGPU_AFFNET ? "k_head_aff" : "k_head_ori", 1, 2.0 * np * (net == MODSGPU_AFFNET ? 12288.0 : 61952.0));
      if (net == MODSGPU_AFFNET)   // 48 KB of weights per CTA: 4 CTAs (32 warps) per SM
        k_head_small<3, false, 1, 8><<<std::min(ceil_div(np, 8), 4 * ctx->num_sms), 256, 3 * 4096 * 4, ctx->stream>>>(
            nw->act[5], nw->slots[5], nw->head_w32, nw->head_b, pout, np, cnt_dev, p0);
      else                         // 4 warps x 4 patches per CTA pass, 160 registers: 3 CTAs per SM
        k_head_small<2, true, 4, 4><<<std::min(ceil_div(np, 16), 3 * ctx->num_sms), 128, 2 * 4096 * 4, ctx->stream>>>(
            nw->act[5], nw->slots[5], nw->head_w32, nw->head_b, pout, np, cnt_dev, p0);
      MG_LAUNCHED(ctx);
    }
  }
  if (net == MODSGPU_HARDNET) {
    static OnceFlags attr;
    if (attr.need(ctx->device)) { MG_CUDA(ctx, cudaFuncSetAttribute(k_head_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, HG_SMEM)); attr.set(ctx->device); }
    MG_PROF(ctx, "k_head_gemm", 1, 2.0 * n * 8192.0 * 128);
    k_head_gemm<<<dim3(m_pad / 128, HG_KSPLIT), 192, HG_SMEM, ctx->stream>>>(act6all, (size_t)m_pad, nw->head_w16, part, m_pad, cnt_dev);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_head_finish", 0, (double)n * 128 * 4 * (HG_KSPLIT + 1));
    k_head_finish<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(part, m_pad, nw->head_b, d_out, n, cnt_dev);
    MG_LAUNCHED(ctx);
  }
  return 0;
}

extern "C" int modsgpu_net_forward_u8(modsgpu_ctx* ctx, modsgpu_net net, const uint8_t* patches, int n, float* out) {
  if (!ctx || (n > 0 && (!patches || !out)) || n < 0 || (int)net < 0 || (int)net > 2) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (n == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const int D = modsgpu_net_out_dim(net);
  MG_CUDA(ctx, ctx->smp_out.ensure((size_t)n * 1024));
  MG_CUDA(ctx, ctx->cnn_out.ensure((size_t)n * D * 4));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->smp_out.p, patches, (size_t)n * 1024, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_net_forward_enqueue(ctx, net, ctx->smp_out.as<uint8_t>(), n, ctx->cnn_out.as<float>(), nullptr);
  if (rc) return rc;
  return mg_read_back_end(ctx, out, ctx->cnn_out.p, (size_t)n * D * 4);
}

int mg_sample_enqueue(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                      double mrSize, int ps, uint8_t* d_out, float* d_outf);

extern "C" int modsgpu_describe(modsgpu_ctx* ctx, modsgpu_net net, const modsgpu_image* img, const modsgpu_region* regs,
                                int n, double mrSize, int patchSize, float* out) {
  if (!ctx || !img || (n > 0 && (!regs || !out)) || n < 0 || (int)net < 0 || (int)net > 2) return MODSGPU_EINVAL;
  if (patchSize != 32) MG_FAIL(ctx, MODSGPU_EINVAL, "the networks take 32x32 patches");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (n == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const int D = modsgpu_net_out_dim(net);
  MG_CUDA(ctx, ctx->smp_out.ensure((size_t)n * 1024));
  MG_CUDA(ctx, ctx->cnn_out.ensure((size_t)n * D * 4));
  int rc = mg_sample_enqueue(ctx, img, regs, n, mrSize, patchSize, ctx->smp_out.as<uint8_t>(), nullptr);
  if (rc) return rc;
  rc = mg_net_forward_enqueue(ctx, net, ctx->smp_out.as<uint8_t>(), n, ctx->cnn_out.as<float>(), nullptr);
  if (rc) return rc;
  return mg_read_back_end(ctx, out, ctx->cnn_out.p, (size_t)n * D * 4);
}

// test-only: D[128 x 32] = A[128 x 64] * B[32 x 64]^T through the tcgen05 path; A/B given row-major fp32
extern "C" int modsgpu_debug_umma_probe(modsgpu_ctx* ctx, const float* A, const float* B, float* D, int swap_lbo_sbo) {
  if (!ctx || !A || !B || !D) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  std::vector<__half> ha(8 * 128 * 8), hb(8 * 32 * 8);
  for (int c8 = 0; c8 < 8; c8++) {
    for (int r = 0; r < 128; r++) for (int e = 0; e < 8; e++) ha[((size_t)c8 * 128 + r) * 8 + e] = __float2half_rn(A[r * 64 + c8 * 8 + e]);
    for (int r = 0; r < 32; r++) for (int e = 0; e < 8; e++) hb[((size_t)c8 * 32 + r) * 8 + e] = __float2half_rn(B[r * 64 + c8 * 8 + e]);
  }
  MG_CUDA(ctx, ctx->io_a.ensure(ha.size() * 2));
  MG_CUDA(ctx, ctx->io_b.ensure(hb.size() * 2));
  MG_CUDA(ctx, ctx->io_c.ensure(128 * 32 * 4));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
  k_umma_probe<<<1, 128, 0, ctx->stream>>>(ctx->io_a.as<__half>(), ctx->io_b.as<__half>(), ctx->io_c.as<float>(), swap_lbo_sbo);
  MG_LAUNCHED(ctx);
  MG_CUDA(ctx, cudaMemcpyAsync(D, ctx->io_c.p, 128 * 32 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

// test / profiling only: see k_umma_pace.  cfg = {n, a_off_bytes, a_lbo, a_sbo, b_lbo, b_sbo, layout, reps, n_acc, a_step_bytes, grid}
extern "C" int modsgpu_debug_umma_pace(modsgpu_ctx* ctx, const int* cfg, double* cycles_per_mma) {
  if (!ctx || !cfg || !cycles_per_mma) return MODSGPU_EINVAL;
  const int n = cfg[0], reps = cfg[7], n_acc = cfg[8], grid = cfg[10];
  if (n < 16 || n > 256 || n % 16 || reps < 1 || reps > (1 << 20) || n_acc < 1 || n_acc * n > 512 || grid < 1 || grid > 148 ||
      cfg[1] < -4096 || cfg[1] > 16384 || (cfg[1] & 15))
    MG_FAIL(ctx, MODSGPU_EINVAL, "umma_pace: bad configuration");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  static OnceFlags attr;
  if (attr.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_umma_pace, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    attr.set(ctx->device);
  }
  MG_CUDA(ctx, ctx->io_c.ensure(148 * 8));
  k_umma_pace<<<grid, 128, 131072, ctx->stream>>>(n, cfg[1], cfg[2], cfg[3], cfg[4], cfg[5], cfg[6], reps, n_acc, cfg[9],
                                                  ctx->io_c.as<long long>());
  MG_LAUNCHED(ctx);
  std::vector<long long> h(grid);
  MG_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->io_c.p, grid * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  *cycles_per_mma = (double)mx / reps;
  return 0;
}
