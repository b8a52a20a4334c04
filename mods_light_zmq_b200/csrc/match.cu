// match.cu -- brute-force descriptor matching with the first-geometrically-inconsistent ratio test
// (seam S3).
//
// Replaces MatchFlannFGINN (matching.cpp:356-460) for vector_matcher=linear and DuplicateFiltering
// (matching.cpp:2615-2679).  Both reference descriptors (HardNet++ bytes, RootSIFT) are integers in
// [0,255], so  d2(i,j) = |q_i|^2 + |t_j|^2 - 2 q_i.t_j  is an exact integer in fp32 when the dot product
// runs on the tensor cores in fp16 x fp16 -> fp32 (255^2*128 < 2^24): neighbour lists are bit-identical
// to the CPU linear index, ties broken by the lower train index like cvflann's KNNSimpleResultSet.
//
//   k_pack_desc   fp32 rows -> fp16 planes [dim/8][rows][8] + squared norms (+ integrality check)
//   k_dist_umma   128 x 128 distance tiles: bulk-copy operands, tcgen05.mma, epilogue from TMEM
//   k_select_fginn per query: exact 3-pass radix select of the nn smallest (dist, idx), rank sort,
//                 then the FGINN walk (matching.cpp:430-457)
//   k_dup_filter  sort by ratio + greedy 2-px suppression, one CTA
#include "common.cuh"
#include "umma.cuh"
#include <cmath>
#include <algorithm>

using namespace umma;

namespace {

__global__ void k_pack_desc(const float* __restrict__ src, int rows, int rows_pad, int dim,
                            __half* __restrict__ planes, float* __restrict__ norms, int* __restrict__ bad, float pad_norm) {
  const int r = blockIdx.x * blockDim.y + threadIdx.y;   // one warp per row
  if (r >= rows_pad) return;
  const int lane = threadIdx.x;
  float ss = 0.f;
  for (int c8 = lane; c8 < dim / 8; c8 += 32) {
    __align__(16) __half h[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      float v = r < rows ? src[(size_t)r * dim + c8 * 8 + e] : 0.f;
      if (!(v >= 0.f && v <= 255.f && v == floorf(v))) atomicOr(bad, 1);
      h[e] = __float2half_rn(v);
      ss += v * v;
    }
    *reinterpret_cast<uint4*>(planes + ((size_t)c8 * rows_pad + r) * 8) = *reinterpret_cast<const uint4*>(h);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) norms[r] = r < rows ? ss : pad_norm;
}

// One CTA = one 128 (queries) x 128 (train) tile.  dim <= 256.
__global__ void __launch_bounds__(128)
k_dist_umma(const __half* __restrict__ qp, int nq_pad, const __half* __restrict__ tp, int nt_pad, int dim,
            const float* __restrict__ qn, const float* __restrict__ tn, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int npl = dim / 8;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + npl * 2048;
  float* tn_s = reinterpret_cast<float*>(smem + 2 * npl * 2048);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tn_s + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role branch
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tn_s[threadIdx.x] = tn[n0 + threadIdx.x];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  if (warp == 0) {
    // the warp waits together; one elected lane issues copies and MMAs (operands stay in uniform registers)
    if (elect_one()) {
      mbar_expect_tx(bars, 2 * npl * 2048);
      for (int pl = 0; pl < npl; pl++) {
        bulk_g2s(a_s + pl * 2048, qp + ((size_t)pl * nq_pad + m0) * 8, 2048, bars);
        bulk_g2s(b_s + pl * 2048, tp + ((size_t)pl * nt_pad + n0) * 8, 2048, bars);
      }
    }
    __syncwarp();
    mbar_wait(bars, 0);
    fence_after_sync();
    constexpr uint32_t idesc = instr_desc_f16(128);
    const uint64_t da = smem_desc(smem_u32(a_s), 2048, 128), db = smem_desc(smem_u32(b_s), 2048, 128);
    if (elect_one()) {
      for (int ks = 0; ks < dim / 16; ks++)
        mma_f16(tb, da + (uint64_t)(ks * 256), db + (uint64_t)(ks * 256), idesc, ks != 0);
      mma_commit(bars + 1);
    }
    __syncwarp();
  }
  mbar_wait(bars + 1, 0);
  fence_after_sync();
  const int m = warp * 32 + lane;
  const float qnv = qn[m0 + m];
  float* drow = D + (size_t)(m0 + m) * nt_pad + n0;
#pragma unroll 1
  for (int cc = 0; cc < 8; cc++) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + cc * 16, v);
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
      float4 o;
      o.x = (qnv + tn_s[cc * 16 + e + 0]) - 2.f * v[e + 0];
      o.y = (qnv + tn_s[cc * 16 + e + 1]) - 2.f * v[e + 1];
      o.z = (qnv + tn_s[cc * 16 + e + 2]) - 2.f * v[e + 2];
      o.w = (qnv + tn_s[cc * 16 + e + 3]) - 2.f * v[e + 3];
      *reinterpret_cast<float4*>(drow + cc * 16 + e) = o;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tb);
}

constexpr int SEL_CAP = 1024;

// Exact top-nn of one distance row + FGINN walk.  One CTA (128 threads) per query.
__global__ void __launch_bounds__(128)
k_select_fginn(const float* __restrict__ D, int nt, int nt_pad, int nn, const double* __restrict__ txy,
               double sqminratio, double contrDistSq, modsgpu_match* __restrict__ matches,
               int* __restrict__ knn_idx, float* __restrict__ knn_dist, int q_base) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long cand[SEL_CAP];
  __shared__ unsigned long long top[64];
  __shared__ unsigned s_prefix, s_remaining, s_count;
  const int q = blockIdx.x, tid = threadIdx.x;
  const float* row = D + (size_t)q * nt_pad;
  const int k = min(nn, nt);
  // distances are non-negative integers < 2^24 held in fp32: radix select on the integer value
  if (tid == 0) { s_prefix = 0; s_remaining = k; }
  for (int pass = 0; pass < 3; pass++) {
    const int shift = 16 - 8 * pass;
    for (int i = tid; i < 256; i += 128) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned hmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int j = tid; j < nt; j += 128) {
      unsigned v = (unsigned)row[j];
      if ((v & hmask) == prefix) atomicAdd(&hist[(v >> shift) & 255], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned rem = s_remaining, acc = 0;
      int b = 0;
      for (; b < 256; b++) { if (acc + hist[b] >= rem) break; acc += hist[b]; }
      s_prefix = prefix | ((unsigned)b << shift);
      s_remaining = rem - acc;
    }
    __syncthreads();
  }
  const unsigned V = s_prefix;   // value of the k-th smallest distance
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int j = tid; j < nt; j += 128) {
    unsigned v = (unsigned)row[j];
    if (v <= V) {
      unsigned slot = atomicAdd(&s_count, 1u);
      if (slot < SEL_CAP) cand[slot] = ((unsigned long long)v << 32) | (unsigned)j;
    }
  }
  __syncthreads();
  const unsigned cnt = s_count;
  if (cnt <= SEL_CAP) {
    // rank sort of the candidates by (dist, idx); the first k ranks are the neighbour list
    for (unsigned i = tid; i < cnt; i += 128) {
      unsigned long long me = cand[i];
      unsigned rank = 0;
      for (unsigned j = 0; j < cnt; j++) rank += cand[j] < me;
      if (rank < (unsigned)k) top[rank] = me;
    }
  } else if (tid == 0) {
    // pathological tie mass at V: serial, index-ordered (exactly the linear index's insertion order)
    int filled = 0;
    // strictly smaller first (at most k-1 of them), then equals in index order
    unsigned long long tmp[64];
    for (int j = 0; j < nt; j++) { unsigned v = (unsigned)row[j]; if (v < V) tmp[filled++] = ((unsigned long long)v << 32) | (unsigned)j; }
    for (int a = 1; a < filled; a++) { unsigned long long x = tmp[a]; int b2 = a - 1; while (b2 >= 0 && tmp[b2] > x) { tmp[b2 + 1] = tmp[b2]; b2--; } tmp[b2 + 1] = x; }
    for (int j = 0; j < nt && filled < k; j++) { unsigned v = (unsigned)row[j]; if (v == V) tmp[filled++] = ((unsigned long long)v << 32) | (unsigned)j; }
    for (int a = 0; a < k; a++) top[a] = tmp[a];
  }
  __syncthreads();
  if (knn_idx) {
    for (int i = tid; i < nn; i += 128) {
      knn_idx[(size_t)(q_base + q) * nn + i] = i < k ? (int)(top[i] & 0xffffffffu) : -1;
      knn_dist[(size_t)(q_base + q) * nn + i] = i < k ? (float)(unsigned)(top[i] >> 32) : INFINITY;
    }
  }
  if (tid == 0) {
    modsgpu_match mt;
    mt.qi = -1; mt.ti = -1; mt.tj_bad = -1; mt.d1 = 0.f; mt.d2 = 0.f; mt._pad = 0; mt.ratio = 0.0;
    if (k >= 1) {
      const int i0 = (int)(top[0] & 0xffffffffu);
      const float d0 = (float)(unsigned)(top[0] >> 32);
      for (int j = 1; j < k; j++) {
        const int ij = (int)(top[j] & 0xffffffffu);
        const float dj = (float)(unsigned)(top[j] >> 32);
        const double ratio = (double)(d0 / dj);
        const double dx = txy[2 * i0] - txy[2 * ij], dy = txy[2 * i0 + 1] - txy[2 * ij + 1];
        const bool contradictive = dx * dx + dy * dy > contrDistSq;
        if (sqminratio >= 1.0) {
          // matching.cpp:395-428 ("to get all points"): every query yields a correspondence -- with its first
          // geometrically inconsistent neighbour, or with the last neighbour of the list
          if (j == nn - 1 || contradictive) {
            mt.qi = q_base + q; mt.ti = i0; mt.tj_bad = ij; mt.d1 = d0; mt.d2 = dj; mt.ratio = sqrt(ratio);
            break;
          }
          continue;
        }
        if (ratio <= sqminratio) {
          mt.qi = q_base + q; mt.ti = i0; mt.tj_bad = ij; mt.d1 = d0; mt.d2 = dj; mt.ratio = sqrt(ratio);
          break;
        }
        if (contradictive) break;
      }
    }
    matches[q_base + q] = mt;
  }
}

// DuplicateFiltering (matching.cpp:2615-2679, mode bestFGINN): stable sort by |ratio|, then greedy
// suppression of later correspondences whose both endpoints lie within r of a kept one.
__global__ void __launch_bounds__(256)
k_dup_filter(const double* __restrict__ xy1, const double* __restrict__ xy2, const double* __restrict__ ratio, int T,
             double r_sq, int* __restrict__ ord, unsigned char* __restrict__ alive, int* __restrict__ out, int* __restrict__ nout) {
  const int tid = threadIdx.x;
  for (int i = tid; i < T; i += 256) {
    const double me = fabs(ratio[i]);
    int rank = 0;
    for (int j = 0; j < T; j++) { double o = fabs(ratio[j]); rank += (o < me) || (o == me && j < i); }
    ord[rank] = i;
    alive[i] = 1;
  }
  __syncthreads();
  for (int i = 0; i < T; i++) {
    if (alive[i]) {   // uniform: alive[i] is final once all smaller indices have been processed
      const int a = ord[i];
      const double ax1 = xy1[2 * a], ay1 = xy1[2 * a + 1], ax2 = xy2[2 * a], ay2 = xy2[2 * a + 1];
      for (int j = i + 1 + tid; j < T; j += 256) {
        if (!alive[j]) continue;
        const int b = ord[j];
        double dx = ax1 - xy1[2 * b], dy = ay1 - xy1[2 * b + 1];
        if (dx * dx + dy * dy > r_sq) continue;
        dx = ax2 - xy2[2 * b]; dy = ay2 - xy2[2 * b + 1];
        if (dx * dx + dy * dy <= r_sq) alive[j] = 0;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    int m = 0;
    for (int i = 0; i < T; i++) if (alive[i]) out[m++] = ord[i];
    *nout = m;
  }
}

// ---- duplicate filter, parallel form: (1) rank sort, (2) T x T conflict bit matrix between sorted
// positions a < b, (3) one warp walks the rows that have conflicts in order and clears the later bits.
// Identical result to the sequential greedy loop: a correspondence dies iff an earlier SURVIVOR conflicts.
constexpr int DUP_MAX_T = 16384;
// (1) rank of every correspondence in the order (|ratio| ascending, index ascending) and the coordinates in that order.
// 64 correspondences per CTA, four lanes each, the keys staged through shared memory in tiles of 256 (one thread per
// correspondence walking all T keys in global memory was 87 us for T = 3000: a chain of dependent loads).
__global__ void __launch_bounds__(256)
k_dup_rank(const double* __restrict__ ratio, const double* __restrict__ xy1, const double* __restrict__ xy2, int T,
           const int* __restrict__ Tp, int* __restrict__ ord, double* __restrict__ sxy) {
  __shared__ double tile[256];
  if (Tp != nullptr) { T = *Tp; if ((int)blockIdx.x * 64 >= T) return; }     // T known on the device only: grid sized by its bound
  const int i = blockIdx.x * 64 + (threadIdx.x >> 2), q = threadIdx.x & 3;
  const double me = i < T ? fabs(ratio[i]) : 0.0;
  int rank = 0;
  for (int base = 0; base < T; base += 256) {
    const int j = base + threadIdx.x;
    tile[threadIdx.x] = j < T ? fabs(ratio[j]) : 0.0;
    __syncthreads();
    const int n = min(256, T - base);
#pragma unroll 16
    for (int t = 0; t < 64; t++) {
      const int idx = 4 * t + q;
      const double o = tile[idx];
      rank += (idx < n) && ((o < me) || (o == me && base + idx < i));
    }
    __syncthreads();
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  if (q == 0 && i < T) {
    ord[rank] = i;
    sxy[4 * (size_t)rank + 0] = xy1[2 * i]; sxy[4 * (size_t)rank + 1] = xy1[2 * i + 1];
    sxy[4 * (size_t)rank + 2] = xy2[2 * i]; sxy[4 * (size_t)rank + 3] = xy2[2 * i + 1];
  }
}
// (2) conflict bit matrix between sorted positions a < b: a warp forms one 32-bit word with a ballot (lane = b), reading
// the sorted coordinates coalesced.  Only the words a row's walk reads (w >= a / 32) are written.
__global__ void __launch_bounds__(256)
k_dup_conflicts(const double* __restrict__ sxy, int T, const int* __restrict__ Tp, int nwords, double r_sq, unsigned* __restrict__ conf,
                int* __restrict__ rowflag) {
  // nwords = row pitch of conf (from the bound of T when T lives on the device)
  const int a = blockIdx.y, wi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (Tp != nullptr) T = *Tp;
  if (a >= T || wi >= (T + 31) / 32 || wi < (a >> 5)) return;
  const int b = wi * 32 + lane;
  bool hit = false;
  if (b > a && b < T) {
    const double ax1 = sxy[4 * (size_t)a], ay1 = sxy[4 * (size_t)a + 1], ax2 = sxy[4 * (size_t)a + 2], ay2 = sxy[4 * (size_t)a + 3];
    double dx = ax1 - sxy[4 * (size_t)b], dy = ay1 - sxy[4 * (size_t)b + 1];
    if (!(dx * dx + dy * dy > r_sq)) {
      dx = ax2 - sxy[4 * (size_t)b + 2]; dy = ay2 - sxy[4 * (size_t)b + 3];
      hit = dx * dx + dy * dy <= r_sq;
    }
  }
  const unsigned bits = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) {
    conf[(size_t)a * nwords + wi] = bits;      // nwords = pitch
    if (bits) rowflag[a] = 1;
  }
}
// (3) one warp walks the rows that have conflicts in order and clears the later bits; a row acts only while it is still
// alive.  The rows it will need are copied into shared memory by the whole CTA first (as many as fit), so the walk is
// not one L2 round trip per flagged row.
__global__ void __launch_bounds__(256)
k_dup_resolve(const unsigned* __restrict__ conf, const int* __restrict__ rowflag, const int* __restrict__ ord, int T,
              const int* __restrict__ Tp, int pitch, int cap_rows, int* __restrict__ out, int* __restrict__ nout) {
  extern __shared__ unsigned rows_s[];                 // cap_rows x pitch
  if (Tp != nullptr) T = *Tp;
  const int nwords = (T + 31) / 32;
  __shared__ unsigned alive[DUP_MAX_T / 32];
  __shared__ unsigned flagged[DUP_MAX_T / 32];
  __shared__ unsigned short slot_base[DUP_MAX_T / 32];    // staged rows before word w
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = warp; w < nwords; w += 8) {
    const int a = w * 32 + lane;
    const unsigned f = __ballot_sync(0xffffffffu, a < T && rowflag[a] != 0);
    if (lane == 0) { flagged[w] = f; alive[w] = 0xffffffffu; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int w = 0; w < nwords; w++) { slot_base[w] = (unsigned short)min(n, 65535); n += __popc(flagged[w]); }
  }
  __syncthreads();
  for (int w = warp; w < nwords; w += 8) {             // stage the flagged rows (a warp per word of the flag mask)
    unsigned m = flagged[w];
    int k = slot_base[w];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      if (k >= cap_rows) break;
      const unsigned* row = conf + (size_t)(w * 32 + bit) * pitch;
      for (int x = w + lane; x < nwords; x += 32) rows_s[(size_t)k * pitch + x] = row[x];
      k++;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  for (int wa = 0; wa < nwords; wa++) {
    unsigned m = flagged[wa];
    int k = slot_base[wa];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int a = wa * 32 + bit;
      const unsigned* row = k < cap_rows ? rows_s + (size_t)k * pitch : conf + (size_t)a * pitch;
      k++;
      if (!((alive[wa] >> bit) & 1u)) continue;
      for (int w = wa + lane; w < nwords; w += 32) alive[w] &= ~row[w];
      __syncwarp();
    }
  }
  int m = 0;
  for (int base = 0; base < T; base += 32) {
    const int a = base + lane;
    const bool keep = a < T && ((alive[a >> 5] >> (a & 31)) & 1u);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out[m + __popc(mask & ((1u << lane) - 1))] = ord[a];
    m += __popc(mask);
  }
  if (lane == 0) *nout = m;
}

// Matches in query order without the holes (qi < 0), and what the duplicate filter reads of them: the two positions and
// the ratio.  One CTA, order preserving (ballot + running base), count left on the device.
__global__ void __launch_bounds__(1024)
k_match_compact(const modsgpu_match* __restrict__ m, int nq, const double* __restrict__ qxy, const double* __restrict__ txy,
                modsgpu_match* __restrict__ mc, double* __restrict__ xy1, double* __restrict__ xy2, double* __restrict__ ratio,
                int* __restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nq; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    modsgpu_match r;
    r.qi = -1;
    if (i < nq) r = m[i];
    const bool keep = r.qi >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; w++) off += wsum[w];
    if (keep) {
      const int k = off + __popc(bal & ((1u << lane) - 1));
      mc[k] = r;
      xy1[2 * k] = qxy[2 * r.qi]; xy1[2 * k + 1] = qxy[2 * r.qi + 1];
      xy2[2 * k] = txy[2 * r.ti]; xy2[2 * k + 1] = txy[2 * r.ti + 1];
      ratio[k] = r.ratio;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; w++) t += wsum[w]; base_s += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base_s;
}

constexpr int RESOLVE_SMEM = 160 * 1024;   // rows of the conflict matrix staged in shared memory by k_dup_resolve

// the three launches of the parallel duplicate filter over T (host value) or *Tp <= T (device value) correspondences
static int dup_enqueue(modsgpu_ctx* ctx, const double* d1, const double* d2, const double* dr, int T, const int* Tp, double r,
                       int* dord, int* dout, int* dnout) {
  const int nwords = (T + 31) / 32;
  MG_CUDA(ctx, ctx->mt_d.ensure((size_t)T * nwords * 4 + (size_t)T * 4 + 16 + (size_t)T * 32));
  unsigned* conf = ctx->mt_d.as<unsigned>();
  int* rowflag = reinterpret_cast<int*>(conf + (size_t)T * nwords);
  double* sxy = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(ctx->mt_d.p) + (((size_t)T * nwords * 4 + (size_t)T * 4 + 15) & ~(size_t)15));
  MG_CUDA(ctx, cudaMemsetAsync(rowflag, 0, (size_t)T * 4, ctx->stream));
  MG_PROF(ctx, "k_dup_rank", 2, (double)T);
  k_dup_rank<<<(T + 63) / 64, 256, 0, ctx->stream>>>(dr, d1, d2, T, Tp, dord, sxy);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_dup_conflicts", 2, (double)T);
  k_dup_conflicts<<<dim3((nwords + 7) / 8, T), 256, 0, ctx->stream>>>(sxy, T, Tp, nwords, r * r, conf, rowflag);
  MG_LAUNCHED(ctx);
  static OnceFlags attr_set;
  if (attr_set.need(ctx->device)) {
    MG_CUDA(ctx, cudaFuncSetAttribute(k_dup_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, RESOLVE_SMEM));
    attr_set.set(ctx->device);
  }
  const int cap_rows = std::min(T, RESOLVE_SMEM / (nwords * 4));
  MG_PROF(ctx, "k_dup_resolve", 2, (double)T);
  k_dup_resolve<<<1, 256, (size_t)cap_rows * nwords * 4, ctx->stream>>>(conf, rowflag, dord, T, Tp, nwords, cap_rows, dout, dnout);
  MG_LAUNCHED(ctx);
  return 0;
}

}  // namespace

// Device part of the matcher.  d_q / d_t: fp32 descriptor rows already on the device, d_txy doubles.
// Results: ctx->mt_out holds nq modsgpu_match records (qi = -1 when the query produced no tentative).
int mg_match_enqueue(modsgpu_ctx* ctx, const float* d_q, int nq, const float* d_t, const double* d_txy, int nt, int dim,
                     double ratio_thr, double contrad_dist, int nn, int* d_knn_idx, float* d_knn_dist) {
  // 255^2 * 2 * dim must stay below 2^24: the fp32 epilogue (qn + tn - 2 dot) and the 24-bit radix select are exact
  // integers only up to dim = 128 (both reference descriptors are 128-d)
  if (dim % 16 != 0 || dim < 16 || dim > 128) MG_FAIL(ctx, MODSGPU_EINVAL, "descriptor dim must be a multiple of 16, <= 128");
  if (nn < 1 || nn > 64) MG_FAIL(ctx, MODSGPU_EINVAL, "nn must be in [1,64]");
  const int nq_pad = (nq + 127) / 128 * 128, nt_pad = (nt + 127) / 128 * 128, npl = dim / 8;
  // operand planes + norms + flag
  size_t q_bytes = (size_t)npl * nq_pad * 16, t_bytes = (size_t)npl * nt_pad * 16;
  MG_CUDA(ctx, ctx->mt_q.ensure(q_bytes + (size_t)nq_pad * 4));
  MG_CUDA(ctx, ctx->mt_t.ensure(t_bytes + (size_t)nt_pad * 4));
  MG_CUDA(ctx, ctx->mt_aux.ensure(64));
  MG_CUDA(ctx, ctx->mt_out.ensure((size_t)std::max(nq, 1) * sizeof(modsgpu_match)));
  __half* qp = ctx->mt_q.as<__half>();
  float* qn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->mt_q.p) + q_bytes);
  __half* tp = ctx->mt_t.as<__half>();
  float* tn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->mt_t.p) + t_bytes);
  int* bad = ctx->mt_aux.as<int>();
  MG_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
  dim3 pb(32, 8);
  MG_PROF(ctx, "k_pack_desc", 0, (double)nq * dim * 6.0);
  k_pack_desc<<<nq_pad / 8, pb, 0, ctx->stream>>>(d_q, nq, nq_pad, dim, qp, qn, bad, 0.f);
  MG_LAUNCHED(ctx);
  MG_PROF(ctx, "k_pack_desc", 0, (double)nt * dim * 6.0);
  k_pack_desc<<<nt_pad / 8, pb, 0, ctx->stream>>>(d_t, nt, nt_pad, dim, tp, tn, bad, 0.f);
  MG_LAUNCHED(ctx);
  // distance matrix in query blocks of <= 256 MB
  int qblk = (int)std::min<size_t>((size_t)nq_pad, std::max<size_t>(128, ((size_t)256 << 20) / ((size_t)nt_pad * 4) / 128 * 128));
  MG_CUDA(ctx, ctx->mt_d.ensure((size_t)qblk * nt_pad * 4));
  const int smem = 2 * npl * 2048 + 512 + 64;
  static OnceFlags attr;
  if (attr.need(ctx->device)) { MG_CUDA(ctx, cudaFuncSetAttribute(k_dist_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32 * 2048 + 512 + 64)); attr.set(ctx->device); }
  const double sq = ratio_thr * ratio_thr, cd = contrad_dist * contrad_dist;
  for (int q0 = 0; q0 < nq; q0 += qblk) {
    const int rows_pad = std::min(qblk, nq_pad - q0), rows = std::min(qblk, nq - q0);
    dim3 grid(nt_pad / 128, rows_pad / 128);
    MG_PROF(ctx, "k_dist_umma", 0, (double)rows_pad * nt_pad * 4.0 + ((double)rows_pad + nt_pad) * dim * 2.0);
    k_dist_umma<<<grid, 128, smem, ctx->stream>>>(qp + (size_t)q0 * 8, nq_pad, tp, nt_pad, dim, qn + q0, tn, ctx->mt_d.as<float>());
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_select_fginn", 0, (double)rows * nt * 4.0);
    k_select_fginn<<<rows, 128, 0, ctx->stream>>>(ctx->mt_d.as<float>(), nt, nt_pad, nn, d_txy, sq, cd,
                                                   ctx->mt_out.as<modsgpu_match>(), d_knn_idx, d_knn_dist, q0);
    MG_LAUNCHED(ctx);
  }
  return 0;
}

extern "C" int modsgpu_match_fginn(modsgpu_ctx* ctx, const float* q, int nq, const float* t, const double* txy, int nt,
                                   int dim, double ratio_thr, double contrad_dist, int nn,
                                   modsgpu_match* out, int* nout, int* knn_idx, float* knn_dist) {
  if (!ctx || !nout || nq < 0 || nt < 0 || (nq > 0 && (!q || !out)) || (nt > 0 && (!t || !txy))) return MODSGPU_EINVAL;
  *nout = 0;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (nq == 0 || nt == 0) {   // matching.cpp:361-366: nothing to match
    if (knn_idx) for (size_t i = 0; i < (size_t)nq * nn; i++) { knn_idx[i] = -1; knn_dist[i] = INFINITY; }
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  }
  size_t qb = (size_t)nq * dim * 4, tb = (size_t)nt * dim * 4, xb = (size_t)nt * 16;
  MG_CUDA(ctx, ctx->io_a.ensure(qb));
  MG_CUDA(ctx, ctx->io_b.ensure(tb));
  MG_CUDA(ctx, ctx->io_c.ensure(xb + (knn_idx ? (size_t)nq * nn * 8 : 0)));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, q, qb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, t, tb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_c.p, txy, xb, cudaMemcpyHostToDevice, ctx->stream));
  int* d_ki = knn_idx ? reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ctx->io_c.p) + xb) : nullptr;
  float* d_kd = knn_idx ? reinterpret_cast<float*>(d_ki + (size_t)nq * nn) : nullptr;
  int rc = mg_match_enqueue(ctx, ctx->io_a.as<float>(), nq, ctx->io_b.as<float>(), ctx->io_c.as<double>(), nt, dim,
                            ratio_thr, contrad_dist, nn, d_ki, d_kd);
  if (rc) return rc;
  MG_CUDA(ctx, ctx->h_stage.ensure((size_t)nq * sizeof(modsgpu_match) + 64));
  modsgpu_match* hm = ctx->h_stage.as<modsgpu_match>();
  int* hbad = reinterpret_cast<int*>(hm + nq);
  MG_CUDA(ctx, cudaMemcpyAsync(hm, ctx->mt_out.p, (size_t)nq * sizeof(modsgpu_match), cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(hbad, ctx->mt_aux.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (knn_idx) {
    MG_CUDA(ctx, cudaMemcpyAsync(knn_idx, d_ki, (size_t)nq * nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, cudaMemcpyAsync(knn_dist, d_kd, (size_t)nq * nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  if (*hbad) MG_FAIL(ctx, MODSGPU_EINVAL, "descriptors must hold integers in [0,255] (HardNet++ bytes / RootSIFT)");
  int m = 0;
  for (int i = 0; i < nq; i++) if (hm[i].qi >= 0) out[m++] = hm[i];   // query order, like TCList (matching.cpp:449)
  *nout = m;
  return 0;
}

// MatchFlannFGINN over descriptor blocks that never left the device (modsgpu_describe_view_dev): only the train
// coordinates go up and the matches come down.  Same kernels, same results as modsgpu_match_fginn.
extern "C" int modsgpu_devdesc_size(const modsgpu_devdesc* dd);
extern "C" int modsgpu_match_fginn_dev(modsgpu_ctx* ctx, const modsgpu_devdesc* q, const modsgpu_devdesc* t, const double* txy,
                                       double ratio_thr, double contrad_dist, int nn, modsgpu_match* out, int* nout) {
  if (!ctx || !nout) return MODSGPU_EINVAL;
  *nout = 0;
  const int nq = modsgpu_devdesc_size(q), nt = modsgpu_devdesc_size(t), dim = 128;
  if ((nq > 0 && !out) || (nt > 0 && !txy)) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (nq == 0 || nt == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const size_t xb = (size_t)nt * 16;
  MG_CUDA(ctx, ctx->io_c.ensure(xb));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_c.p, txy, xb, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_match_enqueue(ctx, mg_devdesc_ptr(q), nq, mg_devdesc_ptr(t), ctx->io_c.as<double>(), nt, dim, ratio_thr, contrad_dist, nn,
                            nullptr, nullptr);
  if (rc) return rc;
  MG_CUDA(ctx, ctx->h_stage.ensure((size_t)nq * sizeof(modsgpu_match) + 64));
  modsgpu_match* hm = ctx->h_stage.as<modsgpu_match>();
  int* hbad = reinterpret_cast<int*>(hm + nq);
  MG_CUDA(ctx, cudaMemcpyAsync(hm, ctx->mt_out.p, (size_t)nq * sizeof(modsgpu_match), cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(hbad, ctx->mt_aux.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  if (*hbad) MG_FAIL(ctx, MODSGPU_EINVAL, "descriptors must hold integers in [0,255] (HardNet++ bytes / RootSIFT)");
  int m = 0;
  for (int i = 0; i < nq; i++) if (hm[i].qi >= 0) out[m++] = hm[i];   // query order, like TCList (matching.cpp:449)
  *nout = m;
  return 0;
}

// MatchFlannFGINN + DuplicateFiltering (matching.cpp:356-460, :2615-2679) in one call over device-resident descriptor
// blocks: the tentative list never visits the host between the two -- it is compacted on the device (k_match_compact),
// the duplicate filter reads its positions / ratios there and takes its length from device memory; ONE read-back brings the
// matches (query order) and the order of the survivors.  Results equal modsgpu_match_fginn_dev + modsgpu_duplicate_filter.
extern "C" int modsgpu_match_dedup_dev(modsgpu_ctx* ctx, const modsgpu_devdesc* q, const modsgpu_devdesc* t, const double* qxy,
                                       const double* txy, double ratio_thr, double contrad_dist, int nn, double dup_radius,
                                       modsgpu_match* matches, int* n_matches, int* order, int* n_unique) {
  if (!ctx || !n_matches || !n_unique) return MODSGPU_EINVAL;
  *n_matches = 0; *n_unique = 0;
  const int nq = modsgpu_devdesc_size(q), nt = modsgpu_devdesc_size(t), dim = 128;
  if ((nq > 0 && (!matches || !order || !qxy)) || (nt > 0 && !txy)) return MODSGPU_EINVAL;
  if (nq > DUP_MAX_T) MG_FAIL(ctx, MODSGPU_EINVAL, "too many queries for the fused matcher + duplicate filter (use the two calls)");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (nq == 0 || nt == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  const size_t xb = (size_t)nt * 16, qb = (size_t)nq * 16;
  // io_c: [txy | qxy]; io_a: compact matches | xy1 | xy2 | ratio; io_b: ord | out
  MG_CUDA(ctx, ctx->io_c.ensure(xb + qb));
  MG_CUDA(ctx, ctx->io_a.ensure((size_t)nq * (sizeof(modsgpu_match) + 40) + 64));
  MG_CUDA(ctx, ctx->io_b.ensure((size_t)nq * 8 + 64));
  double* dtxy = ctx->io_c.as<double>();
  double* dqxy = dtxy + 2 * (size_t)nt;
  MG_CUDA(ctx, cudaMemcpyAsync(dtxy, txy, xb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(dqxy, qxy, qb, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_match_enqueue(ctx, mg_devdesc_ptr(q), nq, mg_devdesc_ptr(t), dtxy, nt, dim, ratio_thr, contrad_dist, nn, nullptr, nullptr);
  if (rc) return rc;
  modsgpu_match* mc = ctx->io_a.as<modsgpu_match>();
  double* d1 = reinterpret_cast<double*>(mc + nq);
  double* d2 = d1 + 2 * (size_t)nq;
  double* dr = d2 + 2 * (size_t)nq;
  int* dord = ctx->io_b.as<int>();
  int* dout = dord + nq;
  int* aux = ctx->mt_aux.as<int>();          // [0] bad-descriptor flag (matcher), [4] survivors, [5] tentatives
  MG_PROF(ctx, "k_match_compact", 2, (double)nq);
  k_match_compact<<<1, 1024, 0, ctx->stream>>>(ctx->mt_out.as<modsgpu_match>(), nq, dqxy, dtxy, mc, d1, d2, dr, aux + 5);
  MG_LAUNCHED(ctx);
  const bool dedup = dup_radius > 0;          // matching.cpp:2621-2625: a radius <= 0 disables the filter
  if (dedup && (rc = dup_enqueue(ctx, d1, d2, dr, nq, aux + 5, dup_radius, dord, dout, aux + 4))) return rc;
  MG_CUDA(ctx, ctx->h_stage.ensure((size_t)nq * (sizeof(modsgpu_match) + 4) + 64));
  modsgpu_match* hm = ctx->h_stage.as<modsgpu_match>();
  int* ho = reinterpret_cast<int*>(hm + nq);
  int* hcnt = ho + nq;                        // [0] bad flag, [4] survivors, [5] tentatives
  MG_CUDA(ctx, cudaMemcpyAsync(hm, mc, (size_t)nq * sizeof(modsgpu_match), cudaMemcpyDeviceToHost, ctx->stream));
  if (dedup) MG_CUDA(ctx, cudaMemcpyAsync(ho, dout, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(hcnt, aux, 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  if (hcnt[0]) MG_FAIL(ctx, MODSGPU_EINVAL, "descriptors must hold integers in [0,255] (HardNet++ bytes / RootSIFT)");
  const int T = hcnt[5];
  memcpy(matches, hm, (size_t)T * sizeof(modsgpu_match));
  *n_matches = T;
  if (dedup) {
    *n_unique = hcnt[4];
    memcpy(order, ho, (size_t)hcnt[4] * 4);
  } else {
    for (int i = 0; i < T; i++) order[i] = i;
    *n_unique = T;
  }
  return 0;
}

// =====================================================================================================================
// MatchFLANNDistance (matching.cpp:574-633): binary descriptors (bytes = floor of the float entries, :596-608), the 2
// nearest train descriptors by Hamming distance, a match whenever the nearest is within matchDistanceThreshold; ratio =
// d1 / d2.  Exact linear search (binary_matcher = linear): ties keep the lower train index first, as cvflann's LinearIndex +
// KNNSimpleResultSet do for every distance type (the reference's default hierarchical-clustering index is approximate and
// randomised -- not a parity target, like the kd-tree of the vector matcher).
// One CTA per query: a thread walks train rows j = tid, tid + 128, ..., keeps its two smallest (distance, index) keys,
// thread 0 merges the 256 keys.
// =====================================================================================================================
namespace {
__global__ void __launch_bounds__(128)
k_hamming_2nn(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t, int nt, int words, int max_distance,
              modsgpu_match* __restrict__ matches) {
  __shared__ uint32_t qs[32];
  __shared__ unsigned long long best[256];
  const int qi = blockIdx.x, tid = threadIdx.x;
  if (tid < words) qs[tid] = q[(size_t)qi * words + tid];
  __syncthreads();
  unsigned long long k1 = ~0ull, k2 = ~0ull;
  for (int j = tid; j < nt; j += 128) {
    const uint32_t* row = t + (size_t)j * words;
    unsigned d = 0;
    for (int w = 0; w < words; w++) d += __popc(qs[w] ^ row[w]);
    const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)j;
    if (key < k1) { k2 = k1; k1 = key; }
    else if (key < k2) k2 = key;
  }
  best[2 * tid] = k1; best[2 * tid + 1] = k2;
  __syncthreads();
  if (tid == 0) {
    unsigned long long a = ~0ull, b = ~0ull;
    for (int i = 0; i < 256; i++) {
      const unsigned long long key = best[i];
      if (key < a) { b = a; a = key; }
      else if (key < b) b = key;
    }
    modsgpu_match mt;
    mt.qi = -1; mt.ti = -1; mt.tj_bad = -1; mt.d1 = 0.f; mt.d2 = 0.f; mt._pad = 0; mt.ratio = 0.0;
    if (a != ~0ull && (int)(a >> 32) <= max_distance) {
      mt.qi = qi; mt.ti = (int)(a & 0xffffffffu); mt.d1 = (float)(unsigned)(a >> 32);
      // with a single train descriptor the reference reads an unset second distance; here d2 = 0 and ratio = inf
      if (b != ~0ull) { mt.tj_bad = (int)(b & 0xffffffffu); mt.d2 = (float)(unsigned)(b >> 32); }
      mt.ratio = (double)mt.d1 / (double)mt.d2;
    }
    matches[qi] = mt;
  }
}
}  // namespace

extern "C" int modsgpu_match_hamming(modsgpu_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                                     double max_distance, modsgpu_match* out, int* nout) {
  if (!ctx || !nout || nq < 0 || nt < 0 || (nq > 0 && (!q || !out)) || (nt > 0 && !t)) return MODSGPU_EINVAL;
  *nout = 0;
  if (dim < 4 || dim > 128 || dim % 4) MG_FAIL(ctx, MODSGPU_EINVAL, "binary descriptor length must be a multiple of 4 bytes, <= 128");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (nq == 0 || nt == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  // Row[j] = floor(desc.vec[j]) into unsigned char (matching.cpp:596-608)
  const size_t qb = (size_t)nq * dim, tb = (size_t)nt * dim;
  MG_CUDA(ctx, ctx->h_stage.ensure(qb + tb + (size_t)nq * sizeof(modsgpu_match) + 64));
  unsigned char* hq = ctx->h_stage.as<unsigned char>();
  unsigned char* ht = hq + ((qb + 15) & ~(size_t)15);
  for (size_t i = 0; i < qb; i++) hq[i] = (unsigned char)floorf(q[i]);
  for (size_t i = 0; i < tb; i++) ht[i] = (unsigned char)floorf(t[i]);
  MG_CUDA(ctx, ctx->io_a.ensure(qb + 16));
  MG_CUDA(ctx, ctx->io_b.ensure(tb + 16));
  MG_CUDA(ctx, ctx->mt_out.ensure((size_t)nq * sizeof(modsgpu_match)));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, hq, qb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, ht, tb, cudaMemcpyHostToDevice, ctx->stream));
  MG_PROF(ctx, "k_hamming_2nn", 2, (double)nq * nt);
  k_hamming_2nn<<<nq, 128, 0, ctx->stream>>>(ctx->io_a.as<uint32_t>(), ctx->io_b.as<uint32_t>(), nt, dim / 4, (int)(float)max_distance,
                                             ctx->mt_out.as<modsgpu_match>());
  MG_LAUNCHED(ctx);
  modsgpu_match* hm = reinterpret_cast<modsgpu_match*>(ht + ((tb + 15) & ~(size_t)15));
  MG_CUDA(ctx, cudaMemcpyAsync(hm, ctx->mt_out.p, (size_t)nq * sizeof(modsgpu_match), cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  int m = 0;
  for (int i = 0; i < nq; i++) if (hm[i].qi >= 0) out[m++] = hm[i];
  *nout = m;
  return 0;
}

extern "C" int modsgpu_duplicate_filter(modsgpu_ctx* ctx, const double* xy1, const double* xy2, const double* ratio,
                                        int T, double r, int* order_out, int* nout) {
  if (!ctx || !nout || T < 0 || (T > 0 && (!xy1 || !xy2 || !ratio || !order_out))) return MODSGPU_EINVAL;
  *nout = 0;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (T == 0) return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  if (r <= 0) {   // matching.cpp:2621-2625: filtering disabled
    for (int i = 0; i < T; i++) order_out[i] = i;
    *nout = T;
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  }
  size_t xb = (size_t)T * 16, rb = (size_t)T * 8;
  MG_CUDA(ctx, ctx->io_a.ensure(2 * xb + rb));
  MG_CUDA(ctx, ctx->io_b.ensure((size_t)T * 9 + 16));
  double* d1 = ctx->io_a.as<double>();
  double* d2 = d1 + 2 * T;
  double* dr = d2 + 2 * T;
  int* dord = ctx->io_b.as<int>();
  int* dout = dord + T;
  unsigned char* alive = reinterpret_cast<unsigned char*>(dout + T);
  MG_CUDA(ctx, ctx->mt_aux.ensure(64));
  MG_CUDA(ctx, cudaMemcpyAsync(d1, xy1, xb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(d2, xy2, xb, cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(dr, ratio, rb, cudaMemcpyHostToDevice, ctx->stream));
  if (T <= DUP_MAX_T) {
    if (int rc = dup_enqueue(ctx, d1, d2, dr, T, nullptr, r, dord, dout, ctx->mt_aux.as<int>() + 4)) return rc;
  } else {
    MG_PROF(ctx, "k_dup_filter", 2, (double)T);
    k_dup_filter<<<1, 256, 0, ctx->stream>>>(d1, d2, dr, T, r * r, dord, alive, dout, ctx->mt_aux.as<int>() + 4);
    MG_LAUNCHED(ctx);
  }
  MG_CUDA(ctx, ctx->h_stage.ensure((size_t)T * 4 + 16));
  int* h = ctx->h_stage.as<int>();
  MG_CUDA(ctx, cudaMemcpyAsync(h, dout, (size_t)T * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(h + T, ctx->mt_aux.as<int>() + 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  *nout = h[T];
  memcpy(order_out, h, (size_t)h[T] * 4);
  return 0;
}
