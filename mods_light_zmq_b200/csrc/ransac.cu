// ransac.cu -- batched LO-RANSAC for a homography on the device (seam S4).
//
// Replaces exp_ransacHcustom (degensac/exp_ranH.c:796-1236) as LORANSACFiltering calls it
// (matching.cpp:731: iter_type 4, oriented constraint on, MSAC score, symmetric check; error function HDs / HDsSym /
// HDsSymMax selected by RANSACPars::errorType, matching.cpp:652-681).  The reference is a sequential loop driven by
// libc rand(); this is the same estimator re-organised for a GPU:
//   k_rs_hyp     one WARP per hypothesis, B hypotheses per launch: counter-based RNG (no state carried
//                between samples) -> 4-point sample -> oriented constraint (Htools.c all_Hori_valid) ->
//                8x9 null space by pivoted Gauss-Jordan (utools.c:97-167) -> |det|/h33^3 test ->
//                error of all T correspondences, lanes striding over T -> MSAC score (rtools.c truncQuad,
//                9/4*th width) by a fixed-order butterfly reduction
//   k_rsb_select one CTA of 256 threads per batch: best hypothesis (max J, lowest index on ties), symmetric
//                transfer check (exp_ranH.c:905-947), bookkeeping of the sequential loop, and -- when the local
//                optimisation is due -- its start model (LSQ on the 8*th band) and inlier list
//   k_rsb_lo     10 CTAs of 256 threads: one inner sample each (exp_inHranicustom / exp_iterHcustom: LSQ on a random
//                subset of the inliers, then 4 shrinking-threshold LSQ steps), every T-long pass spread over the CTA
//   k_rsb_accept one CTA: best inner sample vs the best model, stopping rule nsamples(I+1,T,4,conf)
// The batch loop is on the device: every kernel starts with `if (st->done) return`, so the host enqueues batches
// speculatively and synchronises once per GROUP of batches (the first group, 512 + 1024 hypotheses plus the closing
// kernels, ends almost every real pair: one host synchronisation per call instead of one per batch).
// All arithmetic is fp64 and compiled with --fmad=false; every reduction has a fixed order (thread t of a CTA sums the
// items t, t+256, ... in ascending order; lanes combine by the xor butterfly 16,8,4,2,1; warps are added in order
// 0..7), so a run is reproducible from (u, params) and oracle/ransac_batched.c restates it bit for bit on the CPU.
// Deviations from the reference, on purpose: (1) samples are consumed in batches, so at least 512 are drawn;
// (2) the LSQ null vector comes from 10 steps of inverse iteration on the 9x9 normal matrix instead of
// LAPACK dsyev_ (lapwrap.c:75-97; third-party, unpinned); (3) the inlier-set hash that prunes repeated
// LO iterations (exp_ranH.c __HASHING__) is dropped -- it only skips work whose result is already known;
// (4) LO least squares on all band inliers (the reference, called with inlLimit = 0 under __D3__, refits on a random
// minimal subset).
#include "common.cuh"
#include "ransac_common.cuh"
#include "ransac_h.cuh"
#include <cmath>
#include <algorithm>

namespace {

constexpr int RS_NT = 256, RS_NWARP = RS_NT / 32;

struct RsState {
  double H[9];            // best model (maxS)
  double J; int I;
  double Hs[9];           // best sample so far (maxSs)
  double Js; int Is;
  int max_sam, no_sam, lo_runs, oc_rejects, done, have_sample;
};

struct HypOut { double H[9]; double J; int I; int flag; };   // flag: 0 ok, 1 oriented-constraint reject, 2 degenerate

// ---- the error function of a model (matching.cpp:652-681): Sampson HDs (Htools.c:160-198), symmetric transfer sum
//      HDsSym (:201-242) or maximum HDsSymMax (:243-284) --------------------------------------------------------------
struct HErr {
  double H[9], Hm[9], H1[9];
  int type, ok;
};
__device__ __forceinline__ void herr_setup(HErr& e, const double* H, int type) {
  for (int i = 0; i < 9; i++) e.H[i] = H[i];
  e.type = type; e.ok = 1;
  if (type != MODSGPU_ERR_SAMPSON) {
    const double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]};
    for (int i = 0; i < 9; i++) e.Hm[i] = Hm[i];
    e.ok = inv3(e.Hm, e.H1) ? 1 : 0;
  }
}
__device__ __forceinline__ double herr(const HErr& e, const double* u) {
  if (e.type == MODSGPU_ERR_SAMPSON) return sampson(e.H, u);
  if (!e.ok) return 1e300;
  const double a = e.H1[6] * u[0] + e.H1[7] * u[1] + e.H1[8];
  const double b = e.Hm[6] * u[3] + e.Hm[7] * u[4] + e.Hm[8];
  double xa = (e.H1[0] * u[0] + e.H1[1] * u[1] + e.H1[2]) / a, ya = (e.H1[3] * u[0] + e.H1[4] * u[1] + e.H1[5]) / a;
  double xd = u[3] - xa, yd = u[4] - ya;
  const double d1 = xd * xd + yd * yd;
  xa = (e.Hm[0] * u[3] + e.Hm[1] * u[4] + e.Hm[2]) / b; ya = (e.Hm[3] * u[3] + e.Hm[4] * u[4] + e.Hm[5]) / b;
  xd = u[0] - xa; yd = u[1] - ya;
  const double d2 = xd * xd + yd * yd;
  return e.type == MODSGPU_ERR_SYMM_SUM ? d1 + d2 : (d1 > d2 ? d1 : d2);
}

// ---- hypothesis generation + scoring: one warp per hypothesis -------------------------------------------
__global__ void __launch_bounds__(256)
k_rs_hyp(const double* __restrict__ u, int T, double th, int etype, unsigned long long seed, int base, int nhyp,
         const RsState* __restrict__ st, HypOut* __restrict__ out) {
  if (st->done) return;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nhyp) return;
  const unsigned long long hid = (unsigned long long)(base + w);
  int idx[4];
  draw_sample<4>(seed, hid, T, idx);
  double h[9];
  int flag = 0;
  if (!all_Hori_valid(u, idx)) flag = 1;
  else if (!h_from_4(u, idx, h) || !det_ok(h)) flag = 2;
  int I = 0;
  double J = 0;
  if (flag == 0) {
    HErr E;
    herr_setup(E, h, etype);
    int ci = 0; double cj = 0;
    for (int j = lane; j < T; j += 32) {
      const double e = herr(E, u + 6 * j);
      if (e <= th) ci++;
      cj += truncQuad(e, th);
    }
    I = warp_sum_i(ci); J = warp_sum_d(cj);
  }
  if (lane == 0) {
    HypOut o;
    for (int i = 0; i < 9; i++) o.H[i] = flag == 0 ? h[i] : 0.0;
    o.I = I; o.J = J; o.flag = flag;
    out[w] = o;
  }
}

// ---- block-wide building blocks (every thread of the 256-thread CTA calls them) -----------------------------
struct BlkScratch {
  double c[RS_NWARP][45];
  int ci[RS_NWARP];
  double bj[RS_NWARP]; int bi[RS_NWARP];
};

// sums of N per-thread values over the CTA, the same bits in every thread: lanes by the xor butterfly, warps in order
template <int N>
__device__ __forceinline__ void blk_sum_vec(double* v, BlkScratch& S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < N; i++) v[i] = warp_sum_d(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) S.c[warp][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; i++) {
    double s = S.c[0][i];
    for (int w = 1; w < RS_NWARP; w++) s += S.c[w][i];
    v[i] = s;
  }
}
__device__ __forceinline__ int blk_sum_i(int v, BlkScratch& S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum_i(v);
  __syncthreads();
  if (lane == 0) S.ci[warp] = v;
  __syncthreads();
  int s = 0;
  for (int w = 0; w < RS_NWARP; w++) s += S.ci[w];
  return s;
}

// errors of all T correspondences under E into d[] (optional) + inlier count and MSAC score at th
__device__ void blk_score(const double* __restrict__ u, int T, const HErr& E, double th, double* d, BlkScratch& S, int* I, double* J) {
  int ci = 0;
  double cj[1] = {0};
  for (int j = threadIdx.x; j < T; j += RS_NT) {
    const double e = herr(E, u + 6 * j);
    if (d) d[j] = e;
    if (e <= th) ci++;
    cj[0] += truncQuad(e, th);
  }
  *I = blk_sum_i(ci, S);
  blk_sum_vec<1>(cj, S);
  *J = cj[0];
  __syncthreads();     // d[] complete for every reader
}

// indices j (ascending) with d[j] <= th into idx[]; returns the count (rtools.c inlidxs)
__device__ int blk_compact(const double* d, int T, double th, int* idx, BlkScratch& S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n = 0;
  for (int base = 0; base < T; base += RS_NT) {
    const int j = base + threadIdx.x;
    const bool in = j < T && d[j] <= th;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    __syncthreads();
    if (lane == 0) S.ci[warp] = __popc(m);
    __syncthreads();
    int off = n, tot = 0;
    for (int w = 0; w < RS_NWARP; w++) { if (w < warp) off += S.ci[w]; tot += S.ci[w]; }
    if (in) idx[off + __popc(m & ((1u << lane) - 1))] = j;
    n += tot;
  }
  __syncthreads();
  return n;
}

// normalised-DLT least squares (u2h, Htools.c:100-132: normu + lin_hgN + cov_mat + smallest eigenvector + denormH) over
// the correspondences idx[0..n); H is every thread's private copy (all threads end with the same bits).  n < 4 leaves it.
__device__ void blk_lsq(const double* __restrict__ u, const int* idx, int n, double* H, BlkScratch& S) {
  if (n < 4) return;
  if (n == 4) {
    int id4[4] = {idx[0], idx[1], idx[2], idx[3]};
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = H[i];
    h_from_4(u, id4, h);   // like u2h's len == 4 branch; a degenerate sample keeps the previous H
    for (int i = 0; i < 9; i++) H[i] = h[i];
    return;
  }
  // normu (utools.c:7-50)
  double s4[4] = {0, 0, 0, 0};
  for (int k = threadIdx.x; k < n; k += RS_NT) { const double* p = u + 6 * idx[k]; s4[0] += p[0]; s4[1] += p[1]; s4[2] += p[3]; s4[3] += p[4]; }
  blk_sum_vec<4>(s4, S);
  const double m1x = s4[0] / n, m1y = s4[1] / n, m2x = s4[2] / n, m2y = s4[3] / n;
  double q[2] = {0, 0};
  for (int k = threadIdx.x; k < n; k += RS_NT) {
    const double* p = u + 6 * idx[k];
    double a = p[0] - m1x, b = p[1] - m1y;
    q[0] += sqrt(a * a + b * b);
    a = p[3] - m2x; b = p[4] - m2y;
    q[1] += sqrt(a * a + b * b);
  }
  blk_sum_vec<2>(q, S);
  double A1[3] = {q[0], m1x, m1y}, A2[3] = {q[1], m2x, m2y};
  if (A1[0] != 0) A1[0] = n * sqrt(2.0) / A1[0];
  if (A2[0] != 0) A2[0] = n * sqrt(2.0) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0]; A2[1] *= -A2[0]; A2[2] *= -A2[0];
  // normal matrix C = Z^T Z of the 2n x 9 design matrix (lin_hgN + cov_mat), packed lower triangle
  double C[45];
#pragma unroll
  for (int i = 0; i < 45; i++) C[i] = 0;
  for (int k = threadIdx.x; k < n; k += RS_NT) {
    const double* p = u + 6 * idx[k];
    const double a0 = p[0] * A1[0] + A1[1], a1 = p[1] * A1[0] + A1[2];
    const double b0 = p[3] * A2[0] + A2[1], b1 = p[4] * A2[0] + A2[2], b2 = 1;
    const double r1[9] = {b0, 0, -a0 * b0, b1, 0, -a0 * b1, b2, 0, -a0 * b2};
    const double r2[9] = {0, b0, -a1 * b0, 0, b1, -a1 * b1, 0, b2, -a1 * b2};
    int t = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
      for (int j = 0; j <= i; j++, t++) C[t] += r1[i] * r1[j] + r2[i] * r2[j];
  }
  blk_sum_vec<45>(C, S);
  // smallest eigenvector by inverse iteration on the Cholesky factor (every thread redundantly: deterministic).
  // Every loop has constant bounds and is fully unrolled so that C, L and x live in registers (with run-time indices
  // they sit in local memory and the ~1100 dependent loads of one solve cost more than the T-long passes around it).
  double L[45];
  double maxd = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) { if (C[i * (i + 1) / 2 + i] > maxd) maxd = C[i * (i + 1) / 2 + i]; }
  const double ridge = 1e-13 * maxd, tiny = 1e-30 * maxd + 1e-300;
#pragma unroll
  for (int i = 0; i < 9; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double s = C[i * (i + 1) / 2 + j] + (i == j ? ridge : 0.0);
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      if (i == j) L[i * (i + 1) / 2 + i] = sqrt(s > tiny ? s : tiny);
      else L[i * (i + 1) / 2 + j] = s / L[j * (j + 1) / 2 + j];
    }
  }
  double x[9];
#pragma unroll
  for (int i = 0; i < 9; i++) x[i] = 1.0 + 0.1 * i;
#pragma unroll 1
  for (int it = 0; it < 10; it++) {
#pragma unroll
    for (int i = 0; i < 9; i++) {           // L y = x
      double s = x[i];
#pragma unroll
      for (int k = 0; k < i; k++) s -= L[i * (i + 1) / 2 + k] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
#pragma unroll
    for (int i = 8; i >= 0; i--) {          // L^T z = y
      double s = x[i];
#pragma unroll
      for (int k = i + 1; k < 9; k++) s -= L[k * (k + 1) / 2 + i] * x[k];
      x[i] = s / L[i * (i + 1) / 2 + i];
    }
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) nrm += x[i] * x[i];
    nrm = sqrt(nrm);
    if (!(nrm > 0) || !isfinite(nrm)) return;   // keep the previous H
#pragma unroll
    for (int i = 0; i < 9; i++) x[i] /= nrm;
  }
  // denormH (utools.c:70-90)
  double* F = x;
  double r = A2[0], xx = A2[1], yy = A2[2];
  F[6] += xx * F[0] + yy * F[3];
  F[7] += xx * F[1] + yy * F[4];
  F[8] += xx * F[2] + yy * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r; F[3] *= r; F[4] *= r; F[5] *= r;
  r = 1 / A1[0]; xx = -A1[1] * r; yy = -A1[2] * r;
  for (int i = 0; i < 9; i += 3) {
    F[i] = r * F[i] + xx * F[i + 2];
    F[i + 1] = r * F[i + 1] + yy * F[i + 2];
  }
  for (int i = 0; i < 9; i++) H[i] = F[i];
}

// exp_ranH.c:905-947: at least MIN_GOOD_SYM_PTS+1 correspondences within CHECK_COEF*th symmetric transfer error (HDsSym)
__device__ bool blk_sym_ok(const double* __restrict__ u, int T, const double* H, double th, BlkScratch& S) {
  double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]}, H1[9];
  const bool inv_ok = inv3(Hm, H1);
  int c = 0;
  if (inv_ok)
    for (int j = threadIdx.x; j < T; j += RS_NT) if (sym_err(Hm, H1, u + 6 * j) <= CHECK_COEF * th) c++;
  c = blk_sum_i(c, S);
  return inv_ok && c > MIN_GOOD_SYM_PTS;
}

// exp_iterHcustom (exp_ranH.c:617-737) for one inner sample.  d0 = errors of the start model h (already in memory).
// Returns the best (I,J) seen and leaves the matching model in Hbest.
__device__ void blk_lo_iterate(const double* __restrict__ u, int T, double th, int etype, double* h, const double* d0, double* d, int* idx,
                               BlkScratch& S, int* bestI, double* bestJ, double* Hbest) {
  int mI = 0; double mJv[1] = {0};
  for (int j = threadIdx.x; j < T; j += RS_NT) { if (d0[j] <= th) mI++; mJv[0] += truncQuad(d0[j], th); }
  mI = blk_sum_i(mI, S);
  blk_sum_vec<1>(mJv, S);
  double mJ = mJv[0];
  *bestI = 0; *bestJ = 0;
  if (mI < 4) return;
  for (int i = 0; i < 9; i++) Hbest[i] = h[i];
  int n = blk_compact(d0, T, th * MWM, idx, S);
  blk_lsq(u, idx, n, h, S);
  double ths = TC * th;
  const double dth = (ths - th) / ILSQ_ITERS;
  HErr E;
  for (int it = 0; it < ILSQ_ITERS; it++) {
    int sI; double sJ;
    herr_setup(E, h, etype);
    blk_score(u, T, E, th, d, S, &sI, &sJ);
    n = blk_compact(d, T, ths * MWM, idx, S);
    if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
    if (n < 4) { *bestI = mI; *bestJ = mJ; return; }
    blk_lsq(u, idx, n, h, S);
    ths -= dth;
  }
  int sI; double sJ;
  herr_setup(E, h, etype);
  blk_score(u, T, E, th, nullptr, S, &sI, &sJ);
  if (mJ < sJ) { mJ = sJ; mI = sI; for (int i = 0; i < 9; i++) Hbest[i] = h[i]; }
  *bestI = mI; *bestJ = mJ;
}

// ---- per-batch update: best sample, symmetric check, local optimisation, stopping rule -------------------
// scratch: per inner sample w: dbuf[w][2T] doubles, ibuf[w][T] ints; then dS[T] / inl0[T] of the start model
// force_lo: the closing pass of exp_ranH.c:1085-1197 ("if there were no LOs, do at least one NOW")
__global__ void __launch_bounds__(RS_NT)
k_rsb_select(const double* __restrict__ u, int T, double th, int etype, int do_sym,
             const HypOut* __restrict__ hyp, int nhyp, int force_lo, RsState* st, LoShare* sh, double* dscr, int* iscr) {
  __shared__ BlkScratch S;
  if (force_lo ? (!st->done || st->lo_runs != 0) : st->done) {
    if (threadIdx.x == 0 && force_lo) sh->run_lo = 0;
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* dW = dscr;                                // slot 0 doubles as select's work buffer (LO runs afterwards)
  int* iW = iscr;
  double* dS = dscr + (size_t)RS_NW * 2 * T;        // errors of the LO start model
  int* inl0 = iscr + (size_t)RS_NW * T;             // its inliers at th

  // (a) best hypothesis of the batch: max J, lowest index on ties
  double bj = -1; int bi = -1, rej = 0;
  for (int k = threadIdx.x; k < nhyp; k += RS_NT) {
    const int f = hyp[k].flag;
    if (f == 1) rej++;
    if (f == 0 && (hyp[k].J > bj)) { bj = hyp[k].J; bi = k; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double oj = __shfl_xor_sync(0xffffffffu, bj, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (oj > bj || (oj == bj && (bi < 0 || oi < bi)))) { bj = oj; bi = oi; }
  }
  rej = blk_sum_i(rej, S);
  if (lane == 0) { S.bj[warp] = bj; S.bi[warp] = bi; }
  __syncthreads();
  bj = -1; bi = -1;
  for (int k = 0; k < RS_NWARP; k++) if (S.bi[k] >= 0 && (S.bj[k] > bj || (S.bj[k] == bj && S.bi[k] < bi))) { bj = S.bj[k]; bi = S.bi[k]; }

  // (b) sequential bookkeeping of exp_ranH.c:903-961 for the batch's best sample: every thread snapshots the state,
  // thread 0 writes it back after a barrier
  const double curJ = st->J;
  double curJs = st->Js;
  int have = st->have_sample, curIs = st->Is;
  const int no_sam = st->no_sam, lo_runs = st->lo_runs;
  __syncthreads();
  bool run_lo = false;
  if (bi >= 0) {
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = hyp[bi].H[i];
    const int I = hyp[bi].I;
    if (curJ < bj) {
      const bool ok = !do_sym || blk_sym_ok(u, T, h, th, S);
      if (ok && threadIdx.x == 0) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bj; st->I = I; }
    }
    if (!have || curJs < bj) {
      if (threadIdx.x == 0) { for (int i = 0; i < 9; i++) st->Hs[i] = h[i]; st->Js = bj; st->Is = I; st->have_sample = 1; }
      have = 1; curJs = bj; curIs = I;
      run_lo = no_sam + nhyp > ITER_SAM;
    }
  }
  if (threadIdx.x == 0 && rej) st->oc_rejects += rej;
  if (no_sam + nhyp >= ITER_SAM && lo_runs == 0 && have && curIs > 4) run_lo = true;
  if (force_lo) run_lo = have && lo_runs == 0;
  __syncthreads();
  if (run_lo) {
    // LSQ on the TC*th*MWM band of the best SAMPLE, then its inliers at th (exp_ranH.c:997-1012)
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = st->Hs[i];
    HErr E;
    int I; double J;
    herr_setup(E, h, etype);
    blk_score(u, T, E, th, dW, S, &I, &J);
    int n = blk_compact(dW, T, TC * th * MWM, iW, S);
    blk_lsq(u, iW, n, h, S);
    herr_setup(E, h, etype);
    blk_score(u, T, E, th, dS, S, &I, &J);
    n = blk_compact(dS, T, th, inl0, S);
    if (threadIdx.x == 0) { for (int i = 0; i < 9; i++) sh->h0[i] = h[i]; sh->n0 = n; st->lo_runs = lo_runs + 1; sh->lo_id = lo_runs + 1; }
  }
  if (threadIdx.x == 0) sh->run_lo = run_lo ? 1 : 0;
}

// (c) inner RANSAC (exp_inHranicustom, exp_ranH.c:741-793): one CTA per inner sample
__global__ void __launch_bounds__(RS_NT)
k_rsb_lo(const double* __restrict__ u, int T, double th, int etype, unsigned long long seed, int closing,
         const RsState* __restrict__ st, LoShare* sh, double* dscr, int* iscr) {
  __shared__ BlkScratch S;
  if (closing ? !st->done : st->done) return;      // a batch enqueued after the stopping rule fired does nothing
  if (!sh->run_lo) return;
  const int rep = blockIdx.x;
  double* dW = dscr + (size_t)rep * 2 * T;
  int* iW = iscr + (size_t)rep * T;
  const int* inl0 = iscr + (size_t)RS_NW * T;
  const int n0 = sh->n0;
  int bI = 0; double bJ = 0; double Hb[9], h0[9];
  for (int i = 0; i < 9; i++) { h0[i] = sh->h0[i]; Hb[i] = h0[i]; }
  if (n0 >= 8) {
    int ssiz = n0 / 2; if (ssiz > 12) ssiz = 12;
    // randsubset (rtools.c:25-39) on a private copy of the inlier list
    for (int k = threadIdx.x; k < n0; k += RS_NT) iW[k] = inl0[k];
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long stream = 0x4C4F000000000000ull + (unsigned long long)sh->lo_id * 64 + rep;
      for (int i = 0; i < ssiz; i++) {
        const int s = (int)rs_rand(seed, stream, i, (unsigned)(n0 - i)), j = n0 - i - 1;
        const int q = iW[s]; iW[s] = iW[j]; iW[j] = q;
      }
    }
    __syncthreads();
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = h0[i];
    blk_lsq(u, iW + n0 - ssiz, ssiz, h, S);
    HErr E;
    herr_setup(E, h, etype);
    int I; double J;
    blk_score(u, T, E, th, dW, S, &I, &J);
    blk_lo_iterate(u, T, th, etype, h, dW, dW + T, iW, S, &bI, &bJ, Hb);
  }
  if (threadIdx.x == 0) { sh->loI[rep] = bI; sh->loJ[rep] = bJ; for (int i = 0; i < 9; i++) sh->loH[rep][i] = Hb[i]; }
}

// (d) take the best inner sample (first on ties), accept against maxS, update the stopping rule
__global__ void __launch_bounds__(RS_NT)
k_rsb_accept(const double* __restrict__ u, int T, double th, double conf, int do_sym, int nhyp, int closing, RsState* st,
             const LoShare* sh) {
  __shared__ BlkScratch S;
  if (closing ? !st->done : st->done) return;
  const int run_lo = sh->run_lo;
  if (closing && !run_lo) return;
  if (run_lo) {
    int best = -1; double bJ = 0; int bI = 0;
    for (int k = 0; k < LO_REPS; k++) if (bJ < sh->loJ[k]) { bJ = sh->loJ[k]; bI = sh->loI[k]; best = k; }
    const double curJ = st->J;
    __syncthreads();
    if (best >= 0 && curJ < bJ) {
      double h[9];
      for (int i = 0; i < 9; i++) h[i] = sh->loH[best][i];
      if (det_ok(h) && (!do_sym || blk_sym_ok(u, T, h, th, S))) {
        if (threadIdx.x == 0) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bJ; st->I = bI; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && !closing) {
    st->no_sam += nhyp;
    if (st->I > 0) { const int ns = nsamples(st->I + 1, T, 4, conf); if (ns < st->max_sam) st->max_sam = ns; }
    st->done = st->no_sam >= st->max_sam;
  }
}

// final errors of the accepted model -> inlier mask + residuals (exp_ranH.c:1207-1212, *resids)
__global__ void k_rs_final(const double* __restrict__ u, int T, double th, int etype, const RsState* st, unsigned char* inl,
                           double* __restrict__ resid) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T) return;
  double e = 0;
  if (st->I > 0) {
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = st->H[i];
    HErr E;
    herr_setup(E, h, etype);
    e = herr(E, u + 6 * j);
  }
  resid[j] = e;
  inl[j] = (st->I > 0 && e <= th) ? 1 : 0;
}

}  // namespace

// Device part: correspondences d_u (T x 6 doubles) already on the device.  Batches are enqueued in groups (the kernels
// gate themselves on st->done); one host synchronisation per group.  resid (may be NULL): T doubles, the error of every
// correspondence under the returned model.
int mg_ransac_run(modsgpu_ctx* ctx, const double* d_u, int T, const modsgpu_ransac_params* p,
                  double* H, unsigned char* inl, modsgpu_ransac_result* res, double* resid) {
  const int NW = RS_NW;
  const int etype = p->error_type;
  if (etype < MODSGPU_ERR_SAMPSON || etype > MODSGPU_ERR_SYMM_SUM) MG_FAIL(ctx, MODSGPU_EINVAL, "unknown error_type");
  size_t off_sh = 256, off_hyp = off_sh + ((sizeof(LoShare) + 255) & ~(size_t)255), off_d = off_hyp + sizeof(HypOut) * RS_MAX_B;
  size_t off_i = off_d + sizeof(double) * (size_t)(2 * NW + 1) * T;
  size_t off_inl = off_i + sizeof(int) * (size_t)(NW + 1) * T;
  size_t off_res = (off_inl + T + 63) & ~(size_t)63;
  size_t total = off_res + sizeof(double) * (size_t)T + 64;
  MG_CUDA(ctx, ctx->rs_buf.ensure(total));
  uint8_t* base = ctx->rs_buf.as<uint8_t>();
  RsState* st = reinterpret_cast<RsState*>(base);
  LoShare* sh = reinterpret_cast<LoShare*>(base + off_sh);
  HypOut* hyp = reinterpret_cast<HypOut*>(base + off_hyp);
  double* dscr = reinterpret_cast<double*>(base + off_d);
  int* iscr = reinterpret_cast<int*>(base + off_i);
  unsigned char* dinl = base + off_inl;
  double* dres = reinterpret_cast<double*>(base + off_res);
  const size_t hs_bytes = 2 * sizeof(RsState) + 64 + (size_t)T * 9 + 64;
  MG_CUDA(ctx, ctx->h_stage.ensure(hs_bytes));
  RsState* hinit = ctx->h_stage.as<RsState>();        // [0] initial state (H2D source), [1] read-back
  RsState* hs = hinit + 1;
  memset(hinit, 0, sizeof(RsState));
  hinit->max_sam = p->max_samples;
  MG_CUDA(ctx, cudaMemcpyAsync(st, hinit, sizeof(RsState), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, cudaMemsetAsync(sh, 0, sizeof(LoShare), ctx->stream));
  unsigned char* hinl = reinterpret_cast<unsigned char*>(hs + 1) + 64;
  double* hres = reinterpret_cast<double*>(hinl + (((size_t)T + 63) & ~(size_t)63));
  int basei = 0, batch = 0;
  for (int group = 0;; group++) {
    const int nb = group == 0 ? 2 : 4;
    for (int b = 0; b < nb; b++, batch++) {
      const int B = batch == 0 ? 512 : (batch == 1 ? 1024 : RS_MAX_B);
      MG_PROF(ctx, "k_rs_hyp", 2, (double)B);
      k_rs_hyp<<<ceil_div(B, 8), 256, 0, ctx->stream>>>(d_u, T, p->th, etype, p->seed, basei, B, st, hyp);
      MG_LAUNCHED(ctx);
      MG_PROF(ctx, "k_rs_select", 2, (double)T);
      k_rsb_select<<<1, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, etype, p->do_sym_check, hyp, B, 0, st, sh, dscr, iscr);
      MG_LAUNCHED(ctx);
      MG_PROF(ctx, "k_rs_lo", 2, (double)T);
      k_rsb_lo<<<LO_REPS, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, etype, p->seed, 0, st, sh, dscr, iscr);
      MG_LAUNCHED(ctx);
      MG_PROF(ctx, "k_rs_accept", 2, (double)T);
      k_rsb_accept<<<1, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, p->conf, p->do_sym_check, B, 0, st, sh);
      MG_LAUNCHED(ctx);
      basei += B;
    }
    // closing kernels, speculative: they act only once the stopping rule has fired (st->done)
    // exp_ranH.c:1085-1197: "If there were no LOs, do at least one NOW"
    k_rsb_select<<<1, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, etype, p->do_sym_check, hyp, 0, 1, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rsb_lo<<<LO_REPS, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, etype, p->seed, 1, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rsb_accept<<<1, RS_NT, 0, ctx->stream>>>(d_u, T, p->th, p->conf, p->do_sym_check, 0, 1, st, sh);
    MG_LAUNCHED(ctx);
    k_rs_final<<<ceil_div(T, 256), 256, 0, ctx->stream>>>(d_u, T, p->th, etype, st, dinl, dres);
    MG_LAUNCHED(ctx);
    MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RsState), cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, cudaMemcpyAsync(hinl, dinl, T, cudaMemcpyDeviceToHost, ctx->stream));
    if (resid) MG_CUDA(ctx, cudaMemcpyAsync(hres, dres, sizeof(double) * (size_t)T, cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    if (hs->done) break;
  }
  for (int i = 0; i < 9; i++) H[i] = hs->H[i];
  memcpy(inl, hinl, T);
  if (resid) memcpy(resid, hres, sizeof(double) * (size_t)T);
  if (res) { res->n_inliers = hs->I; res->J = hs->J; res->samples = hs->no_sam; res->lo_runs = hs->lo_runs; res->oc_rejects = hs->oc_rejects; res->degen_runs = 0; res->h_inliers = 0; }
  return 0;
}

static int ransac_H_impl(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                         double* H, unsigned char* inl, modsgpu_ransac_result* res, double* resid) {
  if (!ctx || !p || !H || T < 0 || (T > 0 && (!u || !inl))) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (res) memset(res, 0, sizeof(*res));
  for (int i = 0; i < 9; i++) H[i] = 0;
  if (T < 4) {   // fewer than a minimal sample: no model (LORANSACFiltering requires >= MinimumSamples)
    for (int i = 0; i < T; i++) { inl[i] = 0; if (resid) resid[i] = 0; }
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  }
  MG_CUDA(ctx, ctx->io_a.ensure((size_t)T * 48));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, u, (size_t)T * 48, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_ransac_run(ctx, ctx->io_a.as<double>(), T, p, H, inl, res, resid);
  if (rc) return rc;
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}

extern "C" int modsgpu_ransac_H(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                                double* H, unsigned char* inl, modsgpu_ransac_result* res) {
  return ransac_H_impl(ctx, u, T, p, H, inl, res, nullptr);
}
extern "C" int modsgpu_ransac_H_resid(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                                      double* H, unsigned char* inl, modsgpu_ransac_result* res, double* resid) {
  return ransac_H_impl(ctx, u, T, p, H, inl, res, resid);
}
