// ransac.cu -- batched LO-RANSAC for a homography on the device (seam S4).
//
// Replaces exp_ransacHcustom (degensac/exp_ranH.c:796-1236) as LORANSACFiltering calls it
// (matching.cpp:731: iter_type 4, oriented constraint on, Sampson error, MSAC score, symmetric check).
// The reference is a sequential loop driven by libc rand(); this is the same estimator re-organised
// for a GPU:
//   k_rs_hyp     one WARP per hypothesis, B hypotheses per launch: counter-based RNG (no state carried
//                between samples) -> 4-point sample -> oriented constraint (Htools.c all_Hori_valid) ->
//                8x9 null space by pivoted Gauss-Jordan (utools.c:97-167) -> |det|/h33^3 test ->
//                Sampson error of all T correspondences (Htools.c:160-198), lanes striding over T ->
//                MSAC score (rtools.c truncQuad, 9/4*th width) by a fixed-order butterfly reduction
//   k_rs_select / k_rs_lo / k_rs_accept  per batch: best hypothesis (max J, lowest index on ties), symmetric
//                transfer check (exp_ranH.c:905-947), local optimisation = LSQ on the 8*th band + 10
//                inner samples x 4 shrinking-threshold LSQ steps (exp_inHranicustom / exp_iterHcustom),
//                the 10 inner samples on 10 SMs at once; adaptive stopping nsamples(I+1,T,4,conf)
// All arithmetic is fp64 and compiled with --fmad=false; every reduction has a fixed order, so a run is
// reproducible from (u, params.seed).
// Deviations from the reference, on purpose: (1) samples are consumed in batches, so at least B are drawn;
// (2) the LSQ null vector comes from 10 steps of inverse iteration on the 9x9 normal matrix instead of
// LAPACK dsyev_ (lapwrap.c:75-97; third-party, unpinned); (3) the inlier-set hash that prunes repeated
// LO iterations (exp_ranH.c __HASHING__) is dropped -- it only skips work whose result is already known.
#include "common.cuh"
#include "ransac_common.cuh"
#include "ransac_h.cuh"
#include <cmath>
#include <algorithm>

namespace {


struct RsState {
  double H[9];            // best model (maxS)
  double J; int I;
  double Hs[9];           // best sample so far (maxSs)
  double Js; int Is;
  int max_sam, no_sam, lo_runs, oc_rejects, done, have_sample;
};

struct HypOut { double H[9]; double J; int I; int flag; };   // flag: 0 ok, 1 oriented-constraint reject, 2 degenerate

// ---- hypothesis generation + scoring: one warp per hypothesis -------------------------------------------
__global__ void __launch_bounds__(256)
k_rs_hyp(const double* __restrict__ u, int T, double th, unsigned long long seed, int base, int nhyp, HypOut* __restrict__ out) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nhyp) return;
  const unsigned long long hid = (unsigned long long)(base + w);
  // 4 distinct indices: partial Fisher-Yates over a virtual pool (rtools.c sample()), counter-based draws
  int idx[4], pos[4], val[4];
  for (int i = 0; i < 4; i++) {
    const int s = (int)rs_rand(seed, hid, i, (unsigned)(T - i)), last = T - i - 1;
    int vs = s, vl = last;
    for (int k = 0; k < i; k++) { if (pos[k] == s) vs = val[k]; if (pos[k] == last) vl = val[k]; }
    idx[i] = vs;
    pos[i] = s; val[i] = vl;
  }
  double h[9];
  int flag = 0;
  if (!all_Hori_valid(u, idx)) flag = 1;
  else if (!h_from_4(u, idx, h) || !det_ok(h)) flag = 2;
  int I = 0;
  double J = 0;
  if (flag == 0) score_all(u, T, h, th, nullptr, lane, &I, &J);
  if (lane == 0) {
    HypOut o;
    for (int i = 0; i < 9; i++) o.H[i] = flag == 0 ? h[i] : 0.0;
    o.I = I; o.J = J; o.flag = flag;
    out[w] = o;
  }
}

// ---- per-batch update: best sample, symmetric check, local optimisation, stopping rule -------------------
// Three launches so that the fp64-heavy inner RANSAC spreads over LO_REPS SMs instead of sharing one:
//   k_rs_select  one CTA: best hypothesis of the batch, symmetric check, bookkeeping; decides whether the local
//                optimisation runs and prepares its start model (LSQ on the TC*th*MWM band) and inlier list
//   k_rs_lo      LO_REPS CTAs of one warp: one inner sample each (exp_inHranicustom / exp_iterHcustom)
//   k_rs_accept  one warp: best inner sample vs the best model, stopping rule
// scratch: per inner sample w: dbuf[w][2T] doubles, ibuf[w][T] ints; then dS[T] / inl0[T] of the start model

__global__ void __launch_bounds__(384)
k_rs_select(const double* __restrict__ u, int T, double th, int do_sym,
            const HypOut* __restrict__ hyp, int nhyp, int force_lo, RsState* st, LoShare* sh, double* dscr, int* iscr) {
  __shared__ double sJ[12]; __shared__ int sIdx[12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* dW = dscr;                                // slot 0 doubles as select's work buffer (LO runs afterwards)
  int* iW = iscr;
  double* dS = dscr + (size_t)RS_NW * 2 * T;        // errors of the LO start model
  int* inl0 = iscr + (size_t)RS_NW * T;             // its inliers at th

  // (a) best hypothesis of the batch: max J, lowest index on ties
  double bj = -1; int bi = -1, rej = 0;
  for (int k = threadIdx.x; k < nhyp; k += blockDim.x) {
    const int f = hyp[k].flag;
    if (f == 1) rej++;
    if (f == 0 && (hyp[k].J > bj)) { bj = hyp[k].J; bi = k; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double oj = __shfl_xor_sync(0xffffffffu, bj, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (oj > bj || (oj == bj && (bi < 0 || oi < bi)))) { bj = oj; bi = oi; }
  }
  rej = warp_sum_i(rej);
  if (lane == 0) { sJ[warp] = bj; sIdx[warp] = bi; if (rej) atomicAdd(&st->oc_rejects, rej); }
  __syncthreads();
  if (warp != 0) return;
  bj = -1; bi = -1;
  for (int k = 0; k < nw; k++) if (sIdx[k] >= 0 && (sJ[k] > bj || (sJ[k] == bj && sIdx[k] < bi))) { bj = sJ[k]; bi = sIdx[k]; }

  // (b) sequential bookkeeping of exp_ranH.c:903-961 for the batch's best sample
  // snapshot the state before any lane writes it (lanes of a warp need not run in lockstep)
  const double curJ = st->J;
  double curJs = st->Js;
  int have = st->have_sample, curIs = st->Is;
  const int no_sam = st->no_sam, lo_runs = st->lo_runs;
  __syncwarp();
  bool run_lo = false;
  if (bi >= 0) {
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = hyp[bi].H[i];
    const int I = hyp[bi].I;
    if (curJ < bj) {
      const bool ok = !do_sym || sym_check_ok(u, T, h, th, lane);
      if (ok && lane == 0) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bj; st->I = I; }
    }
    if (!have || curJs < bj) {
      if (lane == 0) { for (int i = 0; i < 9; i++) st->Hs[i] = h[i]; st->Js = bj; st->Is = I; st->have_sample = 1; }
      have = 1; curJs = bj; curIs = I;
      run_lo = no_sam + nhyp > ITER_SAM;
    }
  }
  if (no_sam + nhyp >= ITER_SAM && lo_runs == 0 && have && curIs > 4) run_lo = true;
  if (force_lo) run_lo = have && lo_runs == 0;
  __syncwarp();
  if (run_lo) {
    // LSQ on the TC*th*MWM band of the best SAMPLE, then its inliers at th (exp_ranH.c:997-1012)
    double h[9];
    for (int i = 0; i < 9; i++) h[i] = st->Hs[i];
    int I; double J;
    score_all(u, T, h, th, dW, lane, &I, &J);
    __syncwarp();
    int n = compact_inliers(dW, T, TC * th * MWM, iW, lane);
    lsq_h(u, iW, n, h, lane);
    score_all(u, T, h, th, dS, lane, &I, &J);
    __syncwarp();
    n = compact_inliers(dS, T, th, inl0, lane);
    if (lane == 0) { for (int i = 0; i < 9; i++) sh->h0[i] = h[i]; sh->n0 = n; st->lo_runs = lo_runs + 1; sh->lo_id = lo_runs + 1; }
  }
  if (lane == 0) sh->run_lo = run_lo ? 1 : 0;
}

// (d) take the best inner sample (first on ties), accept against maxS, update the stopping rule
__global__ void __launch_bounds__(32)
k_rs_accept(const double* __restrict__ u, int T, double th, double conf, int do_sym, int nhyp, RsState* st, const LoShare* sh) {
  const int lane = threadIdx.x;
  if (sh->run_lo) {
    int best = -1; double bJ = 0; int bI = 0;
    for (int k = 0; k < LO_REPS; k++) if (bJ < sh->loJ[k]) { bJ = sh->loJ[k]; bI = sh->loI[k]; best = k; }
    const double curJ = st->J;
    __syncwarp();
    if (best >= 0 && curJ < bJ) {
      double h[9];
      for (int i = 0; i < 9; i++) h[i] = sh->loH[best][i];
      if (det_ok(h) && (!do_sym || sym_check_ok(u, T, h, th, lane))) {
        if (lane == 0) { for (int i = 0; i < 9; i++) st->H[i] = h[i]; st->J = bJ; st->I = bI; }
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    st->no_sam += nhyp;
    if (st->I > 0) { const int ns = nsamples(st->I + 1, T, 4, conf); if (ns < st->max_sam) st->max_sam = ns; }
    st->done = st->no_sam >= st->max_sam;
  }
}

// final errors of the accepted model -> inlier mask (exp_ranH.c:1207-1212)
__global__ void k_rs_final(const double* __restrict__ u, int T, double th, const RsState* st, unsigned char* inl) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= T) return;
  inl[j] = (st->I > 0 && sampson(st->H, u + 6 * j) <= th) ? 1 : 0;
}

}  // namespace

// Device part: correspondences d_u (T x 6 doubles) already on the device.  On return (after a stream sync
// inside -- the stopping rule is data dependent) ctx->rs_buf holds the RsState followed by the inlier mask.
int mg_ransac_run(modsgpu_ctx* ctx, const double* d_u, int T, const modsgpu_ransac_params* p,
                  double* H, unsigned char* inl, modsgpu_ransac_result* res) {
  const int NW = RS_NW;
  size_t off_sh = 256, off_hyp = off_sh + ((sizeof(LoShare) + 255) & ~(size_t)255), off_d = off_hyp + sizeof(HypOut) * RS_MAX_B;
  size_t off_i = off_d + sizeof(double) * (size_t)(2 * NW + 1) * T;
  size_t off_inl = off_i + sizeof(int) * (size_t)(NW + 1) * T;
  size_t total = off_inl + T + 64;
  MG_CUDA(ctx, ctx->rs_buf.ensure(total));
  uint8_t* base = ctx->rs_buf.as<uint8_t>();
  RsState* st = reinterpret_cast<RsState*>(base);
  LoShare* sh = reinterpret_cast<LoShare*>(base + off_sh);
  HypOut* hyp = reinterpret_cast<HypOut*>(base + off_hyp);
  double* dscr = reinterpret_cast<double*>(base + off_d);
  int* iscr = reinterpret_cast<int*>(base + off_i);
  unsigned char* dinl = base + off_inl;
  MG_CUDA(ctx, ctx->h_stage.ensure(sizeof(RsState) + T + 64));
  RsState* hs = ctx->h_stage.as<RsState>();
  memset(hs, 0, sizeof(RsState));
  hs->max_sam = p->max_samples;
  MG_CUDA(ctx, cudaMemcpyAsync(st, hs, sizeof(RsState), cudaMemcpyHostToDevice, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));   // hs is reused as the read-back buffer below
  int basei = 0, batch = 0;
  for (;;) {
    int B = batch == 0 ? 512 : (batch == 1 ? 1024 : RS_MAX_B);
    MG_PROF(ctx, "k_rs_hyp", 2, (double)B);
    k_rs_hyp<<<ceil_div(B, 8), 256, 0, ctx->stream>>>(d_u, T, p->th, p->seed, basei, B, hyp);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_rs_select", 2, (double)T);
    k_rs_select<<<1, 384, 0, ctx->stream>>>(d_u, T, p->th, p->do_sym_check, hyp, B, 0, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_rs_lo", 2, (double)T);
    k_rs_lo<<<LO_REPS, 32, 0, ctx->stream>>>(d_u, T, p->th, p->seed, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_rs_accept", 2, (double)T);
    k_rs_accept<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, p->conf, p->do_sym_check, B, st, sh);
    MG_LAUNCHED(ctx);
    MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RsState), cudaMemcpyDeviceToHost, ctx->stream));
    MG_CUDA(ctx, mg_stream_sync(ctx));
    basei += B; batch++;
    if (hs->done) break;
  }
  if (hs->lo_runs == 0) {   // exp_ranH.c:1085-1197: "If there were no LOs, do at least one NOW"
    k_rs_select<<<1, 384, 0, ctx->stream>>>(d_u, T, p->th, p->do_sym_check, hyp, 0, 1, st, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rs_lo<<<LO_REPS, 32, 0, ctx->stream>>>(d_u, T, p->th, p->seed, sh, dscr, iscr);
    MG_LAUNCHED(ctx);
    k_rs_accept<<<1, 32, 0, ctx->stream>>>(d_u, T, p->th, p->conf, p->do_sym_check, 0, st, sh);
    MG_LAUNCHED(ctx);
  }
  k_rs_final<<<ceil_div(T, 256), 256, 0, ctx->stream>>>(d_u, T, p->th, st, dinl);
  MG_LAUNCHED(ctx);
  unsigned char* hinl = reinterpret_cast<unsigned char*>(hs + 1);
  MG_CUDA(ctx, cudaMemcpyAsync(hs, st, sizeof(RsState), cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, cudaMemcpyAsync(hinl, dinl, T, cudaMemcpyDeviceToHost, ctx->stream));
  MG_CUDA(ctx, mg_stream_sync(ctx));
  for (int i = 0; i < 9; i++) H[i] = hs->H[i];
  memcpy(inl, hinl, T);
  if (res) { res->n_inliers = hs->I; res->J = hs->J; res->samples = hs->no_sam; res->lo_runs = hs->lo_runs; res->oc_rejects = hs->oc_rejects; res->degen_runs = 0; res->h_inliers = 0; }
  return 0;
}

extern "C" int modsgpu_ransac_H(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                                double* H, unsigned char* inl, modsgpu_ransac_result* res) {
  if (!ctx || !p || !H || T < 0 || (T > 0 && (!u || !inl))) return MODSGPU_EINVAL;
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  if (res) memset(res, 0, sizeof(*res));
  for (int i = 0; i < 9; i++) H[i] = 0;
  if (T < 4) {   // fewer than a minimal sample: no model (LORANSACFiltering requires >= MinimumSamples)
    for (int i = 0; i < T; i++) inl[i] = 0;
    return mg_end(ctx) ? MODSGPU_ECUDA : 0;
  }
  MG_CUDA(ctx, ctx->io_a.ensure((size_t)T * 48));
  MG_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, u, (size_t)T * 48, cudaMemcpyHostToDevice, ctx->stream));
  int rc = mg_ransac_run(ctx, ctx->io_a.as<double>(), T, p, H, inl, res);
  if (rc) return rc;
  return mg_end(ctx) ? MODSGPU_ECUDA : 0;
}
