// synth.cu -- view synthesis on the device.
//
// Replaces GenerateSynthImageCorr (synth-detection.cpp:324-518, the non-AREA_INTERP branch):
//   rotate by phi (cv::warpAffine, INTER_LINEAR, constant border 128) -> anisotropic anti-aliasing blur
//   (cv::GaussianBlur, separate kernel sizes / sigmas, BORDER_REFLECT_101) -> tilt / zoom (cv::warpAffine).
// The three OpenCV calls are restated with the arithmetic of OpenCV 4.x for CV_32F (oracle/mods_oracle.cpp,
// pinned bit-exactly against cv2 4.13 by tests/golden/synth_pins.npz):
//   warpAffine : inverse matrix in double, source coordinates in 22.10 fixed point rounded to 1/32 px,
//                4 float weights, products accumulated left to right without fusion
//   GaussianBlur: row / column forms of sepFilter2D (fused in the SIMD body, scalar tails as compiled)
// Compiled with --fmad=false; every fused operation is an explicit fmaf().
// These kernels are HBM-bound streaming passes (8 B/px for a warp, 8 B/px per blur pass); they are written for
// exactness first -- one thread per output pixel, taps in shared memory.
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace {

struct WarpArgs { double M[6]; };   // inverted matrix (dst -> src)

__global__ void k_warp_affine(const float* __restrict__ in, int w, int h, WarpArgs a, float* __restrict__ out, int ow, int oh,
                              float border) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  const int AB_SCALE = 1024, round_delta = 16;
  const int adelta = __double2int_rn(a.M[0] * x * AB_SCALE), bdelta = __double2int_rn(a.M[3] * x * AB_SCALE);
  const int X0 = __double2int_rn((a.M[1] * y + a.M[2]) * AB_SCALE) + round_delta;
  const int Y0 = __double2int_rn((a.M[4] * y + a.M[5]) * AB_SCALE) + round_delta;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = X >> 5, sy = Y >> 5;
  const float fa = (float)(X & 31) / 32.0f, fb = (float)(Y & 31) / 32.0f;
  const float w0 = (1.0f - fb) * (1.0f - fa), w1 = (1.0f - fb) * fa, w2 = fb * (1.0f - fa), w3 = fb * fa;
  auto PX = [&](int yy, int xx) -> float {
    return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? in[(size_t)yy * w + xx] : border;
  };
  float v = PX(sy, sx) * w0 + PX(sy, sx + 1) * w1;
  v = v + PX(sy + 1, sx) * w2;
  v = v + PX(sy + 1, sx + 1) * w3;
  out[(size_t)y * ow + x] = v;
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

constexpr int SY_MAX_KS = 127;
struct TapsXY { float k[SY_MAX_KS + 1]; int ks; };

__global__ void k_blur_row_xy(const float* __restrict__ in, float* __restrict__ out, int w, int h, TapsXY tp) {
  __shared__ float k[SY_MAX_KS + 1];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < tp.ks) k[tid] = tp.k[tid];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int ks = tp.ks, r = ks >> 1;
  const float* row = in + (size_t)y * w;
  auto P = [&](int d) -> float { return row[reflect101(d, w)]; };
  float s;
  if (ks == 1) s = P(x) * k[0];
  else if (ks == 3) {
    const float p1 = P(x + 1) + P(x - 1), x0 = P(x);
    if (x < (w & ~1)) s = fmaf(x0, k[1], p1 * k[2]);
    else s = fmaf(p1, k[2], x0 * k[1]);
  } else if (ks == 5) {
    const float p1 = P(x + 1) + P(x - 1), p2 = P(x + 2) + P(x - 2), x0 = P(x);
    if (x < (w & ~1)) { s = p1 * k[3]; s = fmaf(x0, k[2], s); s = fmaf(p2, k[4], s); }
    else { s = x0 * k[2] + p1 * k[3]; s = s + p2 * k[4]; }
  } else if (x < (w & ~3)) {
    s = 0.f;
    for (int t = 0; t < ks; t++) s = fmaf(P(x + t - r), k[t], s);
  } else {
    const int nf = (ks - 1) % 4;
    s = P(x - r) * k[0];
    for (int t = 1; t < ks; t++) {
      if (t >= ks - nf) s = fmaf(P(x + t - r), k[t], s);
      else s = s + P(x + t - r) * k[t];
    }
  }
  out[(size_t)y * w + x] = s;
}

__global__ void k_blur_col_xy(const float* __restrict__ in, float* __restrict__ out, int w, int h, TapsXY tp) {
  __shared__ float k[SY_MAX_KS + 1];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < tp.ks) k[tid] = tp.k[tid];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int ks = tp.ks, r = ks >> 1;
  auto T = [&](int d) -> float { return in[(size_t)reflect101(d, h) * w + x]; };
  float s = T(y) * k[r];
  if (ks == 3 || x < (w & ~7))
    for (int t = 1; t <= r; t++) s = fmaf(T(y - t) + T(y + t), k[r + t], s);
  else
    for (int t = 1; t <= r; t++) s = s + (T(y - t) + T(y + t)) * k[r + t];
  out[(size_t)y * w + x] = s;
}

void invert_affine(const double* Min, double* M) {   // cv::warpAffine's own inversion (imgwarp.cpp)
  for (int i = 0; i < 6; i++) M[i] = Min[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11; M[1] *= -D;
  M[3] *= -D; M[4] = A22;
  double b1 = -M[0] * M[2] - M[1] * M[5];
  double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
}

int taps_xy(modsgpu_ctx* ctx, int ks, double sigma, TapsXY* t) {
  if (ks > SY_MAX_KS) MG_FAIL(ctx, MODSGPU_EINVAL, "anti-aliasing kernel too wide (ksize > 127)");
  const int r = ks / 2;
  std::vector<double> kd(ks);
  double sum = 0;
  for (int i = 0; i < ks; i++) { double x = i - r; kd[i] = std::exp(-x * x / (2.0 * sigma * sigma)); sum += kd[i]; }
  memset(t, 0, sizeof(*t));
  t->ks = ks;
  for (int i = 0; i < ks; i++) t->k[i] = (float)(kd[i] / sum);
  return 0;
}

int launch_warp(modsgpu_ctx* ctx, const float* in, int w, int h, const double* M, float* out, int ow, int oh) {
  WarpArgs a;
  invert_affine(M, a.M);
  dim3 blk(32, 8), grid(ceil_div(ow, 32), ceil_div(oh, 8));
  MG_PROF(ctx, "k_warp_affine", 0, (double)ow * oh * 8.0);
  k_warp_affine<<<grid, blk, 0, ctx->stream>>>(in, w, h, a, out, ow, oh, 128.f);
  MG_LAUNCHED(ctx);
  return 0;
}

}  // namespace

// geometry of GenerateSynthImageCorr (synth-detection.cpp:356-431); returns 1 for the identity view
extern "C" int modsgpu_synth_geometry(int w, int h, double tilt, double phi, double zoom, int* ow, int* oh, double* H) {
  bool vertical = false;
  if (tilt < 0) { tilt = -tilt; vertical = true; }
  const int zoomed = std::fabs(zoom - 1.0f) >= 0.05 ? 1 : 0;
  const int wS1 = (int)(w * zoom), hS1 = (int)(h * zoom);
  for (int i = 0; i < 9; i++) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if ((std::fabs(tilt - 1.) <= 0.1) && (std::abs((int)phi) <= 0.2) && (std::fabs(zoom - 1.) <= 0.1)) {   // int abs(phi), :366
    *ow = w; *oh = h;
    return 1;
  }
  double kV = 1., kH = 1.;
  if (zoomed) { kV = (double)w / (double)wS1; kH = (double)h / (double)hS1; }
  const double tx = vertical ? kH : tilt * kH, ty = vertical ? tilt * kV : kV;
  const double c = std::cos(phi), s = std::sin(phi);
  double w_new, h_new;
  if ((phi >= 0) && (phi < M_PI / 2)) {
    w_new = std::floor((0.5 + c * w + s * h) / tx);
    h_new = std::floor((0.5 + s * w + c * h) / ty);
    H[0] = c / tx; H[1] = s / tx; H[2] = 0;
    H[3] = -s / ty; H[4] = c / ty; H[5] = std::floor(0.5 + s * w / ty);
  } else {
    w_new = std::floor((0.5 - c * w + s * h) / tx);
    h_new = std::floor((0.5 + s * w - c * h) / ty);
    H[0] = c / tx; H[1] = s / tx; H[2] = -std::floor(c * w / tx);
    H[3] = -s / ty; H[4] = c / ty; H[5] = std::floor(0.5 + (s * w - c * h) / ty);
  }
  H[6] = 0; H[7] = 0; H[8] = 1;
  *ow = (int)w_new; *oh = (int)h_new;
  return 0;
}

extern "C" int modsgpu_synth_view(modsgpu_ctx* ctx, const modsgpu_image* in, double tilt, double phi, double zoom,
                                  double InitSigma, int doBlur, modsgpu_image** out, double* H) {
  if (!ctx || !in || !out || !H) return MODSGPU_EINVAL;
  *out = nullptr;
  const int w = in->w, h = in->h;
  int ow, oh;
  const int ident = modsgpu_synth_geometry(w, h, tilt, phi, zoom, &ow, &oh, H);
  if (ow <= 0 || oh <= 0) MG_FAIL(ctx, MODSGPU_EINVAL, "synthesised view has no pixels");
  if (mg_begin(ctx)) return MODSGPU_ECUDA;
  modsgpu_image* img = new modsgpu_image();
  img->w = ow; img->h = oh;
  MG_CUDA(ctx, mg_image_alloc(ctx, (size_t)ow * oh * sizeof(float), &img->d));
  if (ident) {   // "original image cloned" (synth-detection.cpp:366-377)
    MG_CUDA(ctx, cudaMemcpyAsync(img->d, in->d, (size_t)w * h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    if (mg_end(ctx)) return MODSGPU_ECUDA;
    *out = img;
    return 0;
  }
  bool vertical = false;
  if (tilt < 0) { tilt = -tilt; vertical = true; }
  const int zoomed = std::fabs(zoom - 1.0f) >= 0.05 ? 1 : 0;
  const int wS1 = (int)(w * zoom), hS1 = (int)(h * zoom);
  double kV = 1., kH = 1.;
  if (zoomed) { kV = (double)w / (double)wS1; kH = (double)h / (double)hS1; }
  const double sigma_aa_2 = zoomed ? InitSigma / (4.0 * zoom) : InitSigma / 2.0;
  const double sigma_aa = InitSigma * tilt / (2.0 * zoom);
  const double sigma_x = vertical ? sigma_aa_2 : sigma_aa, sigma_y = vertical ? sigma_aa : sigma_aa_2;
  int wr, hr;
  double R[6];
  if ((phi >= 0) && (phi < M_PI / 2)) {
    wr = (int)std::floor((0.5 + std::cos(phi) * w + std::sin(phi) * h));
    hr = (int)std::floor((0.5 + std::sin(phi) * w + std::cos(phi) * h));
    R[0] = std::cos(phi); R[1] = std::sin(phi); R[2] = 0;
    R[3] = -std::sin(phi); R[4] = std::cos(phi); R[5] = std::floor(0.5 + std::sin(phi) * w);
  } else {
    wr = (int)std::floor((0.5 - std::cos(phi) * w + std::sin(phi) * h));
    hr = (int)std::floor((0.5 + std::sin(phi) * w - std::cos(phi) * h));
    R[0] = std::cos(phi); R[1] = std::sin(phi); R[2] = -std::floor(std::cos(phi) * w);
    R[3] = -std::sin(phi); R[4] = std::cos(phi); R[5] = std::floor(0.5 + (std::sin(phi) * w - std::cos(phi) * h));
  }
  if (wr <= 0 || hr <= 0) MG_FAIL(ctx, MODSGPU_EINVAL, "phi must lie in [0, pi)");
  const size_t rb = (size_t)wr * hr * 4;
  MG_CUDA(ctx, ctx->io_a.ensure(rb));
  MG_CUDA(ctx, ctx->io_b.ensure(rb));
  float* rot = ctx->io_a.as<float>();
  float* tmp = ctx->io_b.as<float>();
  int rc = launch_warp(ctx, in->d, w, h, R, rot, wr, hr);
  if (rc) return rc;
  if (doBlur) {
    int kx = (int)std::floor(2.0 * 3.0 * sigma_x + 1.0);
    if (kx % 2 == 0) kx++;
    if (kx < 3) kx = 3;
    int ky = (int)std::floor(2.0 * 3.0 * sigma_y + 1.0);
    if (ky % 2 == 0) ky++;
    if (ky < 3) ky = 3;
    TapsXY tx, ty;
    if ((rc = taps_xy(ctx, kx, sigma_x, &tx)) || (rc = taps_xy(ctx, ky, sigma_y, &ty))) return rc;
    dim3 blk(32, 8), grid(ceil_div(wr, 32), ceil_div(hr, 8));
    MG_PROF(ctx, "k_blur_row_xy", 0, (double)wr * hr * 8.0);
    k_blur_row_xy<<<grid, blk, 0, ctx->stream>>>(rot, tmp, wr, hr, tx);
    MG_LAUNCHED(ctx);
    MG_PROF(ctx, "k_blur_col_xy", 0, (double)wr * hr * 8.0);
    k_blur_col_xy<<<grid, blk, 0, ctx->stream>>>(tmp, rot, wr, hr, ty);   // in-place like the reference: result back in `rot`
    MG_LAUNCHED(ctx);
  }
  double Wm[6] = {0, 0, 0, 0, 0, 0};
  if (vertical) { Wm[0] = 1.0 / kH; Wm[4] = 1.0 / (tilt * kV); }
  else { Wm[0] = 1.0 / (tilt * kH); Wm[4] = 1.0 / kV; }
  rc = launch_warp(ctx, rot, wr, hr, Wm, img->d, ow, oh);
  if (rc) return rc;
  if (mg_end(ctx)) return MODSGPU_ECUDA;
  *out = img;
  return 0;
}
