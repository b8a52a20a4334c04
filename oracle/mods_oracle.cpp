/*
 * mods_oracle.cpp -- CPU restatement of the MODS hot path.  TEST INFRASTRUCTURE ONLY
 * (see mods_oracle.h).  Build: oracle/Makefile (g++ -O2 -ffp-contract=off -mfma).
 *
 * All float arithmetic is written with explicit fmaf() where the reference's
 * third-party back end (OpenCV 4.13 SIMD kernels) fuses, and plain ops elsewhere;
 * the file MUST be compiled with -ffp-contract=off so the compiler adds no fusion.
 *
 * Citations are file:line in the upstream tree (ducha-aiki/mods-light-zmq @ 33c9ba2).
 */
#include "mods_oracle.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

/* ========================================================================== */
/* image primitives                                                           */
/* ========================================================================== */

/* synth-detection.cpp:344-351: convertTo(CV_32FC3); split; (p0+p1+p2)/3.0.
 * The MatExpr "/3.0" is evaluated by cv::Mat::convertTo(alpha = 1.0/3.0) whose
 * 32f->32f kernel multiplies by (float)alpha. */
extern "C" void orc_gray_from_bgr(const uint8_t* bgr, int w, int h, float* gray) {
  const float third = (float)(1.0 / 3.0);
  for (long i = 0; i < (long)w * h; i++) {
    float s = ((float)bgr[3 * i] + (float)bgr[3 * i + 1]) + (float)bgr[3 * i + 2];
    gray[i] = s * third;
  }
}

/* cv::getGaussianKernel(ksize, sigma, CV_32F) as called by cv::GaussianBlur from
 * helpers.cpp:717-731: ksize = (int)(2*3*sigma+1), forced odd. */
extern "C" int orc_gaussian_kernel(float sigmaf, float* taps) {
  double sigma = (double)sigmaf;
  int ks = (int)(2.0 * 3.0 * sigma + 1.0);
  if (ks % 2 == 0) ks++;
  int r = ks / 2;
  std::vector<double> kd(ks);
  double sum = 0;
  for (int i = 0; i < ks; i++) {
    double x = i - r;
    kd[i] = std::exp(-x * x / (2.0 * sigma * sigma));
    sum += kd[i];
  }
  for (int i = 0; i < ks; i++) taps[i] = (float)(kd[i] / sum);
  return ks;
}

/* helpers.cpp:717-731 gaussianBlur()/gaussianBlurInplace(): cv::GaussianBlur with
 * BORDER_REPLICATE.  Arithmetic order = OpenCV 4.13 sepFilter2D for CV_32F as shipped
 * in the cv2 wheel of this image (AVX2/AVX-512 dispatch), established empirically and
 * pinned bit-exactly by tests/test_oracle_cv2_pin.py:
 *   row pass, ksize>=7 : x <  w&~3 : s=0; s=fma(src[x-r+t],k[t],s) for t=0..ks-1
 *                        x >= w&~3 : s=src[x-r]*k[0]; s+=src*k[t] unfused, except the
 *                                    last (ks-1)%4 taps which are fused (scalar tail as
 *                                    compiled in the wheel)
 *   row pass, ksize==5 : x <  w&~1 : s=(x[+1]+x[-1])*k1; s=fma(x0,k0,s); s=fma(x[+2]+x[-2],k2,s)
 *                        else      : s=(x0*k0+(x[+1]+x[-1])*k1)+(x[+2]+x[-2])*k2  (unfused)
 *   col pass           : s=c*k0; for t=1..r: s (+)= (up_t+down_t)*k[t], fused iff x < w&~7
 */
extern "C" void orc_gaussian_blur(const float* in, float* out, int w, int h, float sigma) {
  std::vector<float> k(4096);
  int ks = orc_gaussian_kernel(sigma, k.data());
  int r = ks / 2;
  std::vector<float> tmp((size_t)w * h);
  const int wv = w & ~3, wc = w & ~7;
  auto PX = [&](int y, int xx) -> float {
    xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
    return in[(size_t)y * w + xx];
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float s;
      if (ks == 5) {
        float p1 = PX(y, x + 1) + PX(y, x - 1), p2 = PX(y, x + 2) + PX(y, x - 2), x0 = PX(y, x);
        if (x < (w & ~1)) {
          s = p1 * k[3];
          s = fmaf(x0, k[2], s);
          s = fmaf(p2, k[4], s);
        } else {
          s = x0 * k[2] + p1 * k[3];
          s = s + p2 * k[4];
        }
      } else if (ks < 5) { /* not reachable from the hot path (sigma >= 0.75); plain symmetric sum */
        s = PX(y, x) * k[r];
        for (int t = 1; t <= r; t++) s = s + (PX(y, x + t) + PX(y, x - t)) * k[r + t];
      } else if (x < wv) {
        s = 0;
        for (int t = 0; t < ks; t++) s = fmaf(PX(y, x + t - r), k[t], s);
      } else {
        int nf = (ks - 1) % 4;
        s = PX(y, x - r) * k[0];
        for (int t = 1; t < ks; t++) {
          if (t >= ks - nf) s = fmaf(PX(y, x + t - r), k[t], s);
          else s = s + PX(y, x + t - r) * k[t];
        }
      }
      tmp[(size_t)y * w + x] = s;
    }
  auto TY = [&](int yy, int x) -> float {
    yy = yy < 0 ? 0 : (yy >= h ? h - 1 : yy);
    return tmp[(size_t)yy * w + x];
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float s = tmp[(size_t)y * w + x] * k[r];
      if (x < wc)
        for (int t = 1; t <= r; t++) s = fmaf(TY(y - t, x) + TY(y + t, x), k[r + t], s);
      else
        for (int t = 1; t <= r; t++) s = s + (TY(y - t, x) + TY(y + t, x)) * k[r + t];
      out[(size_t)y * w + x] = s;
    }
}

/* pyramid.cpp:196-254 HessianResponse: 3x3 stencil, interior only.  The reference
 * leaves the 1-px frame uninitialised (Mat without zeroing, :203); the oracle defines it
 * as 0 -- it is never read with border >= 2. */
extern "C" void orc_hessian_response(const float* in, float* out, int w, int h, float norm) {
  const float norm2 = norm * norm;
  std::memset(out, 0, sizeof(float) * (size_t)w * h);
  for (int r = 1; r < h - 1; r++)
    for (int c = 1; c < w - 1; c++) {
      const float* p = in + (size_t)r * w + c;
      float v11 = p[-w - 1], v12 = p[-w], v13 = p[-w + 1];
      float v21 = p[-1], v22 = p[0], v23 = p[1];
      float v31 = p[w - 1], v32 = p[w], v33 = p[w + 1];
      float Lxx = (v21 - 2 * v22 + v23);
      float Lyy = (v12 - 2 * v22 + v32);
      float Lxy = (v13 - v11 + v31 - v33) / 4.0f;
      out[(size_t)r * w + c] = (Lxx * Lyy - Lxy * Lxy) * norm2;
    }
}

/* pyramid.cpp:476 cv::resize(nextBlur, next, Size(0,0), 0.5, 0.5, INTER_LINEAR).
 * dsize = cvRound(w*0.5) (round-half-even).  Sample position 2*d+0.5 -> weights (.5,.5).
 * cv2 4.13 of this image routes 32F INTER_LINEAR through its IPP HAL, whose arithmetic is
 * the lerp form a+(b-a)*0.5 horizontally then vertically (pinned bit-exactly by
 * tests/test_oracle_cv2_pin.py for even sizes; odd sizes clamp the +1 neighbour). */
extern "C" void orc_half_size(int w, int h, int* ow, int* oh) {
  *ow = (int)std::nearbyint(w * 0.5);
  *oh = (int)std::nearbyint(h * 0.5);
}
extern "C" void orc_half_image(const float* in, int w, int h, float* out) {
  int ow, oh;
  orc_half_size(w, h, &ow, &oh);
  for (int y = 0; y < oh; y++) {
    int y0 = std::min(2 * y, h - 1), y1 = std::min(2 * y + 1, h - 1);
    for (int x = 0; x < ow; x++) {
      int x0 = std::min(2 * x, w - 1), x1 = std::min(2 * x + 1, w - 1);
      float a = in[(size_t)y0 * w + x0], b = in[(size_t)y0 * w + x1];
      float c = in[(size_t)y1 * w + x0], d = in[(size_t)y1 * w + x1];
      float r0 = a + (b - a) * 0.5f;
      float r1 = c + (d - c) * 0.5f;
      out[(size_t)y * ow + x] = r0 + (r1 - r0) * 0.5f;
    }
  }
}

/* ========================================================================== */
/* classic stages: dominant orientation, (Root)SIFT                            */
/* ========================================================================== */

/* helpers.cpp:30-72: the literal table holds atan(i/255) to 10 decimals; three entries of the reference carry
 * typos (32, 83, 100) and are reproduced as such.  Values are parsed from their decimal form like the compiler
 * parses the reference's literals. */
extern "C" void orc_atan_lut(double* lut) {
  char buf[64];
  for (int i = 0; i < 256; i++) {
    snprintf(buf, sizeof(buf), "%.10f", std::atan(i / 255.0));
    lut[i] = strtod(buf, nullptr);
  }
  lut[32] = strtod("0.1248376255", nullptr);
  lut[83] = strtod("0.3146752558", nullptr);
  lut[100] = strtod("0.3737268255", nullptr);
}

/* helpers.cpp:160-207 atan2LUTff */
static float atan2LUTff(const double* LUT, float y, float x) {
  const float PI2f = 1.57079632679489661923f, PIf = 3.14159265358979323846f;
  if (x > 0.f) {
    if (y > 0.f) {
      if (x > y) return LUT[(int)(255.f * y / x)];
      return PI2f - LUT[(int)(255 * x / y)];
    } else {
      float absy = std::fabs(y);
      if (x > absy) return -LUT[(int)(255.f * absy / x)];
      return -PI2f + LUT[(int)(255.f * x / absy)];
    }
  } else if (y > 0.f) {
    float absx = std::fabs(x);
    if (absx > y) return PIf - LUT[(int)(255.f * y / absx)];
    return PI2f + LUT[(int)(255.f * absx / y)];
  } else {
    float absx = std::fabs(x), absy = std::fabs(y);
    if (absx > absy) return -PIf + LUT[(int)(255.f * absy / absx)];
    if (x == 0.f) return 0.f;
    return -PI2f - LUT[(int)(255.f * absx / absy)];
  }
}

extern "C" void orc_circular_gauss_mask(float* mask, int size, float sigma) {
  const int halfSize = size >> 1;
  const float r2 = float(halfSize * halfSize);
  const float sigma2 = sigma == 0 ? 0.9f * r2 : 2 * sigma * sigma;
  float* mp = mask;
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++) {
      const float disq = float((i - halfSize) * (i - halfSize) + (j - halfSize) * (j - halfSize));
      *mp++ = (disq < r2) ? std::exp(-disq / sigma2) : 0;     /* float overload, as <cmath> resolves it */
    }
}

extern "C" void orc_dominant_orientation(const float* img, int w, int h, const orc_region* regs, int n, double mrSize,
                                         int patchSize, int maxAngles, double th, int* n_ang, float* angles) {
  const int pS = patchSize, bins = 36;
  double LUT[256];
  orc_atan_lut(LUT);
  std::vector<float> orimask((size_t)pS * pS), patch((size_t)pS * pS), gmag((size_t)pS * pS, 0.f), gori((size_t)pS * pS, 0.f);
  orc_circular_gauss_mask(orimask.data(), pS, pS / 3.0f);
  const double mrScale = (double)mrSize;
  const int patchImageSize = 2 * int(mrScale) + 1;
  const double imageToPatchScale = double(patchImageSize) / (double)patchSize;
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
  for (int i = 0; i < n; i++) {
    const orc_region& k = regs[i];
    n_ang[i] = 0;
    const float curr_sc = imageToPatchScale * k.s;
    if (orc_interpolate_check_borders(w, h, (float)k.x, (float)k.y, (float)k.a11, (float)k.a12, (float)k.a21, (float)k.a22,
                                      (int)(k_sigma * k.s), (int)(k_sigma * k.s))) { n_ang[i] = -1; continue; }
    if (maxAngles <= 0) continue;
    orc_interpolate(img, w, h, (float)k.x, (float)k.y, (float)k.a11 * curr_sc, (float)k.a12 * curr_sc, (float)k.a21 * curr_sc,
                    (float)k.a22 * curr_sc, patch.data(), pS, pS);
    /* computeGradientMagnitudeAndOrientation (helpers.cpp:840-863): interior only, the frame keeps its zeros */
    for (int r = 1; r < pS - 1; ++r)
      for (int c = 1; c < pS - 1; ++c) {
        const float xgrad = patch[r * pS + c + 1] - patch[r * pS + c - 1];
        const float ygrad = patch[(r + 1) * pS + c] - patch[(r - 1) * pS + c];
        gmag[r * pS + c] = std::sqrt(xgrad * xgrad + ygrad * ygrad);
        gori[r * pS + c] = atan2LUTff(LUT, ygrad, xgrad);
      }
    float hist[bins + 1];
    for (int b = 0; b < bins; b++) hist[b] = 0.0f;
    const int maskPixels = pS * (pS - 2);
    for (int q = 0; q < maskPixels; ++q) {
      const float m = orimask[pS + q], g = gmag[pS + q], o = gori[pS + q];
      if (m > 0 && g > 1.0) {
        int bin = (int)(bins * (o / float(M_PI) + 1.0f) / 2.0f);
        hist[bin] += g * m;
      }
    }
    for (int it = 0; it < 6; it++) {   /* smoothCircularBuffer, synth-detection.cpp:811-822 */
      float first = hist[0], prev = hist[bins - 1];
      for (int b = 0; b < bins - 1; b++) { float cur = hist[b]; hist[b] = prev + cur + hist[b + 1]; prev = cur; }
      hist[bins - 1] = prev + hist[bins - 1] + first;
    }
    float thresh = 0.0;
    for (int b = 0; b < bins; b++) if (hist[b] > thresh) thresh = hist[b];
    thresh *= th;
    std::vector<float> ang, peaks;
    auto addPeak = [&](int a, int b, int c) {
      if (hist[b] >= thresh && hist[b] > hist[a] && hist[b] > hist[c]) {
        float pp = (hist[a] - hist[c]) / (hist[a] - 2.0f * hist[b] + hist[c]) / 2.0f;
        ang.push_back(2.0f * float(M_PI) * (b + 0.5f + pp) / bins - float(M_PI));
        peaks.push_back(hist[b]);
      }
    };
    addPeak(bins - 1, 0, 1);
    for (int b = 1; b < bins - 1; b++) addPeak(b - 1, b, b + 1);
    addPeak(bins - 2, bins - 1, 0);
    int mA = std::min(maxAngles, (int)peaks.size());
    int cnt = 0;
    for (int a = 0; a < mA; a++) {
      if (peaks[a] >= thresh) angles[(size_t)i * maxAngles + cnt++] = ang[a];
      else break;
    }
    n_ang[i] = cnt;
  }
}

namespace {
struct SiftTables {
  int ps = 0;
  std::vector<int> bin0, bin1;
  std::vector<double> w0, w1;
  std::vector<float> mask;
  void init(int patchSize) {   /* siftdesc.cpp:22-71 precomputeBinsAndWeights (spatialBins 4, orientationBins 8) */
    ps = patchSize;
    const int spatialBins = 4, orientationBins = 8;
    const int halfSize = ps >> 1;
    const float step = float(spatialBins + 1) / (2 * halfSize);
    bin0.resize(ps); bin1.resize(ps); w0.resize(ps); w1.resize(ps);
    for (int i = 0; i < ps; i++) {
      float x = step * i;
      int xi = (int)(x);
      bin0[i] = xi - 1; bin1[i] = xi;
      w1[i] = x - xi;
      w0[i] = 1.0f - w1[i];
      if (bin0[i] < 0) { bin0[i] = 0; w0[i] = 0; }
      if (bin0[i] >= spatialBins) { bin0[i] = spatialBins - 1; w0[i] = 0; }
      if (bin1[i] < 0) { bin1[i] = 0; w1[i] = 0; }
      if (bin1[i] >= spatialBins) { bin1[i] = spatialBins - 1; w1[i] = 0; }
      bin0[i] *= orientationBins; bin1[i] *= orientationBins;
    }
    mask.resize((size_t)ps * ps);
    orc_circular_gauss_mask(mask.data(), ps, 0.f);
  }
};
double normalize_d(std::vector<double>& v) {   /* siftdesc.cpp:132-159 */
  double len = 0.0;
  for (size_t i = 0; i < v.size(); i += 4) {
    const double sq0 = v[i] * v[i], sq1 = v[i + 1] * v[i + 1], sq2 = v[i + 2] * v[i + 2], sq3 = v[i + 3] * v[i + 3];
    len += sq0 + sq1 + sq2 + sq3;
  }
  len = std::sqrt(len);
  const double fac = 1.0 / len;
  for (size_t i = 0; i < v.size(); i++) v[i] *= fac;
  return len;
}
}  // namespace

extern "C" void orc_describe_sift(const float* img, int w, int h, const orc_region* regs, int n, double mrSize, int patchSize,
                                  int photoNorm, int rootSift, float* desc) {
  const int ps = patchSize, spatialBins = 4, orientationBins = 8;
  const double maxBinValue = 0.2;       /* [SIFTDescriptor] maxBinValue = 0.2 read with GetDouble (io_mods.cpp:427) */
  const double M_PI_DOUBLED = 6.28318530718;
  double LUT[256];
  orc_atan_lut(LUT);
  SiftTables T;
  T.init(ps);
  std::vector<float> pmask((size_t)ps * ps);
  orc_circular_gauss_mask(pmask.data(), ps, 0.f);   /* DescribeRegions' own mask (synth-detection.hpp:182-183) */
  std::vector<float> patch((size_t)ps * ps), grad((size_t)ps * ps), ori((size_t)ps * ps);
  for (int i = 0; i < n; i++) {
    orc_extract_patches(img, w, h, regs + i, 1, mrSize, ps, patch.data());
    if (photoNorm) {   /* helpers.cpp:666-716 */
      float sum = 0, gsum = 0;
      for (int q = 0; q < ps * ps; q++) if (pmask[q] > 0) { sum += patch[q]; gsum++; }
      sum = sum / gsum;
      float var = 0;
      for (int q = 0; q < ps * ps; q++) if (pmask[q] > 0) var += (sum - patch[q]) * (sum - patch[q]);
      var = std::sqrt(var / gsum);
      if (!(var < 0.0001)) {
        float fac = 50.0f / var;
        for (int q = 0; q < ps * ps; q++) {
          float v = 128 + fac * (patch[q] - sum);
          if (v > 255) v = 255;
          if (v < 0) v = 0;
          patch[q] = v;
        }
      }
    }
    /* gradients with one-sided differences on the frame (siftdesc.cpp:279-302) */
    for (int r = 0; r < ps; ++r)
      for (int c = 0; c < ps; ++c) {
        float xgrad, ygrad;
        if (c == 0) xgrad = patch[r * ps + c + 1] - patch[r * ps + c];
        else if (c == ps - 1) xgrad = patch[r * ps + c] - patch[r * ps + c - 1];
        else xgrad = patch[r * ps + c + 1] - patch[r * ps + c - 1];
        if (r == 0) ygrad = patch[(r + 1) * ps + c] - patch[r * ps + c];
        else if (r == ps - 1) ygrad = patch[r * ps + c] - patch[(r - 1) * ps + c];
        else ygrad = patch[(r + 1) * ps + c] - patch[(r - 1) * ps + c];
        grad[r * ps + c] = std::sqrt(xgrad * xgrad + ygrad * ygrad);
        ori[r * ps + c] = atan2LUTff(LUT, ygrad, xgrad);
      }
    /* samplePatch (siftdesc.cpp:73-130): vec is double, the per-pixel weights are float */
    std::vector<double> vec((size_t)spatialBins * spatialBins * orientationBins, 0.0);
    const bool magnLess = false;
    for (int r = 0; r < ps; ++r) {
      const int br0 = spatialBins * T.bin0[r];
      const float wr0 = T.w0[r];
      const int br1 = spatialBins * T.bin1[r];
      const float wr1 = T.w1[r];
      for (int c = 0; c < ps; ++c) {
        float val = float(magnLess) * 1.0 + (1.0 - float(magnLess)) * T.mask[r * ps + c] * grad[r * ps + c];
        const int bc0 = T.bin0[c];
        const float wc0 = T.w0[c] * val;
        const int bc1 = T.bin1[c];
        const float wc1 = T.w1[c] * val;
        const float o = float(orientationBins) * (ori[r * ps + c] + M_PI_DOUBLED) / M_PI_DOUBLED;
        int bo0 = (int)o;
        const float wo1 = o - bo0;
        bo0 %= orientationBins;
        int bo1 = (bo0 + 1) % orientationBins;
        const float wo0 = 1.0f - wo1;
        val = wr0 * wc0;
        if (val > 0) { vec[br0 + bc0 + bo0] += val * wo0; vec[br0 + bc0 + bo1] += val * wo1; }
        val = wr0 * wc1;
        if (val > 0) { vec[br0 + bc1 + bo0] += val * wo0; vec[br0 + bc1 + bo1] += val * wo1; }
        val = wr1 * wc0;
        if (val > 0) { vec[br1 + bc0 + bo0] += val * wo0; vec[br1 + bc0 + bo1] += val * wo1; }
        val = wr1 * wc1;
        if (val > 0) { vec[br1 + bc1 + bo0] += val * wo0; vec[br1 + bc1 + bo1] += val * wo1; }
      }
    }
    /* SIFTnorm / RootSIFTnorm on doubles (siftdesc.cpp:196-249) */
    normalize_d(vec);
    bool changed = false;
    for (size_t q = 0; q < vec.size(); q++) if (vec[q] > maxBinValue) { vec[q] = maxBinValue; changed = true; }
    if (changed) normalize_d(vec);
    if (rootSift) {
      double sum = 0.;
      for (size_t q = 0; q < vec.size(); q++) sum += std::fabs(vec[q]);
      for (size_t q = 0; q < vec.size(); q++) vec[q] = std::sqrt(vec[q] / sum);
      for (size_t q = 0; q < vec.size(); q++) {
        int b = std::max(0, std::min((int)(512.0 * vec[q] + 0.5), 255));
        desc[(size_t)i * 128 + q] = (float)double(b);
      }
    } else {
      for (size_t q = 0; q < vec.size(); q++) {
        int b = std::max(0, std::min((int)(512.0f * vec[q] + 0.5), 255));
        desc[(size_t)i * 128 + q] = (float)double(b);
      }
    }
  }
}

/* ========================================================================== */
/* view synthesis                                                             */
/* ========================================================================== */

/* cv::warpAffine for CV_32F / INTER_LINEAR / BORDER_CONSTANT, OpenCV 4.x imgwarp.cpp: the forward matrix is
 * inverted in double, source coordinates are evaluated in 22.10 fixed point (AB_BITS = 10) with per-column
 * tables adelta/bdelta and rounded to 1/32 px (INTER_BITS = 5); remapBilinear then blends the 4 neighbours with
 * float weights (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx accumulated left to right without fusion.  Out-of-image
 * neighbours take the border value.  Pinned bit-exactly against cv2 4.13 (tests/golden/synth_pins.npz). */
extern "C" void orc_warp_affine(const float* in, int w, int h, const double* Min, float* out, int ow, int oh, float border) {
  double M[6];
  for (int i = 0; i < 6; i++) M[i] = Min[i];
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1. / D : 0;
  double A11 = M[4] * D, A22 = M[0] * D;
  M[0] = A11; M[1] *= -D;
  M[3] *= -D; M[4] = A22;
  double b1 = -M[0] * M[2] - M[1] * M[5];
  double b2 = -M[3] * M[2] - M[4] * M[5];
  M[2] = b1; M[5] = b2;
  const int AB_BITS = 10, AB_SCALE = 1 << AB_BITS, INTER_BITS = 5, INTER_TAB_SIZE = 1 << INTER_BITS;
  const int round_delta = AB_SCALE / INTER_TAB_SIZE / 2;
  std::vector<int> adelta(ow), bdelta(ow);
  for (int x = 0; x < ow; x++) {
    adelta[x] = (int)std::lrint(M[0] * x * AB_SCALE);
    bdelta[x] = (int)std::lrint(M[3] * x * AB_SCALE);
  }
  auto PX = [&](int yy, int xx) -> float {
    return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? in[(size_t)yy * w + xx] : border;
  };
  for (int y = 0; y < oh; y++) {
    const int X0 = (int)std::lrint((M[1] * y + M[2]) * AB_SCALE) + round_delta;
    const int Y0 = (int)std::lrint((M[4] * y + M[5]) * AB_SCALE) + round_delta;
    for (int x = 0; x < ow; x++) {
      const int X = (X0 + adelta[x]) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta[x]) >> (AB_BITS - INTER_BITS);
      const int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
      const float a = (float)(X & (INTER_TAB_SIZE - 1)) / 32.0f, b = (float)(Y & (INTER_TAB_SIZE - 1)) / 32.0f;
      const float w0 = (1.0f - b) * (1.0f - a), w1 = (1.0f - b) * a, w2 = b * (1.0f - a), w3 = b * a;
      float v = PX(sy, sx) * w0 + PX(sy, sx + 1) * w1;
      v = v + PX(sy + 1, sx) * w2;
      v = v + PX(sy + 1, sx + 1) * w3;
      out[(size_t)y * ow + x] = v;
    }
  }
}

static void gaussian_taps_d(int ks, double sigma, std::vector<float>& k) {
  const int r = ks / 2;
  std::vector<double> kd(ks);
  double sum = 0;
  for (int i = 0; i < ks; i++) { double x = i - r; kd[i] = std::exp(-x * x / (2.0 * sigma * sigma)); sum += kd[i]; }
  k.resize(ks);
  for (int i = 0; i < ks; i++) k[i] = (float)(kd[i] / sum);
}
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

/* cv::GaussianBlur with separate kernel sizes / sigmas and BORDER_REFLECT_101 (synth-detection.cpp:499).
 * Arithmetic order of OpenCV 4.13 sepFilter2D for CV_32F (pinned bit-exactly, tests/golden/synth_pins.npz):
 *   row, ks == 3: x < w&~1: fma(x0,k1,(x[+1]+x[-1])*k2)      else fma(x[+1]+x[-1],k2,x0*k1)
 *   row, ks == 5 / ks >= 7: as orc_gaussian_blur
 *   col, ks == 3: fma(up+down,k2,c*k1) everywhere;  ks >= 5: as orc_gaussian_blur (fused iff x < w&~7) */
extern "C" void orc_gaussian_blur_xy(const float* in, float* out, int w, int h, int kx, int ky, double sigma_x, double sigma_y) {
  std::vector<float> KX, KY;
  gaussian_taps_d(kx, sigma_x, KX);
  gaussian_taps_d(ky, sigma_y, KY);
  std::vector<float> tmp((size_t)w * h);
  {
    const int ks = kx, r = ks / 2;
    const float* k = KX.data();
    for (int y = 0; y < h; y++) {
      auto P = [&](int d) -> float { return in[(size_t)y * w + reflect101(d, w)]; };
      for (int x = 0; x < w; x++) {
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
          s = 0;
          for (int t = 0; t < ks; t++) s = fmaf(P(x + t - r), k[t], s);
        } else {
          const int nf = (ks - 1) % 4;
          s = P(x - r) * k[0];
          for (int t = 1; t < ks; t++) {
            if (t >= ks - nf) s = fmaf(P(x + t - r), k[t], s);
            else s = s + P(x + t - r) * k[t];
          }
        }
        tmp[(size_t)y * w + x] = s;
      }
    }
  }
  {
    const int ks = ky, r = ks / 2;
    const float* k = KY.data();
    const int wc = w & ~7;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        auto T = [&](int d) -> float { return tmp[(size_t)reflect101(d, h) * w + x]; };
        float s = T(y) * k[r];
        if (ks == 3 || x < wc)
          for (int t = 1; t <= r; t++) s = fmaf(T(y - t) + T(y + t), k[r + t], s);
        else
          for (int t = 1; t <= r; t++) s = s + (T(y - t) + T(y + t)) * k[r + t];
        out[(size_t)y * w + x] = s;
      }
  }
}

/* synth-detection.cpp:356-431: output size and H of the synthesised view */
static void synth_rot(int w, int h, double phi, int* wr, int* hr, double* R) {
  if ((phi >= 0) && (phi < M_PI / 2)) {
    *wr = (int)std::floor((0.5 + std::cos(phi) * w + std::sin(phi) * h));
    *hr = (int)std::floor((0.5 + std::sin(phi) * w + std::cos(phi) * h));
    R[0] = std::cos(phi); R[1] = std::sin(phi); R[2] = 0;
    R[3] = -std::sin(phi); R[4] = std::cos(phi); R[5] = std::floor(0.5 + std::sin(phi) * w);
  } else {
    *wr = (int)std::floor((0.5 - std::cos(phi) * w + std::sin(phi) * h));
    *hr = (int)std::floor((0.5 + std::sin(phi) * w - std::cos(phi) * h));
    R[0] = std::cos(phi); R[1] = std::sin(phi); R[2] = -std::floor(std::cos(phi) * w);
    R[3] = -std::sin(phi); R[4] = std::cos(phi); R[5] = std::floor(0.5 + (std::sin(phi) * w - std::cos(phi) * h));
  }
}
extern "C" int orc_synth_geometry(int w, int h, double tilt, double phi, double zoom, int* ow, int* oh, double* H) {
  bool vertical = false;
  if (tilt < 0) { tilt = -tilt; vertical = true; }
  const int zoomed = std::fabs(zoom - 1.0f) >= 0.05 ? 1 : 0;
  const int wS1 = (int)(w * zoom), hS1 = (int)(h * zoom);
  for (int i = 0; i < 9; i++) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if ((std::fabs(tilt - 1.) <= 0.1) && (std::abs((int)phi) <= 0.2) && (std::fabs(zoom - 1.) <= 0.1)) {   /* abs(phi): int abs, :366 */
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

/* synth-detection.cpp:324-518 (non-AREA_INTERP branch): rotate -> anisotropic anti-alias blur -> tilt/zoom */
extern "C" void orc_synth_view(const float* gray, int w, int h, double tilt, double phi, double zoom, double InitSigma,
                               int doBlur, float* out) {
  int ow, oh;
  double H[9];
  if (orc_synth_geometry(w, h, tilt, phi, zoom, &ow, &oh, H)) {
    std::memcpy(out, gray, sizeof(float) * (size_t)w * h);
    return;
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
  synth_rot(w, h, phi, &wr, &hr, R);
  std::vector<float> rot((size_t)wr * hr), blurred;
  orc_warp_affine(gray, w, h, R, rot.data(), wr, hr, 128.f);
  const float* src = rot.data();
  if (doBlur) {
    int kx = (int)std::floor(2.0 * 3.0 * sigma_x + 1.0);
    if (kx % 2 == 0) kx++;
    if (kx < 3) kx = 3;
    int ky = (int)std::floor(2.0 * 3.0 * sigma_y + 1.0);
    if (ky % 2 == 0) ky++;
    if (ky < 3) ky = 3;
    blurred.resize(rot.size());
    orc_gaussian_blur_xy(rot.data(), blurred.data(), wr, hr, kx, ky, sigma_x, sigma_y);
    src = blurred.data();
  }
  double Wm[6] = {0, 0, 0, 0, 0, 0};
  if (vertical) { Wm[0] = 1.0 / kH; Wm[4] = 1.0 / (tilt * kV); }
  else { Wm[0] = 1.0 / (tilt * kH); Wm[4] = 1.0 / kV; }
  orc_warp_affine(src, wr, hr, Wm, out, ow, oh, 128.f);
}

/* ========================================================================== */
/* detector                                                                   */
/* ========================================================================== */

/* helpers.cpp:309-368 solveLinear3x3 (pivoted Gauss, float) */
static void solveLinear3x3(float* A, float* b) {
  int i = 0;
  float* pr = A;
  float vp = std::fabs(A[0]);
  float tmp = std::fabs(A[3]);
  if (tmp > vp) { pr = A + 3; i = 1; vp = tmp; }
  if (std::fabs(A[6]) > vp) { pr = A + 6; i = 2; }
  if (pr != A) {
    std::swap(pr[0], A[0]); std::swap(pr[1], A[1]); std::swap(pr[2], A[2]);
    std::swap(b[i], b[0]);
  }
  vp = A[3] / A[0]; A[4] -= vp * A[1]; A[5] -= vp * A[2]; b[1] -= vp * b[0];
  vp = A[6] / A[0]; A[7] -= vp * A[1]; A[8] -= vp * A[2]; b[2] -= vp * b[0];
  if (std::fabs(A[4]) < std::fabs(A[7])) {
    std::swap(A[7], A[4]); std::swap(A[8], A[5]); std::swap(b[2], b[1]);
  }
  vp = A[7] / A[4]; A[8] -= vp * A[5]; b[2] -= vp * b[1];
  b[2] = (b[2]) / A[8];
  b[1] = (b[1] - A[5] * b[2]) / A[4];
  b[0] = (b[0] - A[2] * b[2] - A[1] * b[1]) / A[0];
}

/* helpers.cpp:413-440 computeGaussMask */
extern "C" void orc_gauss_mask(float* mask, int size) {
  const int halfSize = size >> 1;
  const float scale = float(halfSize) / 3.0f;
  const float scale2 = -2.0f * scale * scale;
  std::vector<float> tmp(halfSize + 1);
  for (int i = 0; i <= halfSize; i++) tmp[i] = std::exp((float(i * i) / scale2));
  const int endSize = int(std::ceil(scale * 5.0f) - halfSize);
  for (int i = 1; i < endSize; i++) tmp[halfSize - i] += std::exp((float((i + halfSize) * (i + halfSize)) / scale2));
  for (int i = 0; i <= halfSize; i++)
    for (int j = 0; j <= halfSize; j++) {
      const float v = tmp[i] * tmp[j];
      mask[(i + halfSize) * size + (-j + halfSize)] = v;
      mask[(-i + halfSize) * size + (j + halfSize)] = v;
      mask[(i + halfSize) * size + (j + halfSize)] = v;
      mask[(-i + halfSize) * size + (-j + halfSize)] = v;
    }
}

/* helpers.cpp:461-503 invSqrt */
static void invSqrt(float& a, float& b, float& c, float& l1, float& l2) {
  double t, r;
  if (b != 0) {
    r = double(c - a) / (2 * b);
    if (r >= 0) t = 1.0 / (r + std::sqrt(1 + r * r));
    else t = -1.0 / (-r + std::sqrt(1 + r * r));
    r = 1.0 / std::sqrt(1 + t * t);
    t = t * r;
  } else { r = 1; t = 0; }
  double x, z, d;
  x = 1.0 / std::sqrt(r * r * a - 2 * r * t * b + t * t * c);
  z = 1.0 / std::sqrt(t * t * a + 2 * r * t * b + r * r * c);
  d = std::sqrt(x * z);
  x /= d; z /= d;
  if (x < z) { l1 = float(z); l2 = float(x); }
  else { l1 = float(x); l2 = float(z); }
  a = float(r * r * x + t * t * z);
  b = float(-r * t * x + t * r * z);
  c = float(t * t * x + r * r * z);
}
/* helpers.cpp:504-515 */
static bool getEigenvaluesF(float a, float b, float c, float d, float& l1, float& l2) {
  float trace = a + d;
  float delta1 = (trace * trace - 4 * (a * d - b * c));
  if (delta1 < 0) return false;
  float delta = std::sqrt(delta1);
  l1 = (trace + delta) / 2.0f;
  l2 = (trace - delta) / 2.0f;
  return true;
}

/* affine.cpp:26-158 findAffineShape, doBaumberg = 1, AFF_BMBRG_SMM.  blur = the pyramid level handed over by
 * localizeKeypoint (prevBlur).  Returns true and U when the iteration converged. */
static bool findAffineShape(const float* blur, int w, int h, float x, float y, float s, float pixelDistance,
                            const orc_affshape_params& par, const std::vector<float>& mask, float* U) {
  float eigen_ratio_act = 0.0f, eigen_ratio_bef = 0.0f;
  float u11 = 1.0f, u12 = 0.0f, u21 = 0.0f, u22 = 1.0f, l1 = 1.0f, l2 = 1.0f;
  const float lx = x / pixelDistance, ly = y / pixelDistance;
  const float ratio = s / (par.initialSigma * pixelDistance);
  const int ws = par.smmWindowSize, maskPixels = ws * ws;
  std::vector<float> img((size_t)maskPixels), fx((size_t)maskPixels), fy((size_t)maskPixels);
  for (int l = 0; l < par.maxIterations; l++) {
    float a = 0, b = 0, c = 0;
    orc_interpolate(blur, w, h, lx, ly, u11 * ratio, u12 * ratio, u21 * ratio, u22 * ratio, img.data(), ws, ws);
    for (int r = 0; r < ws; ++r)       /* computeGradient, helpers.cpp:779-797 */
      for (int cc = 0; cc < ws; ++cc) {
        float xgrad, ygrad;
        if (cc == 0) xgrad = img[r * ws + cc + 1] - img[r * ws + cc];
        else if (cc == ws - 1) xgrad = img[r * ws + cc] - img[r * ws + cc - 1];
        else xgrad = img[r * ws + cc + 1] - img[r * ws + cc - 1];
        if (r == 0) ygrad = img[(r + 1) * ws + cc] - img[r * ws + cc];
        else if (r == ws - 1) ygrad = img[r * ws + cc] - img[(r - 1) * ws + cc];
        else ygrad = img[(r + 1) * ws + cc] - img[(r - 1) * ws + cc];
        fx[r * ws + cc] = xgrad; fy[r * ws + cc] = ygrad;
      }
    for (int i = 0; i < maskPixels; ++i) {
      const float v = mask[i], gxx = fx[i], gyy = fy[i], gxy = gxx * gyy;
      a += gxx * gxx * v;
      b += gxy * v;
      c += gyy * gyy * v;
    }
    a /= maskPixels; b /= maskPixels; c /= maskPixels;
    invSqrt(a, b, c, l1, l2);
    if ((a != a) || (b != b) || (c != c)) break;
    eigen_ratio_bef = eigen_ratio_act;
    eigen_ratio_act = 1.0 - l2 / l1;
    float u11t = u11, u12t = u12;
    u11 = a * u11t + b * u21;
    u12 = a * u12t + b * u22;
    u21 = b * u11t + c * u21;
    u22 = b * u12t + c * u22;
    if (!getEigenvaluesF(u11, u12, u21, u22, l1, l2)) break;
    if ((l1 / l2 > 6) || (l2 / l1 > 6)) break;
    if (eigen_ratio_act < par.convergenceThreshold && eigen_ratio_bef < par.convergenceThreshold) {
      U[0] = u11; U[1] = u12; U[2] = u21; U[3] = u22;
      return true;
    }
  }
  return false;
}

namespace {
struct Detector {
  orc_pyr_params P;
  bool doBaumberg = false;
  orc_affshape_params AP;
  std::vector<float> smmMask;
  std::vector<float> keyA;
  /* pyramid.h:46-66 derived constants */
  double edgeScoreThreshold;
  float finalThreshold, positiveThreshold, negativeThreshold;
  int w = 0, h = 0;
  std::vector<float> low, cur, high, blur, prevBlur;
  std::vector<unsigned char> octaveMap;
  std::vector<orc_keypoint> keys;
  int octave = 0, level = 0;

  explicit Detector(const orc_pyr_params& p) : P(p) {
    edgeScoreThreshold = (p.edgeEigenValueRatio + 1.0f) * (p.edgeEigenValueRatio + 1.0f) / p.edgeEigenValueRatio;
    finalThreshold = p.threshold;
    positiveThreshold = (float)(0.8 * finalThreshold);
    negativeThreshold = -positiveThreshold;
    finalThreshold = p.threshold * p.threshold; /* DET_HESSIAN, FIXED_TH */
  }

  /* pyramid.cpp:41-63 */
  bool isMax(float val, const std::vector<float>& pix, int row, int col) const {
    for (int r = row - 1; r <= row + 1; r++)
      for (int c = col - 1; c <= col + 1; c++)
        if (pix[(size_t)r * w + c] > val) return false;
    return true;
  }
  bool isMin(float val, const std::vector<float>& pix, int row, int col) const {
    for (int r = row - 1; r <= row + 1; r++)
      for (int c = col - 1; c <= col + 1; c++)
        if (pix[(size_t)r * w + c] < val) return false;
    return true;
  }

  /* pyramid.cpp:281-403 */
  void localizeKeypoint(int r, int c, float curScale, float pixelDistance) {
    const int cols = w, rows = h;
    const int r0 = r, c0 = c;
    float b[3] = {};
    float val = 0;
    int nr = r, nc = c;
    for (int iter = 0; iter < 5; iter++) {
      r = nr; c = nc;
      const float* cur0 = &cur[(size_t)(r - 1) * w]; const float* cur1 = &cur[(size_t)r * w]; const float* cur2 = &cur[(size_t)(r + 1) * w];
      const float* low0 = &low[(size_t)(r - 1) * w]; const float* low1 = &low[(size_t)r * w]; const float* low2 = &low[(size_t)(r + 1) * w];
      const float* high0 = &high[(size_t)(r - 1) * w]; const float* high1 = &high[(size_t)r * w]; const float* high2 = &high[(size_t)(r + 1) * w];
      float dxx = cur1[c - 1] - 2.0f * cur1[c] + cur1[c + 1];
      float dyy = cur0[c] - 2.0f * cur1[c] + cur2[c];
      float dss = low1[c] - 2.0f * cur1[c] + high1[c];
      float dxy = 0.25f * (cur2[c + 1] - cur2[c - 1] - cur0[c + 1] + cur0[c - 1]);
      if (0 == iter) {
        float edgeScore = (dxx + dyy) * (dxx + dyy) / (dxx * dyy - dxy * dxy);
        if (edgeScore >= edgeScoreThreshold || edgeScore < 0) return;
      }
      float dxs = 0.25f * (high1[c + 1] - high1[c - 1] - low1[c + 1] + low1[c - 1]);
      float dys = 0.25f * (high2[c] - high0[c] - low2[c] + low0[c]);
      float A[9] = {dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss};
      float dx = 0.5f * (cur1[c + 1] - cur1[c - 1]);
      float dy = 0.5f * (cur2[c] - cur0[c]);
      float ds = 0.5f * (high1[c] - low1[c]);
      b[0] = -dx; b[1] = -dy; b[2] = -ds;
      solveLinear3x3(A, b);
      if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2])) return;
      val = cur1[c] + 0.5f * (dx * b[0] + dy * b[1] + ds * b[2]);
      /* MAX_SUBPIXEL_SHIFT 0.6 is a double literal (pyramid.cpp:25): float promoted */
      if (b[0] > 0.6) { if (c < cols - 3) nc++; else return; }
      if (b[1] > 0.6) { if (r < rows - 3) nr++; else return; }
      if (b[0] < -0.6) { if (c > 3) nc--; else return; }
      if (b[1] < -0.6) { if (r > 3) nr--; else return; }
      if (nr == r && nc == c) break;
    }
    if (std::fabs(b[0]) > 1.5 || std::fabs(b[1]) > 1.5 || std::fabs(b[2]) > 1.5 ||
        std::fabs(val) < finalThreshold || octaveMap[(size_t)r * w + c] > 0)
      return;
    octaveMap[(size_t)r * w + c] = 1;
    /* pyramid.cpp:392: curScale * pow(2.0f, b[2]/numberOfScales)  -- powf in the reference;
     * the oracle (and the kernel) evaluate 2^e in double and round once. */
    float e = b[2] / P.numberOfScales;
    float scale = curScale * (float)std::exp2((double)e);
    /* pyramid.cpp:65-124 getPointType on `blur` (the image of `cur`) */
    int type;
    if (val < 0) type = 2;
    else {
      const float* ptr = &blur[(size_t)r * w + c];
      float Lxx = (ptr[-1] - 2 * ptr[0] + ptr[1]);
      type = (Lxx < 0) ? 0 : 1;
    }
    orc_keypoint k;
    k.x = pixelDistance * (c + b[0]);
    k.y = pixelDistance * (r + b[1]);
    k.s = pixelDistance * scale;
    if (doBaumberg) {   /* onKeypointDetected -> findAffineShape(prevBlur, ...) (scale-space-detector.hpp:47-55) */
      float U[4];
      if (!findAffineShape(prevBlur.data(), w, h, k.x, k.y, k.s, pixelDistance, AP, smmMask, U)) return;
      keyA.insert(keyA.end(), U, U + 4);
    }
    k.response = val;
    k.type = type;
    k.octave = octave; k.level = level;
    k.r0 = r0; k.c0 = c0; k.r = r; k.c = c;
    k.seq = (int)keys.size();
    keys.push_back(k);
  }

  /* pyramid.cpp:405-425 */
  void findLevelKeypoints(float curScale, float pixelDistance) {
    for (int r = P.border; r < (h - P.border); r++)
      for (int c = P.border; c < (w - P.border); c++) {
        const float val = cur[(size_t)r * w + c];
        if ((val > positiveThreshold && (isMax(val, cur, r, c) && isMax(val, low, r, c) && isMax(val, high, r, c))) ||
            (val < negativeThreshold && (isMin(val, cur, r, c) && isMin(val, low, r, c) && isMin(val, high, r, c))))
          localizeKeypoint(r, c, curScale, pixelDistance);
      }
  }

  /* pyramid.cpp:428-494 */
  void detectOctave(const std::vector<float>& firstLevel, int W, int H, float pixelDistance,
                    std::vector<float>& nextFirst, int& nW, int& nH) {
    w = W; h = H;
    octaveMap.assign((size_t)w * h, 0);
    float sigmaStep = std::pow(2.0f, 1.0f / (float)P.numberOfScales);
    float curSigma = P.initialSigma;
    int numLevels = 1;
    blur = firstLevel;
    cur.resize((size_t)w * h); high.resize((size_t)w * h);
    orc_hessian_response(blur.data(), cur.data(), w, h, curSigma * curSigma);
    level = 0;
    for (int i = 1; i < P.numberOfScales + 2; i++) {
      float sigma = curSigma * std::sqrt(sigmaStep * sigmaStep - 1.0f);
      std::vector<float> nextBlur((size_t)w * h);
      orc_gaussian_blur(blur.data(), nextBlur.data(), w, h, sigma);
      sigma = curSigma * sigmaStep;
      orc_hessian_response(nextBlur.data(), high.data(), w, h, sigma * sigma);
      numLevels++;
      if (numLevels == 3) {
        level = i - 1;
        findLevelKeypoints(curSigma, pixelDistance);
        numLevels--;
      }
      if (i == P.numberOfScales) {
        orc_half_size(w, h, &nW, &nH);
        nextFirst.resize((size_t)nW * nH);
        orc_half_image(nextBlur.data(), w, h, nextFirst.data());
      }
      prevBlur.swap(blur);
      blur.swap(nextBlur);
      low.swap(cur);
      cur.swap(high);
      high.resize((size_t)w * h);
      curSigma *= sigmaStep;
    }
  }
};
}  // namespace

/* pyramid.cpp:496-529 detectPyramidKeypoints + scale-space-detector.hpp:120-131 sortKeys.
 * std::sort in the reference is unstable; the oracle fixes the total order
 * (|response| desc, push order asc) == what a stable sort yields. */
extern "C" int orc_detect_hessian(const float* gray, int w, int h, const orc_pyr_params* p,
                                  orc_keypoint* out, int cap) {
  Detector D(*p);
  float curSigma = 0.5f;
  float pixelDistance = 1.0f;
  std::vector<float> first(gray, gray + (size_t)w * h);
  if (p->initialSigma > curSigma) {
    float sigma = std::sqrt(p->initialSigma * p->initialSigma - curSigma * curSigma);
    std::vector<float> t((size_t)w * h);
    orc_gaussian_blur(first.data(), t.data(), w, h, sigma);
    first.swap(t);
  }
  int minSize = 2 * p->border + 2;
  int W = w, H = h;
  D.octave = 0;
  while (H > minSize && W > minSize) {
    std::vector<float> next;
    int nW = 0, nH = 0;
    D.detectOctave(first, W, H, pixelDistance, next, nW, nH);
    pixelDistance *= 2.0;
    first.swap(next);
    W = nW; H = nH;
    D.octave++;
  }
  std::stable_sort(D.keys.begin(), D.keys.end(), [](const orc_keypoint& a, const orc_keypoint& b) {
    return std::fabs(a.response) > std::fabs(b.response);
  });
  int n = (int)D.keys.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = D.keys[i];
  return n;
}

/* the same with the in-pyramid Baumberg iteration (doBaumberg = 1) */
extern "C" int orc_detect_hessian_affine(const float* gray, int w, int h, const orc_pyr_params* p, const orc_affshape_params* ap,
                                         orc_keypoint* out, float* A, int cap) {
  Detector D(*p);
  D.doBaumberg = true; D.AP = *ap;
  D.smmMask.resize((size_t)ap->smmWindowSize * ap->smmWindowSize);
  orc_gauss_mask(D.smmMask.data(), ap->smmWindowSize);
  float curSigma = 0.5f;
  float pixelDistance = 1.0f;
  std::vector<float> first(gray, gray + (size_t)w * h);
  if (p->initialSigma > curSigma) {
    float sigma = std::sqrt(p->initialSigma * p->initialSigma - curSigma * curSigma);
    std::vector<float> t((size_t)w * h);
    orc_gaussian_blur(first.data(), t.data(), w, h, sigma);
    first.swap(t);
  }
  int minSize = 2 * p->border + 2;
  int W = w, H = h;
  D.octave = 0;
  while (H > minSize && W > minSize) {
    std::vector<float> next;
    int nW = 0, nH = 0;
    D.detectOctave(first, W, H, pixelDistance, next, nW, nH);
    pixelDistance *= 2.0;
    first.swap(next);
    W = nW; H = nH;
    D.octave++;
  }
  std::stable_sort(D.keys.begin(), D.keys.end(), [](const orc_keypoint& a, const orc_keypoint& b) {
    return std::fabs(a.response) > std::fabs(b.response);
  });
  int n = (int)D.keys.size();
  for (int i = 0; i < n && i < cap; i++) {
    out[i] = D.keys[i];
    for (int k = 0; k < 4; k++) A[4 * i + k] = D.keyA[(size_t)4 * D.keys[i].seq + k];
  }
  return n;
}

/* ========================================================================== */
/* patch sampler                                                              */
/* ========================================================================== */

/* helpers.cpp:524-549 */
extern "C" int orc_interpolate_check_borders(int orig_img_w, int orig_img_h, float ofsx, float ofsy,
                                             float a11, float a12, float a21, float a22, int res_w, int res_h) {
  const int width = orig_img_w - 2;
  const int height = orig_img_h - 2;
  const float halfWidth = std::ceil((float)res_w / 2.0);
  const float halfHeight = std::ceil((float)res_h / 2.0);
  float x[4] = {-halfWidth, -halfWidth, +halfWidth, +halfWidth};
  float y[4] = {-halfHeight, +halfHeight, -halfHeight, +halfHeight};
  for (int i = 0; i < 4; i++) {
    float imx = ofsx + x[i] * a11 + y[i] * a12;
    float imy = ofsy + x[i] * a21 + y[i] * a22;
    if (std::floor(imx) <= 0 || std::floor(imy) <= 0 || std::ceil(imx) >= width || std::ceil(imy) >= height) return 1;
  }
  return 0;
}

/* helpers.cpp:551-626 interpolate(): bilinear affine resample; sample coordinates are
 * accumulated incrementally in float (WX += a11), which fixes the rounding. */
extern "C" int orc_interpolate(const float* im, int w, int h, float ofsx, float ofsy, float a11, float a12,
                               float a21, float a22, float* res, int res_w, int res_h) {
  bool ret = false;
  const int width = w - 1, height = h - 1;
  const int halfWidth = res_w / 2, halfHeight = res_h / 2;
  float* out = res;
  float rx = ofsx - (float)halfHeight * a12;
  float ry = ofsy - (float)halfHeight * a22;
  bool touch = orc_interpolate_check_borders(w, h, ofsx, ofsy, a11, a12, a21, a22, res_w, res_h);
  for (int j = -halfHeight; j < res_h - halfHeight; ++j) {
    float WX = rx - (float)halfWidth * a11;
    float WY = ry - (float)halfWidth * a21;
    for (int i = -halfWidth; i < res_w - halfWidth; ++i) {
      int x, y;
      bool inside;
      if (!touch) { x = (int)WX; y = (int)WY; inside = true; }
      else {
        x = (int)std::floor(WX); y = (int)std::floor(WY);
        inside = (WX >= 0 && WY >= 0 && x < width && y < height);
      }
      if (inside) {
        const float wx = WX - (float)x;
        const float* Row0 = im + (size_t)y * w;
        const float* Row1 = im + (size_t)(y + 1) * w;
        const float I1 = wx * (Row0[x + 1] - Row0[x]) + Row0[x];
        *out++ = (WY - y) * (wx * (Row1[x + 1] - Row1[x]) + Row1[x] - I1) + I1;
      } else {
        *out++ = 0;
        ret = true;
      }
      WX += a11;
      WY += a21;
    }
    rx += a12;
    ry += a22;
  }
  return ret;
}

/* synth-detection.cpp:38-132 ExtractPatchesColumn, fast_extraction=false, photoNorm=false */
extern "C" void orc_extract_patches(const float* img, int w, int h, const orc_region* regs, int n,
                                    double mrSize, int patchSize, float* outp) {
  std::vector<float> smoothed, blurred;
  for (int i = 0; i < n; i++) {
    float* roi = outp + (size_t)i * patchSize * patchSize;
    const orc_region& k = regs[i];
    float mrScale = std::ceil(k.s * mrSize); /* double product, ceil, -> float */
    int patchImageSize = patchSize % 2 != 0 ? 2 * int(mrScale) + 1 : 2 * int(mrScale);
    float imageToPatchScale = float(patchImageSize) / float(patchSize);
    if (imageToPatchScale > 0.4) {
      patchImageSize += 2;
      size_t np = (size_t)patchImageSize * patchImageSize;
      smoothed.resize(np); blurred.resize(np);
      orc_interpolate(img, w, h, (float)k.x, (float)k.y, (float)k.a11, (float)k.a12, (float)k.a21, (float)k.a22,
                      smoothed.data(), patchImageSize, patchImageSize);
      orc_gaussian_blur(smoothed.data(), blurred.data(), patchImageSize, patchImageSize, 1.5f * imageToPatchScale);
      orc_interpolate(blurred.data(), patchImageSize, patchImageSize, (float)(patchImageSize / 2), (float)(patchImageSize / 2),
                      imageToPatchScale, 0, 0, imageToPatchScale, roi, patchSize, patchSize);
    } else {
      orc_interpolate(img, w, h, (float)k.x, (float)k.y, (float)k.a11 * imageToPatchScale, (float)k.a12 * imageToPatchScale,
                      (float)k.a21 * imageToPatchScale, (float)k.a22 * imageToPatchScale, roi, patchSize, patchSize);
    }
  }
}

/* imagerepresentation.cpp:45 cv::imencode(".png", CV_32F Mat): OpenCV converts to 8 bit with
 * saturate_cast<uchar>(float) = cvRound (round-half-even) + clamp. */
extern "C" void orc_quantize_u8(const float* in, uint8_t* out, long n) {
  for (long i = 0; i < n; i++) {
    long v = std::lrintf(in[i]);
    out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

/* ========================================================================== */
/* region filters                                                             */
/* ========================================================================== */

/* helpers.cpp:378-387 rectifyAffineTransformationUpIsUp(double&...) */
static void rectifyUpIsUp(double& a11, double& a12, double& a21, double& a22) {
  double a = a11, b = a12, c = a21, d = a22;
  double det = std::sqrt(std::fabs(a * d - b * c));
  double b2a2 = std::sqrt(b * b + a * a);
  a11 = b2a2 / det;
  a12 = 0;
  a21 = (d * b + c * a) / (b2a2 * det);
  a22 = det / b2a2;
}
/* helpers.cpp:504-515 getEigenvalues (float) */
static bool getEigenvalues(float a, float b, float c, float d, float& l1, float& l2) {
  float trace = a + d;
  float delta1 = (trace * trace - 4 * (a * d - b * c));
  if (delta1 < 0) return false;
  float delta = std::sqrt(delta1);
  l1 = (trace + delta) / 2.0f;
  l2 = (trace - delta) / 2.0f;
  return true;
}

/* imagerepresentation.cpp:803-845 */
extern "C" int orc_affnet_postprocess(const orc_region* in, const float* aff3, int n, int w, int h,
                                      double mrSize, orc_region* out, int* src_index) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    orc_region t = in[i];
    t.a11 = aff3[3 * i + 0]; t.a12 = 0; t.a21 = aff3[3 * i + 1]; t.a22 = aff3[3 * i + 2];
    rectifyUpIsUp(t.a11, t.a12, t.a21, t.a22);
    float l1 = 1.0f, l2 = 1.0f;
    if (!getEigenvalues((float)t.a11, (float)t.a12, (float)t.a21, (float)t.a22, l1, l2)) continue;
    if ((l1 / l2 > 6) || (l2 / l1 > 6)) continue;
    /* double mrSize*s is passed into `const int res_w, res_h` (truncation) */
    if (orc_interpolate_check_borders(w, h, (float)t.x, (float)t.y, (float)t.a11, (float)t.a12, (float)t.a21, (float)t.a22,
                                      (int)(mrSize * t.s), (int)(mrSize * t.s)))
      continue;
    if (src_index) src_index[m] = i;
    out[m++] = t;
  }
  return m;
}

/* imagerepresentation.cpp:881-899 */
extern "C" void orc_orinet_postprocess(const orc_region* in, const float* ori2, int n, orc_region* out) {
  for (int i = 0; i < n; i++) {
    const orc_region& c = in[i];
    double angle = std::atan2((double)ori2[2 * i + 0], (double)ori2[2 * i + 1]);
    double ci = std::cos(angle), si = std::sin(angle);
    orc_region t = c;
    t.a11 = c.a11 * ci - c.a12 * si;
    t.a12 = c.a11 * si + c.a12 * ci;
    t.a21 = c.a21 * ci - c.a22 * si;
    t.a22 = c.a21 * si + c.a22 * ci;
    out[i] = t;
  }
}

/* synth-detection.cpp:631-706 ReprojectRegions for the identity view (H = I): keep when the
 * centre is strictly inside and the k_sigma*s frame (k_sigma = 2*3*sqrt(3), :21) does not
 * touch the border. */
extern "C" int orc_reproject_filter(const orc_region* in, int n, int w, int h, orc_region* out, int* src_index) {
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
  int m = 0;
  for (int i = 0; i < n; i++) {
    const orc_region& p = in[i];
    if ((p.x < w) && (p.y < h) && (p.x > 0) && (p.y > 0)) {
      if (!orc_interpolate_check_borders(w, h, (float)p.x, (float)p.y, (float)p.a11, (float)p.a12, (float)p.a21, (float)p.a22,
                                         (int)(k_sigma * p.s), (int)(k_sigma * p.s))) {
        if (src_index) src_index[m] = i;
        out[m++] = p;
      }
    }
  }
  return m;
}

/* ========================================================================== */
/* matching                                                                   */
/* ========================================================================== */

/* cvflann::L2<float>::operator() (OpenCV flann/dist.h, third-party; version unpinned by the
 * reference): float accumulation of squared differences in blocks of 4.  With the integer-
 * valued HardNet descriptors (0..255, dim 128) every partial sum is exactly representable,
 * so any summation order gives the same value. */
static float l2sq(const float* a, const float* b, int dim) {
  float result = 0;
  int i = 0;
  for (; i + 3 < dim; i += 4) {
    float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < dim; i++) { float d0 = a[i] - b[i]; result += d0 * d0; }
  return result;
}

/* cvflann::LinearIndex::findNeighbors + KNNSimpleResultSet::addPoint: visit train points in
 * index order; insert when dist < worst; equal distances keep insertion (index) order. */
extern "C" void orc_knn_linear(const float* q, int nq, const float* t, int nt, int dim, int nn, int* idx, float* dist) {
  for (int i = 0; i < nq; i++) {
    int* I = idx + (size_t)i * nn;
    float* Dd = dist + (size_t)i * nn;
    int count = 0;
    for (int j = 0; j < nn; j++) { I[j] = -1; Dd[j] = INFINITY; }
    float worst = INFINITY;
    for (int j = 0; j < nt; j++) {
      float d = l2sq(q + (size_t)i * dim, t + (size_t)j * dim, dim);
      if (d >= worst) continue;
      int k;
      for (k = count; k > 0; --k) {
        if (Dd[k - 1] > d) { if (k < nn) { Dd[k] = Dd[k - 1]; I[k] = I[k - 1]; } }
        else break;
      }
      if (count < nn) ++count;
      Dd[k] = d; I[k] = j;
      worst = Dd[nn - 1];
    }
  }
}

/* matching.cpp:356-460 MatchFlannFGINN, sqminratio < 1 branch (:430-457) */
extern "C" int orc_match_fginn(const float* q, const double* qxy, int nq, const float* t, const double* txy,
                               int nt, int dim, double ratio_thr, double contrad_dist, int nn, orc_match* out) {
  (void)qxy;
  if (nq == 0 || nt == 0) return 0;
  double sqminratio = ratio_thr * ratio_thr;
  double contrDistSq = contrad_dist * contrad_dist;
  std::vector<int> idx((size_t)nq * nn);
  std::vector<float> dist((size_t)nq * nn);
  orc_knn_linear(q, nq, t, nt, dim, nn, idx.data(), dist.data());
  int m = 0;
  for (int i = 0; i < nq; i++) {
    const int* I = &idx[(size_t)i * nn];
    const float* Dd = &dist[(size_t)i * nn];
    for (int j = 1; j < nn; j++) {
      if (I[j] < 0) break; /* fewer than nn train points (defined as "stop") */
      double ratio = Dd[0] / Dd[j]; /* float division, widened */
      if (sqminratio >= 1.0) { /* matching.cpp:395-428: every query yields a correspondence */
        double dx1 = txy[2 * I[0]] - txy[2 * I[j]], dy1 = txy[2 * I[0] + 1] - txy[2 * I[j] + 1];
        if ((j == nn - 1) || (dx1 * dx1 + dy1 * dy1 > contrDistSq)) {
          orc_match mt;
          mt.qi = i; mt.ti = I[0]; mt.tj_bad = I[j]; mt.d1 = Dd[0]; mt.d2 = Dd[j];
          mt.ratio = std::sqrt(ratio);
          out[m++] = mt;
          break;
        }
        continue;
      }
      if (ratio <= sqminratio) {
        orc_match mt;
        mt.qi = i; mt.ti = I[0]; mt.tj_bad = I[j]; mt.d1 = Dd[0]; mt.d2 = Dd[j];
        mt.ratio = std::sqrt(ratio);
        out[m++] = mt;
        break;
      }
      double dx = txy[2 * I[0]] - txy[2 * I[j]], dy = txy[2 * I[0] + 1] - txy[2 * I[j] + 1];
      if (dx * dx + dy * dy > contrDistSq) break;
    }
  }
  return m;
}

/* matching.cpp:2615-2679 DuplicateFiltering, mode MODE_FGINN (bestFGINN).  The reference's
 * std::sort is unstable; the oracle uses a stable sort by |ratio| (ties keep query order). */
extern "C" int orc_duplicate_filter(const double* xy1, const double* xy2, const double* ratio, int T,
                                    double r, int* order_out) {
  std::vector<int> ord(T);
  for (int i = 0; i < T; i++) ord[i] = i;
  if (r <= 0) { for (int i = 0; i < T; i++) order_out[i] = i; return T; }
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return std::fabs(ratio[a]) < std::fabs(ratio[b]); });
  double r_sq = r * r;
  std::vector<char> uniq(T, 1);
  for (int i = 0; i < T; i++) {
    if (!uniq[i]) continue;
    int a = ord[i];
    for (int j = i + 1; j < T; j++) {
      if (!uniq[j]) continue;
      int b = ord[j];
      double dx = xy1[2 * a] - xy1[2 * b], dy = xy1[2 * a + 1] - xy1[2 * b + 1];
      if (dx * dx + dy * dy > r_sq) continue;
      dx = xy2[2 * a] - xy2[2 * b]; dy = xy2[2 * a + 1] - xy2[2 * b + 1];
      if (dx * dx + dy * dy <= r_sq) uniq[j] = 0;
    }
  }
  int m = 0;
  for (int i = 0; i < T; i++) if (uniq[i]) order_out[m++] = ord[i];
  return m;
}
