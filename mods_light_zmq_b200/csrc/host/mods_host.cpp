// mods_host.cpp -- host mirror of the reference operators over the C ABI (see mods_host.h).
// Plain C++ (no CUDA): everything heavy is behind modsgpu_* calls; what remains here is the small
// per-region double/float arithmetic the reference also does on the host between its stages.
#include "mods_host.h"
#include "../npz.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cmath>
#include <cstring>
#include <fstream>
#include <thread>

namespace modsb200 {

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// CPU time of the calling thread (MODSGPU_HOST_PROFILE=1: per-stage host cost of a pair, printed on stderr)
static double cpu_ms() {
  timespec ts;
  clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static bool host_profile() {
  static const bool on = [] { const char* e = getenv("MODSGPU_HOST_PROFILE"); return e && atoi(e) != 0; }();
  return on;
}

// helpers.cpp:524-549.  NB callers pass doubles into the int res_w/res_h (truncation).
bool interpolateCheckBorders(int orig_img_w, int orig_img_h, float ofsx, float ofsy, float a11, float a12, float a21,
                             float a22, int res_w, int res_h) {
  const int width = orig_img_w - 2;
  const int height = orig_img_h - 2;
  const float halfWidth = std::ceil((float)res_w / 2.0);
  const float halfHeight = std::ceil((float)res_h / 2.0);
  const float x[4] = {-halfWidth, -halfWidth, +halfWidth, +halfWidth};
  const float y[4] = {-halfHeight, +halfHeight, -halfHeight, +halfHeight};
  for (int i = 0; i < 4; i++) {
    float imx = ofsx + x[i] * a11 + y[i] * a12;
    float imy = ofsy + x[i] * a21 + y[i] * a22;
    if (std::floor(imx) <= 0 || std::floor(imy) <= 0 || std::ceil(imx) >= width || std::ceil(imy) >= height) return true;
  }
  return false;
}

// helpers.cpp:401-410
void rectifyAffineTransformationUpIsUp(double& a11, double& a12, double& a21, double& a22) {
  double a = a11, b = a12, c = a21, d = a22;
  double det = std::sqrt(std::fabs(a * d - b * c));
  double b2a2 = std::sqrt(b * b + a * a);
  a11 = b2a2 / det;
  a12 = 0;
  a21 = (d * b + c * a) / (b2a2 * det);
  a22 = det / b2a2;
}

// helpers.cpp:504-515
bool getEigenvalues(float a, float b, float c, float d, float& l1, float& l2) {
  float trace = a + d;
  float delta1 = (trace * trace - 4 * (a * d - b * c));
  if (delta1 < 0) return false;
  float delta = std::sqrt(delta1);
  l1 = (trace + delta) / 2.0f;
  l2 = (trace - delta) / 2.0f;
  return true;
}

ImageRepresentation::ImageRepresentation(modsgpu_ctx* ctx, modsgpu_image* img, bool owns_image)
    : ctx_(ctx), img_(img), owns_(owns_image) {}
ImageRepresentation::~ImageRepresentation() {
  if (devdesc_) modsgpu_devdesc_free(ctx_, devdesc_);
  if (owns_ && img_) modsgpu_image_free(ctx_, img_);
}

static void to_regions(const AffineRegionVector& v, std::vector<modsgpu_region>& r) {
  r.resize(v.size());
  for (size_t i = 0; i < v.size(); i++) {
    const AffineKeypoint& k = v[i].det_kp;
    r[i].x = k.x; r[i].y = k.y; r[i].s = k.s; r[i].a11 = k.a11; r[i].a12 = k.a12; r[i].a21 = k.a21; r[i].a22 = k.a22;
  }
}

static bool invert3h(const double* A, double* R) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0 || !std::isfinite(det)) { for (int i = 0; i < 9; i++) R[i] = 0; return false; }
  const double id = 1.0 / det;
  R[0] = c0 * id; R[1] = (A[2] * A[7] - A[1] * A[8]) * id; R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c1 * id; R[4] = (A[0] * A[8] - A[2] * A[6]) * id; R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c2 * id; R[7] = (A[1] * A[6] - A[0] * A[7]) * id; R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}
// synth-detection.cpp:143-149
static bool HIsEye(const double* H) {
  const double eps1 = 0.01;
  return (std::fabs(H[0] - 1.0) + std::fabs(H[1]) + std::fabs(H[2]) + std::fabs(H[3]) + std::fabs(H[4] - 1.0) + std::fabs(H[5]) +
          std::fabs(H[6]) + std::fabs(H[7]) + std::fabs(H[8] - 1.0) < eps1);
}
// synth-detection.cpp:578-587
static void ReprojectByH(const AffineKeypoint& in_kp, AffineKeypoint& out_kp, const double* H) {
  out_kp.x = (H[0] * in_kp.x + H[1] * in_kp.y + H[2]);
  out_kp.y = (H[3] * in_kp.x + H[4] * in_kp.y + H[5]);
  out_kp.a11 = (H[0] * in_kp.a11 + H[1] * in_kp.a21);
  out_kp.a12 = (H[0] * in_kp.a12 + H[1] * in_kp.a22);
  out_kp.a21 = (H[3] * in_kp.a11 + H[4] * in_kp.a21);
  out_kp.a22 = (H[3] * in_kp.a12 + H[4] * in_kp.a22);
}

// synth-detection.cpp:191-322
int SetVSPars(const std::vector<double>& scale_set, const std::vector<double>& tilt_set, double phi_base,
              std::vector<ViewSynthParameters>& par, std::vector<ViewSynthParameters>& prev_par, double InitSigma, int doBlur) {
  const double eps1 = 0.01;
  par.clear();
  std::vector<ViewSynthParameters> prev_par_tmp(prev_par), pars_tmp;
  auto mk = [&](double phi, double tilt, double zoom, int blur) {
    ViewSynthParameters t;
    t.phi = phi; t.tilt = tilt; t.zoom = zoom; t.InitSigma = InitSigma; t.doBlur = blur;
    return t;
  };
  if (scale_set.empty() || tilt_set.empty()) pars_tmp.push_back(mk(0, 0, 0, 0));
  for (size_t sc = 0; sc < scale_set.size(); sc++)
    for (size_t t = 0; t < tilt_set.size(); t++) {
      if (std::fabs(tilt_set[t] - 1) > eps1) {
        int n_rot1 = (int)std::floor(180.0 * tilt_set[t] / phi_base);
        double delta_phi = M_PI / n_rot1;
        if (n_rot1 < 0) {   // no rotation mode if negative, add vertical tilt
          n_rot1 = 1;
          delta_phi = 0;
          pars_tmp.push_back(mk(0, -tilt_set[t], scale_set[sc], doBlur));
        }
        for (int r = 0; r < n_rot1; r++) pars_tmp.push_back(mk(delta_phi * r, tilt_set[t], scale_set[sc], doBlur));
      } else {
        pars_tmp.push_back(mk(0, tilt_set[t], scale_set[sc], doBlur));
      }
    }
  for (const ViewSynthParameters& p : pars_tmp) {
    bool unique = true;
    for (const ViewSynthParameters& q : prev_par_tmp)
      if ((std::fabs(p.zoom - q.zoom) <= eps1) && (std::fabs(p.tilt - q.tilt) <= eps1) && (std::fabs(p.phi - q.phi) <= eps1)) { unique = false; break; }
    if (unique) par.push_back(p);
  }
  for (const ViewSynthParameters& p : par) prev_par_tmp.push_back(p);
  prev_par = prev_par_tmp;
  return (int)par.size();
}

// imagerepresentation.cpp:686-1104 for one detector (HessianAffine) and the identity view (H = I,
// synth-detection.cpp:366-377), deep configuration (config_aff_ori_desc_zeromq.ini + iters_HessianZMQ.ini).
int ImageRepresentation::SynthDetectDescribeKeypoints(const DetectPars& par) {
  int w, h;
  modsgpu_image_size(img_, &w, &h);
  regions_.clear();
  n_views = 1;
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  return DescribeView(img_, I3, w, h, par, regions_);
}

int ImageRepresentation::SynthDetectDescribeKeypoints(const std::vector<ViewSynthParameters>& views, const DetectPars& par) {
  regions_.clear();
  n_views = 0;
  n_keypoints = n_affine = 0;
  return AddViews(views, par);
}

int ImageRepresentation::AddViews(const std::vector<ViewSynthParameters>& views, const DetectPars& par) {
  int w, h;
  modsgpu_image_size(img_, &w, &h);
  const int view_base = n_views;
  for (size_t v = 0; v < views.size(); v++) {
    double t0 = now_ms();
    modsgpu_image* view = nullptr;
    double H[9];
    int rc = modsgpu_synth_view(ctx_, img_, views[v].tilt, views[v].phi, views[v].zoom, views[v].InitSigma, views[v].doBlur, &view, H);
    if (rc) return rc;
    TimeSpent.SynthTime += now_ms() - t0;
    AffineRegionVector one;
    const int kp0 = n_keypoints, af0 = n_affine;
    DetectPars vpar = par;     // DetectAffineKeypoints shrinks reg_number on tilted / zoomed-out views (scale-space-detector.cpp:19-21)
    vpar.pyr.reg_number = modsgpu_reg_number_for_view(par.pyr.reg_number, views[v].tilt, views[v].zoom);
    rc = DescribeView(view, H, w, h, vpar, one);
    modsgpu_image_free(ctx_, view);
    if (rc < 0) return rc;
    n_keypoints += kp0; n_affine += af0;
    for (AffineRegion& r : one) {
      r.img_reproj_id = view_base + (int)v;
      r.id = (int)regions_.size();
      regions_.push_back(r);
    }
    n_views++;
  }
  return (int)regions_.size();
}

// one view: detect -> AffNet -> reproject -> OriNet -> reproject + frame test -> HardNet++.  `view` is the
// synthesised image (or the original), H maps original -> view; regions keep det_kp in view coordinates and get
// reproj_kp in the original image.
int ImageRepresentation::DescribeView(modsgpu_image* view, const double* H, int orig_w, int orig_h, const DetectPars& par,
                                      AffineRegionVector& result) {
  int w, h;
  modsgpu_image_size(view, &w, &h);
  result.clear();
  // Default route: the whole view on the device (modsgpu_describe_view, csrc/chain.cu) -- same stages, same arithmetic,
  // two host synchronisations instead of four and no host-side list handling between the nets.  MODSGPU_SEAM_CHAIN=1
  // keeps the seam-by-seam route below (modsgpu_detect + 3 x modsgpu_describe), which the parity tests compare it with.
  const char* seam_env = getenv("MODSGPU_SEAM_CHAIN");      // read per view: the parity tests switch routes inside one process
  const bool seam_chain = seam_env && atoi(seam_env) != 0;
  if (!seam_chain && par.patchSize == 32) {
    double t0 = now_ms();
    modsgpu_view_region* rows = nullptr;
    float* desc = nullptr;
    int n = 0, counts[3] = {0, 0, 0};
    const bool on_device = par.desc_on_device && &result == &regions_;
    if (devdesc_) { modsgpu_devdesc_free(ctx_, devdesc_); devdesc_ = nullptr; }
    int rc = on_device ? modsgpu_describe_view_dev(ctx_, view, H, orig_w, orig_h, &par.pyr, par.mrSize, par.patchSize, &rows, &devdesc_, &n, counts)
                       : modsgpu_describe_view(ctx_, view, H, orig_w, orig_h, &par.pyr, par.mrSize, par.patchSize, &rows, &desc, &n, counts);
    if (rc) return rc;
    n_keypoints = counts[0];
    n_affine = counts[1];
    std::shared_ptr<const std::vector<float>> blk;
    if (!on_device) blk = std::make_shared<const std::vector<float>>(desc, desc + (size_t)n * 128);
    result.resize(n);
    for (int i = 0; i < n; i++) {
      AffineRegion& r = result[i];
      r.id = i; r.parent_id = -1; r.type = 1 /* DET_HESSIAN */;
      AffineKeypoint& k = r.det_kp;
      const modsgpu_region& d = rows[i].det;
      k.x = d.x; k.y = d.y; k.s = d.s; k.a11 = d.a11; k.a12 = d.a12; k.a21 = d.a21; k.a22 = d.a22;
      k.response = rows[i].response; k.octave_number = rows[i].octave; k.sub_type = rows[i].type;
      r.reproj_kp = k;
      const modsgpu_region& q = rows[i].reproj;
      r.reproj_kp.x = q.x; r.reproj_kp.y = q.y;
      r.reproj_kp.a11 = q.a11; r.reproj_kp.a12 = q.a12; r.reproj_kp.a21 = q.a21; r.reproj_kp.a22 = q.a22;
      if (!on_device) r.desc.view(blk, (size_t)i * 128, 128);
    }
    modsgpu_free(rows);
    if (desc) modsgpu_free(desc);
    TimeSpent.DescTime += now_ms() - t0;
    return n;
  }
  double Hinv[9];
  invert3h(H, Hinv);
  const bool eye = HIsEye(H);
  double t0 = now_ms();
  // ---- DetectAffineRegions (synth-detection.hpp:79-112) over DetectAffineKeypoints
  modsgpu_keypoint* kps = nullptr;
  int n = 0;
  const bool hp = host_profile();
  double hc[5] = {0, 0, 0, 0, 0}, hc0 = hp ? cpu_ms() : 0;
  int rc = modsgpu_detect(ctx_, view, &par.pyr, &kps, &n);
  if (hp) hc[0] = cpu_ms() - hc0;
  if (rc) return rc;
  n_keypoints = n;
  AffineRegionVector temp_kp1(n);
  for (int i = 0; i < n; i++) {
    AffineRegion& r = temp_kp1[i];
    r.id = i; r.parent_id = -1; r.type = 1 /* DET_HESSIAN */;
    AffineKeypoint& k = r.det_kp;
    k.x = kps[i].x; k.y = kps[i].y; k.s = kps[i].s;      // s * sqrt(|det I|), rectifyTransformation(I) = I
    k.a11 = 1; k.a12 = 0; k.a21 = 0; k.a22 = 1;
    k.response = kps[i].response; k.octave_number = kps[i].octave; k.sub_type = kps[i].type;
  }
  modsgpu_free(kps);
  std::vector<modsgpu_region> regs;
  std::vector<float> out;
  // ---- AffNet (imagerepresentation.cpp:797-845); the border test is against the VIEW
  to_regions(temp_kp1, regs);
  out.resize((size_t)n * 3 + 1);
  if (hp) hc0 = cpu_ms();
  rc = modsgpu_describe(ctx_, MODSGPU_AFFNET, view, regs.data(), n, par.mrSize, par.patchSize, out.data());
  if (hp) hc[1] = cpu_ms() - hc0;
  if (rc) return rc;
  AffineRegionVector temp_kp_aff;
  temp_kp_aff.reserve(n);
  for (int i = 0; i < n; i++) {
    // built in place at the back of the list and popped again when a test rejects it (one copy per region)
    temp_kp_aff.push_back(temp_kp1[i]);
    AffineRegion& t = temp_kp_aff.back();
    t.det_kp.a11 = out[3 * i + 0];
    t.det_kp.a12 = 0;
    t.det_kp.a21 = out[3 * i + 1];
    t.det_kp.a22 = out[3 * i + 2];
    rectifyAffineTransformationUpIsUp(t.det_kp.a11, t.det_kp.a12, t.det_kp.a21, t.det_kp.a22);
    float l1 = 1.0f, l2 = 1.0f;
    bool keep = getEigenvalues((float)t.det_kp.a11, (float)t.det_kp.a12, (float)t.det_kp.a21, (float)t.det_kp.a22, l1, l2);
    keep = keep && !((l1 / l2 > 6) || (l2 / l1 > 6));
    keep = keep && !interpolateCheckBorders(w, h, (float)t.det_kp.x, (float)t.det_kp.y, (float)t.det_kp.a11, (float)t.det_kp.a12,
                                            (float)t.det_kp.a21, (float)t.det_kp.a22, (int)(par.mrSize * t.det_kp.s),
                                            (int)(par.mrSize * t.det_kp.s));
    if (!keep) temp_kp_aff.pop_back();
  }
  n_affine = (int)temp_kp_aff.size();
  TimeSpent.DetectTime += now_ms() - t0;
  t0 = now_ms();
  // ---- ReprojectRegionsAndRemoveTouchBoundary(dontRemove = true) (synth-detection.cpp:151-190)
  // filtered in place (the survivors keep their order)
  {
    size_t nk = 0;
    for (auto& r : temp_kp_aff) {
      r.reproj_kp = r.det_kp;
      if (!eye) ReprojectByH(r.det_kp, r.reproj_kp, Hinv);
      if ((r.reproj_kp.x < orig_w) && (r.reproj_kp.y < orig_h) && (r.reproj_kp.x > 0) && (r.reproj_kp.y > 0)) {
        if (&temp_kp_aff[nk] != &r) temp_kp_aff[nk] = r;
        nk++;
      }
    }
    temp_kp_aff.resize(nk);
  }
  AffineRegionVector& kept = temp_kp_aff;
  // ---- OriNet (imagerepresentation.cpp:876-899), patches from the view
  const int n2 = (int)kept.size();
  to_regions(kept, regs);
  out.resize((size_t)n2 * 2 + 1);
  if (hp) hc0 = cpu_ms();
  rc = modsgpu_describe(ctx_, MODSGPU_ORINET, view, regs.data(), n2, par.mrSize, par.patchSize, out.data());
  if (hp) hc[2] = cpu_ms() - hc0;
  if (rc) return rc;
  // rotated in place: every new entry is computed from the OLD four entries of the same region
  for (int i = 0; i < n2; i++) {
    AffineKeypoint& k = kept[i].det_kp;
    double angle = std::atan2((double)out[2 * i + 0], (double)out[2 * i + 1]);
    double ci = std::cos(angle), si = std::sin(angle);
    const double a11 = k.a11, a12 = k.a12, a21 = k.a21, a22 = k.a22;
    k.a11 = a11 * ci - a12 * si;
    k.a12 = a11 * si + a12 * ci;
    k.a21 = a21 * ci - a22 * si;
    k.a22 = a21 * si + a22 * ci;
  }
  AffineRegionVector& oriented = kept;
  TimeSpent.OrientTime += now_ms() - t0;
  t0 = now_ms();
  // ---- ReprojectRegions (synth-detection.cpp:631-706): centre inside + k_sigma*s frame inside the ORIGINAL image
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);   // synth-detection.cpp:21
  {
    size_t nk = 0;
    for (auto& r : oriented) {
      r.reproj_kp = r.det_kp;
      if (!eye) ReprojectByH(r.det_kp, r.reproj_kp, Hinv);
      const AffineKeypoint& p = r.reproj_kp;
      if ((p.x < orig_w) && (p.y < orig_h) && (p.x > 0) && (p.y > 0) &&
          !interpolateCheckBorders(orig_w, orig_h, (float)p.x, (float)p.y, (float)p.a11, (float)p.a12, (float)p.a21, (float)p.a22,
                                   (int)(k_sigma * p.s), (int)(k_sigma * p.s))) {
        if (&oriented[nk] != &r) oriented[nk] = r;
        nk++;
      }
    }
    oriented.resize(nk);
  }
  AffineRegionVector& final_regs = oriented;
  // ---- HardNet++ (imagerepresentation.cpp:992-1006), patches from the view
  const int n3 = (int)final_regs.size();
  to_regions(final_regs, regs);
  out.resize((size_t)n3 * 128 + 1);
  if (hp) hc0 = cpu_ms();
  rc = modsgpu_describe(ctx_, MODSGPU_HARDNET, view, regs.data(), n3, par.mrSize, par.patchSize, out.data());
  if (hp) hc[3] = cpu_ms() - hc0;
  if (rc) return rc;
  {
    out.resize((size_t)n3 * 128);
    auto blk = std::make_shared<const std::vector<float>>(std::move(out));
    for (int i = 0; i < n3; i++) {
      final_regs[i].desc.view(blk, (size_t)i * 128, 128);
      final_regs[i].id = i;
    }
  }
  result.swap(final_regs);
  if (hp) fprintf(stderr, "[modsgpu host]   view: cpu ms in detect %.2f affnet %.2f orinet %.2f hardnet %.2f\n", hc[0], hc[1], hc[2], hc[3]);
  TimeSpent.DescTime += now_ms() - t0;
  return n3;
}

// synth-detection.cpp:134-143
static void rectifyTransformation(double& a11, double& a12, double& a21, double& a22) {
  double a = a11, b = a12, c = a21, d = a22;
  double det = std::sqrt(std::fabs(a * d - b * c));
  double b2a2 = std::sqrt(b * b + a * a);
  a11 = b2a2 / det;
  a12 = 0;
  a21 = (d * b + c * a) / (b2a2 * det);
  a22 = det / b2a2;
}

// imagerepresentation.cpp:686-1104 for config_affori_classic.ini + iters_HessianSIFT.ini (identity view):
// DetectAffineRegions over the Baumberg detector -> ReprojectRegionsAndRemoveTouchBoundary(dontRemove) ->
// DetectOrientation (:903) -> ReprojectRegions (:951) -> DescribeRegions<RootSIFT> (:958-965)
int ImageRepresentation::SynthDetectDescribeKeypointsClassic(const DetectPars& par) {
  int w, h;
  modsgpu_image_size(img_, &w, &h);
  regions_.clear();
  desc_name = "RootSIFT";
  n_views = 1;
  double t0 = now_ms();
  modsgpu_keypoint* kps = nullptr;
  float* A = nullptr;
  int n = 0;
  int rc = modsgpu_detect_affine(ctx_, img_, &par.pyr, &par.aff, &kps, &A, &n);
  if (rc) return rc;
  n_keypoints = n;
  AffineRegionVector temp_kp1(n);
  for (int i = 0; i < n; i++) {   // DetectAffineRegions, synth-detection.hpp:79-112
    AffineRegion& r = temp_kp1[i];
    r.id = i; r.parent_id = -1; r.type = 1 /* DET_HESSIAN */;
    AffineKeypoint& k = r.det_kp;
    double a11 = A[4 * i], a12 = A[4 * i + 1], a21 = A[4 * i + 2], a22 = A[4 * i + 3];
    k.s = kps[i].s * std::sqrt(std::fabs(a11 * a22 - a12 * a21));
    rectifyTransformation(a11, a12, a21, a22);
    k.x = kps[i].x; k.y = kps[i].y;
    k.a11 = a11; k.a12 = a12; k.a21 = a21; k.a22 = a22;
    k.response = kps[i].response; k.octave_number = kps[i].octave; k.sub_type = kps[i].type;
  }
  modsgpu_free(kps);
  modsgpu_free(A);
  n_affine = n;
  TimeSpent.DetectTime += now_ms() - t0;
  t0 = now_ms();
  // ReprojectRegionsAndRemoveTouchBoundary(dontRemove = true), H = I: centre inside
  AffineRegionVector kept;
  kept.reserve(n);
  for (auto& r : temp_kp1) {
    r.reproj_kp = r.det_kp;
    if ((r.reproj_kp.x < w) && (r.reproj_kp.y < h) && (r.reproj_kp.x > 0) && (r.reproj_kp.y > 0)) kept.push_back(r);
  }
  // DetectOrientation (synth-detection.cpp:1039-1149)
  std::vector<modsgpu_region> regs;
  to_regions(kept, regs);
  const int n2 = (int)kept.size();
  std::vector<int> n_ang(n2 + 1);
  std::vector<float> ang((size_t)n2 * par.maxAngles + 1);
  rc = modsgpu_dominant_orientation(ctx_, img_, regs.data(), n2, par.mrSize, par.oriPatchSize, par.maxAngles, par.oriThreshold,
                                    n_ang.data(), ang.data());
  if (rc) return rc;
  AffineRegionVector oriented;
  oriented.reserve(n2);
  int count = 0;
  for (int i = 0; i < n2; i++) {
    if (n_ang[i] < 0) continue;                       // frame touches the border
    AffineRegion c = kept[i];
    c.id = count;
    for (int j = 0; j < n_ang[i]; j++) {
      const double a = ang[(size_t)i * par.maxAngles + j];
      const double ci = std::cos(-a), si = std::sin(-a);
      AffineRegion t = c;
      t.det_kp.a11 = c.det_kp.a11 * ci - c.det_kp.a12 * si;
      t.det_kp.a12 = c.det_kp.a11 * si + c.det_kp.a12 * ci;
      t.det_kp.a21 = c.det_kp.a21 * ci - c.det_kp.a22 * si;
      t.det_kp.a22 = c.det_kp.a21 * si + c.det_kp.a22 * ci;
      oriented.push_back(t);
      count++;
    }
  }
  TimeSpent.OrientTime += now_ms() - t0;
  t0 = now_ms();
  // ReprojectRegions (synth-detection.cpp:631-706), H = I
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
  AffineRegionVector final_regs;
  final_regs.reserve(oriented.size());
  for (auto& r : oriented) {
    r.reproj_kp = r.det_kp;
    const AffineKeypoint& p = r.reproj_kp;
    if ((p.x < w) && (p.y < h) && (p.x > 0) && (p.y > 0)) {
      if (!interpolateCheckBorders(w, h, (float)p.x, (float)p.y, (float)p.a11, (float)p.a12, (float)p.a21, (float)p.a22,
                                   (int)(k_sigma * p.s), (int)(k_sigma * p.s)))
        final_regs.push_back(r);
    }
  }
  // DescribeRegions<SIFTDescriptor> (synth-detection.hpp:170-263)
  const int n3 = (int)final_regs.size();
  to_regions(final_regs, regs);
  std::vector<float> out((size_t)n3 * 128 + 1);
  rc = modsgpu_describe_sift(ctx_, img_, regs.data(), n3, par.mrSize, par.siftPatchSize, par.photoNorm, par.rootSift, out.data());
  if (rc) return rc;
  for (int i = 0; i < n3; i++) {
    final_regs[i].desc.assign(out.begin() + (size_t)i * 128, out.begin() + (size_t)(i + 1) * 128);
    final_regs[i].id = i;
  }
  regions_.swap(final_regs);
  TimeSpent.DescTime += now_ms() - t0;
  return n3;
}

// imagerepresentation.cpp:113-126: sc = s*sqrt(|det A|)*3*sqrt(3); A <- upIsUp(A) cast to float; SVD A = U W V^T;
// ellipse = U diag(1/(w_i^2 sc^2)) U^T = (A A^T)^-1 / sc^2.  cv::SVD (third-party, float) is replaced by the closed
// form evaluated in double on the float-cast A and rounded to float; the file prints 6 significant digits.
void OxAffEllipse(const AffineKeypoint& k, float& a, float& b, float& c) {
  double a11 = k.a11, a12 = k.a12, a21 = k.a21, a22 = k.a22;
  const double sc = k.s * std::sqrt(std::fabs(a11 * a22 - a12 * a21)) * 3.0 * std::sqrt(3.0);
  rectifyAffineTransformationUpIsUp(a11, a12, a21, a22);
  const double f11 = (float)a11, f12 = (float)a12, f21 = (float)a21, f22 = (float)a22;
  // M = A A^T
  const double m11 = f11 * f11 + f12 * f12, m12 = f11 * f21 + f12 * f22, m22 = f21 * f21 + f22 * f22;
  const double det = m11 * m22 - m12 * m12, q = 1.0 / (det * sc * sc);
  a = (float)(m22 * q); b = (float)(-m12 * q); c = (float)(m11 * q);
}

// operator<<(ostream, float / double) with the default format = printf("%g") (6 significant digits); descriptor entries
// are small integers almost always, which a table prints without any formatting call.  Byte-identical to the iostream
// form the reference writes (the writer tests compare with files produced through std::ofstream).
namespace {
struct NumText {
  char txt[256][4];
  unsigned char len[256];
  NumText() { for (int i = 0; i < 256; i++) len[i] = (unsigned char)snprintf(txt[i], 4, "%d", i); }
};
inline void put_g(std::string& out, double v) {
  char buf[40];
  const int n = snprintf(buf, sizeof(buf), "%g", v);
  out.append(buf, (size_t)n);
}
inline void put_desc(std::string& out, float v) {
  static const NumText T;
  const int iv = (int)v;
  if (v >= 0.f && v < 256.f && (float)iv == v) out.append(T.txt[iv], T.len[iv]);
  else put_g(out, (double)v);
}
}  // namespace

int SaveRegionsMichal(const AffineRegionVector& regions, const std::string& fname) {
  std::string out;
  out.reserve(regions.size() * 460 + 64);
  out += "128\n";
  out += std::to_string(regions.size());
  out += "\n";
  for (const AffineRegion& ar : regions) {
    float a, b, c;
    OxAffEllipse(ar.reproj_kp, a, b, c);
    put_g(out, ar.reproj_kp.x); out += ' ';
    put_g(out, ar.reproj_kp.y); out += ' ';
    put_g(out, (double)a); out += ' ';
    put_g(out, (double)b); out += ' ';
    put_g(out, (double)c); out += ' ';
    const size_t nd = ar.desc.size();
    const float* d = ar.desc.data();
    for (size_t i = 0; i < nd; i++) { put_desc(out, d[i]); out += ' '; }
    out += '\n';
  }
  FILE* f = fopen(fname.c_str(), "wb");
  if (!f) return -1;
  const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  return (fclose(f) == 0 && ok) ? 0 : -1;
}

// matching.cpp:356-460 (vector_matcher = linear): list1 = queries (image 1), list2 = train (image 2)
int MatchFlannFGINN(modsgpu_ctx* ctx, const AffineRegionVector& list1, const AffineRegionVector& list2,
                    TentativeCorrespListExt& corresp, const MatchPars& par) {
  corresp.TCList.clear();
  const int n1 = (int)list1.size(), n2 = (int)list2.size();
  if (n1 == 0 || n2 == 0) return 0;
  const int dim = (int)list1[0].desc.size();
  // descriptors of one describe call sit back to back in their block: hand that block over as it is
  auto flat = [dim](const AffineRegionVector& l, std::vector<float>& tmp) -> const float* {
    const std::vector<float>* b = l[0].desc.block();
    bool contiguous = b != nullptr;
    for (size_t i = 0; contiguous && i < l.size(); i++)
      contiguous = l[i].desc.block() == b && l[i].desc.size() == (size_t)dim && l[i].desc.offset() == l[0].desc.offset() + i * dim;
    if (contiguous) return l[0].desc.data();
    tmp.resize(l.size() * (size_t)dim);
    for (size_t i = 0; i < l.size(); i++) memcpy(&tmp[i * dim], l[i].desc.data(), dim * sizeof(float));
    return tmp.data();
  };
  std::vector<float> qtmp, ttmp;
  const float* q = flat(list1, qtmp);
  const float* t = flat(list2, ttmp);
  std::vector<double> txy((size_t)n2 * 2);
  for (int i = 0; i < n2; i++) { txy[2 * i] = list2[i].reproj_kp.x; txy[2 * i + 1] = list2[i].reproj_kp.y; }
  std::vector<modsgpu_match> m(n1);
  int nm = 0;
  int rc = modsgpu_match_fginn(ctx, q, n1, t, txy.data(), n2, dim, par.FGINNThreshold, par.contradDist,
                               par.nn, m.data(), &nm, nullptr, nullptr);
  if (rc) return rc;
  corresp.TCList.reserve(nm);
  for (int k = 0; k < nm; k++) {
    TentativeCorrespExt tc;
    tc.first = list1[m[k].qi];
    tc.second = list2[m[k].ti];
    tc.secondbad_idx = m[k].tj_bad;
    tc.d1 = m[k].d1; tc.d2 = m[k].d2; tc.ratio = m[k].ratio;
    corresp.TCList.push_back(tc);
  }
  return nm;
}

int MatchFlannFGINNDevice(modsgpu_ctx* ctx, const ImageRepresentation& img1, const ImageRepresentation& img2,
                          TentativeCorrespListExt& corresp, const MatchPars& par) {
  corresp.TCList.clear();
  const AffineRegionVector& list1 = img1.GetAffineRegionVector();
  const AffineRegionVector& list2 = img2.GetAffineRegionVector();
  const int n1 = (int)list1.size(), n2 = (int)list2.size();
  if (n1 == 0 || n2 == 0) return 0;
  const modsgpu_devdesc *q = img1.device_descriptors(), *t = img2.device_descriptors();
  if (!q || !t || modsgpu_devdesc_size(q) != n1 || modsgpu_devdesc_size(t) != n2) return MODSGPU_ESTATE;
  std::vector<double> txy((size_t)n2 * 2);
  for (int i = 0; i < n2; i++) { txy[2 * i] = list2[i].reproj_kp.x; txy[2 * i + 1] = list2[i].reproj_kp.y; }
  std::vector<modsgpu_match> m(n1);
  int nm = 0;
  int rc = modsgpu_match_fginn_dev(ctx, q, t, txy.data(), par.FGINNThreshold, par.contradDist, par.nn, m.data(), &nm);
  if (rc) return rc;
  corresp.TCList.reserve(nm);
  for (int k = 0; k < nm; k++) {
    TentativeCorrespExt tc;
    tc.first = list1[m[k].qi];
    tc.second = list2[m[k].ti];
    tc.secondbad_idx = m[k].tj_bad;
    tc.d1 = m[k].d1; tc.d2 = m[k].d2; tc.ratio = m[k].ratio;
    corresp.TCList.push_back(tc);
  }
  return nm;
}

int MatchDedupDevice(modsgpu_ctx* ctx, const ImageRepresentation& img1, const ImageRepresentation& img2,
                     TentativeCorrespListExt& corresp, const MatchPars& par, int* n_tentatives) {
  corresp.TCList.clear();
  if (n_tentatives) *n_tentatives = 0;
  const AffineRegionVector& list1 = img1.GetAffineRegionVector();
  const AffineRegionVector& list2 = img2.GetAffineRegionVector();
  const int n1 = (int)list1.size(), n2 = (int)list2.size();
  if (n1 == 0 || n2 == 0) return 0;
  const modsgpu_devdesc *q = img1.device_descriptors(), *t = img2.device_descriptors();
  if (!q || !t || modsgpu_devdesc_size(q) != n1 || modsgpu_devdesc_size(t) != n2) return MODSGPU_ESTATE;
  std::vector<double> qxy((size_t)n1 * 2), txy((size_t)n2 * 2);
  for (int i = 0; i < n1; i++) { qxy[2 * i] = list1[i].reproj_kp.x; qxy[2 * i + 1] = list1[i].reproj_kp.y; }
  for (int i = 0; i < n2; i++) { txy[2 * i] = list2[i].reproj_kp.x; txy[2 * i + 1] = list2[i].reproj_kp.y; }
  std::vector<modsgpu_match> m(n1);
  std::vector<int> ord(n1);
  int nm = 0, nu = 0;
  int rc = modsgpu_match_dedup_dev(ctx, q, t, qxy.data(), txy.data(), par.FGINNThreshold, par.contradDist, par.nn,
                                   par.doubleFilteringRadius, m.data(), &nm, ord.data(), &nu);
  if (rc) return rc;
  if (n_tentatives) *n_tentatives = nm;
  corresp.TCList.reserve(nu);
  for (int k = 0; k < nu; k++) {
    const modsgpu_match& mk = m[ord[k]];
    TentativeCorrespExt tc;
    tc.first = list1[mk.qi];
    tc.second = list2[mk.ti];
    tc.secondbad_idx = mk.tj_bad;
    tc.d1 = mk.d1; tc.d2 = mk.d2; tc.ratio = mk.ratio;
    corresp.TCList.push_back(tc);
  }
  return nu;
}

// ---- keyed region store (imagerepresentation.cpp:600-660) ------------------------------------------------------------
void ImageRepresentation::AddRegions(const AffineRegionVector& RegionsToAdd, const std::string& det, const std::string& desc) {
  AffineRegionVector& dst = RegionVectorMap[det][desc];       // appended to an existing list, created otherwise
  dst.insert(dst.end(), RegionsToAdd.begin(), RegionsToAdd.end());
}
AffineRegionVector ImageRepresentation::GetAffineRegionVector(const std::string& desc, const std::string& det) const {
  AffineRegionVector out;
  if (det == det_name && desc == desc_name) out = regions_;   // the lists this object extracted itself
  auto d = RegionVectorMap.find(det);
  if (d != RegionVectorMap.end()) {
    auto e = d->second.find(desc);
    if (e != d->second.end()) out.insert(out.end(), e->second.begin(), e->second.end());
  }
  return out;
}

// ---- CorrespondenceBank (correspondencebank.cpp) ------------------------------------------------------------------------
void CorrespondenceBank::AddCorrespondences(const TentativeCorrespListExt& CorrsToAdd, const std::string& det, const std::string& desc) {
  auto& byDet = CorrespondencesMapMap[desc];                  // :174-199
  auto it = byDet.find(det);
  if (it != byDet.end()) it->second.TCList.insert(it->second.TCList.end(), CorrsToAdd.TCList.begin(), CorrsToAdd.TCList.end());
  else byDet[det] = CorrsToAdd;
}
void CorrespondenceBank::ClearCorrespondences(const std::string& det, const std::string& desc) {
  auto d = CorrespondencesMapMap.find(desc);                  // :200-211
  if (d == CorrespondencesMapMap.end()) return;
  auto e = d->second.find(det);
  if (e != d->second.end()) e->second.TCList.clear();
}
TentativeCorrespListExt CorrespondenceBank::GetCorresponcesVector(const std::string& desc, const std::string& det) const {
  TentativeCorrespListExt corrs;                              // :115-172: descriptors, then detectors, in map (= name) order
  for (const auto& d : CorrespondencesMapMap) {
    if (desc != "All" && d.first != desc) continue;
    for (const auto& e : d.second) {
      if (det != "All" && e.first != det) continue;
      corrs.TCList.insert(corrs.TCList.end(), e.second.TCList.begin(), e.second.TCList.end());
    }
  }
  return corrs;
}
int CorrespondenceBank::GetCorrespondencesNumber(const std::string& desc, const std::string& det) const {
  return (int)GetCorresponcesVector(desc, det).TCList.size();
}
int CorrespondenceBank::MatchImgReps(const ImageRepresentation& imgrep1, const ImageRepresentation& imgrep2, const WhatToMatch& what,
                                     const MatchPars& par, const std::map<std::string, double>& fginn) {
  auto threshold = [&](const std::string& desc) { auto it = fginn.find(desc); return it != fginn.end() ? it->second : 0.0; };
  auto match = [&](const AffineRegionVector& queries, const AffineRegionVector& trains, double thr, TentativeCorrespListExt& out) {
    if (!(thr > 0)) return 0;                                  // currMatchRatio == 0: this descriptor is not matched (:262-277)
    MatchPars mp = par;
    mp.FGINNThreshold = thr;
    return MatchFlannFGINN(ctx_, queries, trains, out, mp);
  };
  // grouped (:246-285): one pooled list per group descriptor
  for (const std::string& curr_desc : what.group_descriptors) {
    ClearCorrespondences("Group", curr_desc);
    AffineRegionVector queries, trains;
    for (const std::string& curr_det : what.group_detectors) {
      const AffineRegionVector t = imgrep2.GetAffineRegionVector(curr_desc, curr_det);
      trains.insert(trains.end(), t.begin(), t.end());
      const AffineRegionVector q = imgrep1.GetAffineRegionVector(curr_desc, curr_det);
      queries.insert(queries.end(), q.begin(), q.end());
    }
    TentativeCorrespListExt current_tents;
    const int rc = match(queries, trains, threshold(curr_desc), current_tents);
    if (rc < 0) return rc;
    AddCorrespondences(current_tents, "Group", curr_desc);
  }
  // separate (:288-340): every separate detector x separate descriptor on its own
  for (const std::string& curr_det : what.separate_detectors)
    for (const std::string& curr_desc : what.separate_descriptors) {
      ClearCorrespondences(curr_det, curr_desc);
      TentativeCorrespListExt current_tents;
      const int rc = match(imgrep1.GetAffineRegionVector(curr_desc, curr_det), imgrep2.GetAffineRegionVector(curr_desc, curr_det),
                           threshold(curr_desc), current_tents);
      if (rc < 0) return rc;
      AddCorrespondences(current_tents, curr_det, curr_desc);
    }
  return 0;
}

// matching.cpp:2615-2679, mode bestFGINN (mods.cpp:283)
int DuplicateFiltering(modsgpu_ctx* ctx, TentativeCorrespListExt& in_corresp, double r) {
  const int T = (int)in_corresp.TCList.size();
  if (r <= 0 || T == 0) return T;
  std::vector<double> xy1(2 * (size_t)T), xy2(2 * (size_t)T), ratio(T);
  for (int i = 0; i < T; i++) {
    const TentativeCorrespExt& c = in_corresp.TCList[i];
    xy1[2 * i] = c.first.reproj_kp.x; xy1[2 * i + 1] = c.first.reproj_kp.y;
    xy2[2 * i] = c.second.reproj_kp.x; xy2[2 * i + 1] = c.second.reproj_kp.y;
    ratio[i] = c.ratio;
  }
  std::vector<int> ord(T);
  int nout = 0;
  int rc = modsgpu_duplicate_filter(ctx, xy1.data(), xy2.data(), ratio.data(), T, r, ord.data(), &nout);
  if (rc) return rc;
  std::vector<TentativeCorrespExt> keep;
  keep.reserve(nout);
  for (int i = 0; i < nout; i++) keep.push_back(in_corresp.TCList[ord[i]]);
  in_corresp.TCList.swap(keep);
  return nout;
}

static bool invert3(const double* A, double* R) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0 || !std::isfinite(det)) { for (int i = 0; i < 9; i++) R[i] = 0; return false; }
  const double id = 1.0 / det;
  R[0] = c0 * id; R[1] = (A[2] * A[7] - A[1] * A[8]) * id; R[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  R[3] = c1 * id; R[4] = (A[0] * A[8] - A[2] * A[6]) * id; R[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  R[6] = c2 * id; R[7] = (A[1] * A[6] - A[0] * A[7]) * id; R[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

// matching.cpp:1014-1043
static int NaiveHCheck(const TentativeCorrespListExt& corresp, const double* H, double error) {
  const double err_sq = error * error;
  double Hinv[9];
  invert3(H, Hinv);
  int corr_numb = 0;
  for (const auto& c : corresp.TCList) {
    const double x1 = c.first.reproj_kp.x, y1 = c.first.reproj_kp.y, x2 = c.second.reproj_kp.x, y2 = c.second.reproj_kp.y;
    double xa = (H[0] * x1 + H[1] * y1 + H[2]) / (H[6] * x1 + H[7] * y1 + H[8]);
    double ya = (H[3] * x1 + H[4] * y1 + H[5]) / (H[6] * x1 + H[7] * y1 + H[8]);
    const double d1 = (x2 - xa) * (x2 - xa) + (y2 - ya) * (y2 - ya);
    xa = (Hinv[0] * x2 + Hinv[1] * y2 + Hinv[2]) / (Hinv[6] * x2 + Hinv[7] * y2 + Hinv[8]);
    ya = (Hinv[3] * x2 + Hinv[4] * y2 + Hinv[5]) / (Hinv[6] * x2 + Hinv[7] * y2 + Hinv[8]);
    const double d2 = (x1 - xa) * (x1 - xa) + (y1 - ya) * (y1 - ya);
    if ((d1 <= err_sq) && (d2 <= err_sq)) corr_numb++;
  }
  return corr_numb;
}

// Htools.c HDsSymMax for one point pair: Hm = Hl transposed (degensac convention: column-major, image 2 -> image 1),
// H1 = inv(Hm); both are the same for every pair of a call and are formed once by the caller
static double HDsSymMax1(const double* Hm, const double* H1, const double* u) {
  const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8];
  const double b = Hm[6] * u[3] + Hm[7] * u[4] + Hm[8];
  double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a, ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
  double xd = u[3] - xa, yd = u[4] - ya;
  const double d1 = xd * xd + yd * yd;
  xa = (Hm[0] * u[3] + Hm[1] * u[4] + Hm[2]) / b; ya = (Hm[3] * u[3] + Hm[4] * u[4] + Hm[5]) / b;
  xd = u[0] - xa; yd = u[1] - ya;
  const double d2 = xd * xd + yd * yd;
  return d1 > d2 ? d1 : d2;
}

// matching.cpp:250-308 with HDS1 = HDsSymMax
static void H_LAF_check(const std::vector<TentativeCorrespExt>& in, const double* Hl, std::vector<TentativeCorrespExt>& res,
                        double affineFerror) {
  const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
  res.clear();
  if (!(affineFerror > 0)) { res = in; return; }
  res.reserve(in.size());
  const double Hm[9] = {Hl[0], Hl[3], Hl[6], Hl[1], Hl[4], Hl[7], Hl[2], Hl[5], Hl[8]};
  double H1[9];
  invert3(Hm, H1);
  for (const auto& c : in) {
    const AffineKeypoint &f = c.first.reproj_kp, &s = c.second.reproj_kp;
    double u[18];
    u[0] = f.x; u[1] = f.y; u[2] = 1.0; u[3] = s.x; u[4] = s.y; u[5] = 1.0;
    u[6] = u[0] + k_sigma * f.a12 * f.s; u[7] = u[1] + k_sigma * f.a22 * f.s; u[8] = 1.0;
    u[9] = u[3] + k_sigma * s.a12 * s.s; u[10] = u[4] + k_sigma * s.a22 * s.s; u[11] = 1.0;
    u[12] = u[0] + k_sigma * f.a11 * f.s; u[13] = u[1] + k_sigma * f.a21 * f.s; u[14] = 1.0;
    u[15] = u[3] + k_sigma * s.a11 * s.s; u[16] = u[4] + k_sigma * s.a21 * s.s; u[17] = 1.0;
    const double sumErr = std::sqrt(HDsSymMax1(Hm, H1, u) + HDsSymMax1(Hm, H1, u + 6) + HDsSymMax1(Hm, H1, u + 12));
    if (!(sumErr > affineFerror)) res.push_back(c);
  }
}

// matching.cpp:637-823, H branch
int LORANSACFiltering(modsgpu_ctx* ctx, TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& ransac_corresp,
                      double* H, const RANSACPars& pars) {
  const int MIN_POINTS = 8;   // matching.hpp:27
  const int tent_size = (int)in_corresp.TCList.size();
  ransac_corresp.TCList.clear();
  int max_samples = pars.max_samples;
  if (tent_size <= 20) max_samples = 1000;
  if (tent_size < MIN_POINTS) return 0;
  std::vector<double> u2((size_t)tent_size * 6);
  for (int i = 0; i < tent_size; i++) {
    const TentativeCorrespExt& c = in_corresp.TCList[i];
    u2[6 * i + 0] = c.first.reproj_kp.x; u2[6 * i + 1] = c.first.reproj_kp.y; u2[6 * i + 2] = 1.;
    u2[6 * i + 3] = c.second.reproj_kp.x; u2[6 * i + 4] = c.second.reproj_kp.y; u2[6 * i + 5] = 1.;
  }
  std::vector<unsigned char> inl2(tent_size);
  double Hloran[9];
  modsgpu_ransac_params rp;
  rp.th = pars.err_threshold * pars.err_threshold;
  rp.conf = pars.confidence;
  rp.max_samples = max_samples;
  rp.do_sym_check = pars.doSymmCheck;
  rp.seed = pars.seed;
  rp.error_type = pars.errorType;
  rp._pad = 0;
  modsgpu_ransac_result rr;
  int rc = pars.useF ? modsgpu_ransac_F(ctx, u2.data(), tent_size, &rp, Hloran, inl2.data(), &rr)
                     : modsgpu_ransac_H(ctx, u2.data(), tent_size, &rp, Hloran, inl2.data(), &rr);
  if (rc) return rc;
  ransac_corresp.TCList.reserve(tent_size);
  for (int i = 0; i < tent_size; i++) {
    in_corresp.TCList[i].isTrue = inl2[i];
    // matching.cpp:738-762: justMarkOutliers keeps every tentative in the list and only flags it
    if (inl2[i] || pars.justMarkOutliers) ransac_corresp.TCList.push_back(in_corresp.TCList[i]);
  }
  return EmpiricalChecks(ransac_corresp, Hloran, pars, H);
}

// The empirical checks at the end of LORANSACFiltering (matching.cpp:764-820) on the list RANSAC returned:
//   H: model = inv(Hloran^T) (all-zero inverse -> empty list), NaiveHCheck (:1014-1043: fewer than MIN_POINTS
//      correspondences within 10 px under H and under H^-1 empty the list), H_LAF_check (:250-308) with HDsSymMax on the
//      three LAF points at 3 * HLAFCoef * err_threshold, fewer than MIN_POINTS survivors empty the list;
//   F: F_LAF_check (:192-249) with FDs at LAFCoef * err_threshold, same MIN_POINTS rule; model = Hloran as is.
int EmpiricalChecks(TentativeCorrespListExt& ransac_corresp, const double* Hloran, const RANSACPars& pars, double* H) {
  const int MIN_POINTS = 8;   // matching.hpp:27
  if (pars.useF) {
    std::vector<TentativeCorrespExt> checked;
    const double affineFerror = pars.LAFCoef * pars.err_threshold;
    const double k_sigma = 2 * 3.0 * std::sqrt(3.0);
    auto fds = [&](const double* u) {
      const double* F = Hloran;
      const double rxc = F[0] * u[3] + F[3] * u[4] + F[6], ryc = F[1] * u[3] + F[4] * u[4] + F[7], rwc = F[2] * u[3] + F[5] * u[4] + F[8];
      const double r = (u[0] * rxc + u[1] * ryc + rwc);
      const double rx = F[0] * u[0] + F[1] * u[1] + F[2], ry = F[3] * u[0] + F[4] * u[1] + F[5];
      return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
    };
    for (const auto& c : ransac_corresp.TCList) {
      if (!(affineFerror > 0)) { checked.push_back(c); continue; }
      const AffineKeypoint &f = c.first.reproj_kp, &s = c.second.reproj_kp;
      double u[18];
      u[0] = f.x; u[1] = f.y; u[2] = 1.0; u[3] = s.x; u[4] = s.y; u[5] = 1.0;
      u[6] = u[0] + k_sigma * f.a12 * f.s; u[7] = u[1] + k_sigma * f.a22 * f.s; u[8] = 1.0;
      u[9] = u[3] + k_sigma * s.a12 * s.s; u[10] = u[4] + k_sigma * s.a22 * s.s; u[11] = 1.0;
      u[12] = u[0] + k_sigma * f.a11 * f.s; u[13] = u[1] + k_sigma * f.a21 * f.s; u[14] = 1.0;
      u[15] = u[3] + k_sigma * s.a11 * s.s; u[16] = u[4] + k_sigma * s.a21 * s.s; u[17] = 1.0;
      const double sumErr = std::sqrt(fds(u)) + std::sqrt(fds(u + 6)) + std::sqrt(fds(u + 12));
      if (!(sumErr > affineFerror)) checked.push_back(c);
    }
    if ((int)checked.size() < MIN_POINTS) checked.clear();
    ransac_corresp.TCList.swap(checked);
    for (int i = 0; i < 9; i++) { ransac_corresp.H[i] = Hloran[i]; H[i] = Hloran[i]; }
    return (int)ransac_corresp.TCList.size();
  }
  // H = inv(Hloran^T)  (matching.cpp:767-784)
  const double Ht[9] = {Hloran[0], Hloran[3], Hloran[6], Hloran[1], Hloran[4], Hloran[7], Hloran[2], Hloran[5], Hloran[8]};
  double Hinv[9];
  invert3(Ht, Hinv);
  bool nonzero = false;
  for (int i = 0; i < 9; i++) nonzero = nonzero || (Hinv[i] != 0.0);
  if (!nonzero) { ransac_corresp.TCList.clear(); return 0; }
  for (int i = 0; i < 9; i++) { ransac_corresp.H[i] = Hinv[i]; H[i] = Hinv[i]; }
  if (NaiveHCheck(ransac_corresp, ransac_corresp.H, 10.0) < MIN_POINTS) ransac_corresp.TCList.clear();
  std::vector<TentativeCorrespExt> checked;
  H_LAF_check(ransac_corresp.TCList, Hloran, checked, 3.0 * pars.HLAFCoef * pars.err_threshold);
  if ((int)checked.size() < MIN_POINTS) checked.clear();
  ransac_corresp.TCList.swap(checked);
  return (int)ransac_corresp.TCList.size();
}

// Htools.c:138-198 pinvJ + HDs for one correspondence (H in the degensac convention)
static double HDs1(const double* H, const double* u) {
  const double x1 = u[0], y1 = u[1], x2 = u[3], y2 = u[4], w2 = u[5];
  double r1 = 0, r2 = 0;
  r1 += H[0] * x2; r1 += H[2] * (-x1 * x2); r1 += H[3] * y2; r1 += H[5] * (-x1 * y2); r1 += H[6] * w2; r1 += H[8] * (-x1 * w2);
  r2 += H[1] * x2; r2 += H[2] * (-y1 * x2); r2 += H[4] * y2; r2 += H[5] * (-y1 * y2); r2 += H[7] * w2; r2 += H[8] * (-y1 * w2);
  const double a = H[0] - H[2] * x1, b = H[3] - H[5] * x1, c = -H[8] - H[2] * x2 - H[5] * y2;
  const double d = H[1] - H[2] * y1, e = H[4] - H[5] * y1;
  const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
  const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
  double pJ[8];
  pJ[0] = -b * de + a * (c2 + e2); pJ[1] = b * c2pd2 - a * de; pJ[2] = c * (c2pd2 + e2); pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e; pJ[5] = -ab * d + e * (a2 + c2); pJ[6] = pJ[3]; pJ[7] = c * (a2 + b2 + c2);
  const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
  double s = 0;
  for (int j = 0; j < 4; j++) { const double t = (pJ[j] / N) * r1 + (pJ[j + 4] / N) * r2; s += t * t; }
  return s;
}

// matching.cpp:917-1013: verification against a KNOWN homography (ver_type GR_TRUTH, mods.cpp:292-303).  NB the
// reference packs u as (second, first) here -- H is expected in the degensac convention of THAT order -- and hands back
// the transposed H.  isExtended keeps every tentative and only sets isTrue.
int HMatrixFiltering(TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& true_corresp, const double* H, int isExtended,
                     const RANSACPars& pars) {
  const size_t T = in_corresp.TCList.size();
  true_corresp.TCList.clear();
  const float th = (float)(pars.err_threshold * pars.err_threshold);       // float in the reference (:968)
  double Hm[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]}, H1[9];
  if (pars.errorType != 0) invert3(Hm, H1);
  int true_size = 0;
  if (isExtended) true_corresp.TCList = in_corresp.TCList;
  for (size_t i = 0; i < T; i++) {
    const TentativeCorrespExt& c = in_corresp.TCList[i];
    const double u[6] = {c.second.reproj_kp.x, c.second.reproj_kp.y, 1., c.first.reproj_kp.x, c.first.reproj_kp.y, 1.};
    double d;
    if (pars.errorType == 0) d = HDs1(H, u);
    else {
      // HDsSymMax / HDsSym (Htools.c:201-284): max / sum of the two transfer errors
      const double a = H1[6] * u[0] + H1[7] * u[1] + H1[8], b = Hm[6] * u[3] + Hm[7] * u[4] + Hm[8];
      double xa = (H1[0] * u[0] + H1[1] * u[1] + H1[2]) / a, ya = (H1[3] * u[0] + H1[4] * u[1] + H1[5]) / a;
      double xd = u[3] - xa, yd = u[4] - ya;
      const double d1 = xd * xd + yd * yd;
      xa = (Hm[0] * u[3] + Hm[1] * u[4] + Hm[2]) / b; ya = (Hm[3] * u[3] + Hm[4] * u[4] + Hm[5]) / b;
      xd = u[0] - xa; yd = u[1] - ya;
      const double d2 = xd * xd + yd * yd;
      d = pars.errorType == 1 ? (d1 > d2 ? d1 : d2) : d1 + d2;
    }
    const int inl = d <= th ? 1 : 0;
    true_size += inl;
    if (isExtended) true_corresp.TCList[i].isTrue = inl;
    else if (inl) true_corresp.TCList.push_back(c);
  }
  const double Ht[9] = {H[0], H[3], H[6], H[1], H[4], H[7], H[2], H[5], H[8]};
  for (int i = 0; i < 9; i++) true_corresp.H[i] = Ht[i];
  return true_size;
}

// matching.cpp:574-633 on the device (modsgpu_match_hamming): binary descriptors, 2-NN by Hamming distance
int MatchFLANNDistance(modsgpu_ctx* ctx, const AffineRegionVector& list1, const AffineRegionVector& list2,
                       TentativeCorrespListExt& corresp, double matchDistanceThreshold) {
  corresp.TCList.clear();
  const int n1 = (int)list1.size(), n2 = (int)list2.size();
  if (n1 == 0 || n2 == 0) return 0;
  const int dim = (int)list1[0].desc.size();
  std::vector<float> q((size_t)n1 * dim), t((size_t)n2 * dim);
  for (int i = 0; i < n1; i++) memcpy(&q[(size_t)i * dim], list1[i].desc.data(), dim * sizeof(float));
  for (int i = 0; i < n2; i++) memcpy(&t[(size_t)i * dim], list2[i].desc.data(), dim * sizeof(float));
  std::vector<modsgpu_match> m(n1);
  int nm = 0;
  int rc = modsgpu_match_hamming(ctx, q.data(), n1, t.data(), n2, dim, matchDistanceThreshold, m.data(), &nm);
  if (rc) return rc;
  corresp.TCList.reserve(nm);
  for (int k = 0; k < nm; k++) {
    TentativeCorrespExt tc;
    tc.first = list1[m[k].qi];
    tc.second = list2[m[k].ti];
    tc.d1 = m[k].d1; tc.d2 = m[k].d2; tc.ratio = m[k].ratio;
    corresp.TCList.push_back(tc);
  }
  return nm;
}

// mods.cpp:202-356, HessianAffine steps only (MSER and the other detectors stay on the reference's CPU path)
int MODSPair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const std::vector<IterationStep>& steps,
             int minMatches, const RANSACPars& rp, MODSResult& res, TentativeCorrespListExt& verified) {
  return MODSPair(ctx, img1, img2, steps, minMatches, DetectPars(), MatchPars(), rp, res, verified);
}
int MODSPair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const std::vector<IterationStep>& steps,
             int minMatches, const DetectPars& dp, const MatchPars& mp0, const RANSACPars& rp, MODSResult& res,
             TentativeCorrespListExt& verified) {
  res = MODSResult();
  verified.TCList.clear();
  ImageRepresentation r1(ctx, img1, false), r2(ctx, img2, false);
  std::vector<ViewSynthParameters> hist;   // SetVSPars history: a view is synthesised once per run
  int curr_matches = 0;
  for (size_t step = 0; step < steps.size() && curr_matches < minMatches; step++) {
    const IterationStep& st = steps[step];
    std::vector<ViewSynthParameters> views;
    SetVSPars(st.ScaleSet, st.TiltSet, st.Phi, views, hist, st.initSigma, st.doBlur);
    int rc = r1.AddViews(views, dp);
    if (rc < 0) return rc;
    rc = r2.AddViews(views, dp);
    if (rc < 0) return rc;
    res.steps_done = (int)step + 1;
    res.views[0] = r1.n_views; res.views[1] = r2.n_views;
    res.regions[0] = (int)r1.GetAffineRegionVector().size(); res.regions[1] = (int)r2.GetAffineRegionVector().size();
    MatchPars mp = mp0;
    mp.FGINNThreshold = st.FGINNThreshold;
    TentativeCorrespListExt tent;
    int nt = MatchFlannFGINN(ctx, r1.GetAffineRegionVector(), r2.GetAffineRegionVector(), tent, mp);
    if (nt < 0) return nt;
    res.tentatives = nt;
    int nu = DuplicateFiltering(ctx, tent, mp.doubleFilteringRadius);
    if (nu < 0) return nu;
    res.unique_tentatives = nu;
    int ni = LORANSACFiltering(ctx, tent, verified, res.model, rp);
    if (ni < 0) return ni;
    res.inliers = ni;
    curr_matches = ni;
  }
  return 0;
}

}  // namespace modsb200

// HMatrixFiltering (matching.cpp:917-1013) on caller-supplied centres: xy1 / xy2 = reproj_kp centres of the n tentatives in
// image 1 / image 2, H in the convention the reference's caller passes (ground-truth file, degensac layout for u = (image 2,
// image 1)).  keep[i] = within err_threshold under the selected error; H_out = H transposed (what true_corresp.H receives).
extern "C" int modsgpu_hmatrix_filter(const double* xy1, const double* xy2, int n, const double* H, int error_type, double err_threshold,
                                      unsigned char* keep, double* H_out, int* n_out) {
  using namespace modsb200;
  if (n < 0 || (n > 0 && (!xy1 || !xy2 || !keep)) || !H || !n_out || error_type < 0 || error_type > 2) return MODSGPU_EINVAL;
  TentativeCorrespListExt in, out;
  in.TCList.resize(n);
  for (int i = 0; i < n; i++) {
    in.TCList[i].first.reproj_kp.x = xy1[2 * i]; in.TCList[i].first.reproj_kp.y = xy1[2 * i + 1];
    in.TCList[i].second.reproj_kp.x = xy2[2 * i]; in.TCList[i].second.reproj_kp.y = xy2[2 * i + 1];
  }
  RANSACPars pars;
  pars.errorType = error_type;
  pars.err_threshold = err_threshold;
  *n_out = HMatrixFiltering(in, out, H, 1, pars);
  for (int i = 0; i < n; i++) keep[i] = (unsigned char)out.TCList[i].isTrue;
  if (H_out) for (int i = 0; i < 9; i++) H_out[i] = out.H[i];
  return 0;
}

// host-only seam of the checks above (no device work, callable without a GPU): kp1 / kp2 = reproj_kp of the n
// correspondences RANSAC kept; model = the degensac-convention H (column-major, image 2 -> 1) or F; keep[i] = survives.
extern "C" int modsgpu_empirical_checks(const modsgpu_region* kp1, const modsgpu_region* kp2, int n, const double* model, int use_F,
                                        double err_threshold, double laf_coef, unsigned char* keep, double* model_out, int* n_out) {
  using namespace modsb200;
  if (n < 0 || (n > 0 && (!kp1 || !kp2 || !keep)) || !model || !model_out || !n_out) return MODSGPU_EINVAL;
  TentativeCorrespListExt list;
  list.TCList.resize(n);
  auto put = [](AffineKeypoint& k, const modsgpu_region& r) { k.x = r.x; k.y = r.y; k.s = r.s; k.a11 = r.a11; k.a12 = r.a12; k.a21 = r.a21; k.a22 = r.a22; };
  for (int i = 0; i < n; i++) {
    put(list.TCList[i].first.reproj_kp, kp1[i]);
    put(list.TCList[i].second.reproj_kp, kp2[i]);
    list.TCList[i].first.id = i;      // carries the input index through the filters
  }
  RANSACPars pars;
  pars.useF = use_F;
  pars.err_threshold = err_threshold;
  if (use_F) pars.LAFCoef = laf_coef; else pars.HLAFCoef = laf_coef;
  for (int i = 0; i < 9; i++) model_out[i] = 0;
  const int m = EmpiricalChecks(list, model, pars, model_out);
  memset(keep, 0, (size_t)n);
  for (const auto& c : list.TCList) keep[c.first.id] = 1;
  *n_out = m;
  return 0;
}

// CorrespondenceBank::MatchImgReps over caller-supplied region lists (test / integration seam).  lists: which image (1 or 2),
// detector and descriptor name each list is filed under.  Name sets and thresholds are comma-separated strings
// ("HessianAffine,MSER", "ZMQ=0.8,RootSIFT=0.85").  out: 7 doubles per tentative (x1 y1 x2 y2 d1 d2 ratio) in the order of
// GetCorresponcesVector("All", "All").
extern "C" int modsgpu_match_imgreps(modsgpu_ctx* ctx, const modsgpu_region_list* lists, int n_lists, const char* group_detectors,
                                     const char* group_descriptors, const char* separate_detectors, const char* separate_descriptors,
                                     const char* fginn_thresholds, double* out, int capacity, int* n_out) {
  using namespace modsb200;
  if (!ctx || n_lists < 0 || (n_lists > 0 && !lists) || !n_out) return MODSGPU_EINVAL;
  auto split = [](const char* s) {
    std::vector<std::string> v;
    std::string cur;
    for (const char* p = s ? s : ""; ; p++) {
      if (*p == ',' || *p == 0) { if (!cur.empty()) v.push_back(cur); cur.clear(); if (!*p) break; }
      else if (*p != ' ') cur.push_back(*p);
    }
    return v;
  };
  ImageRepresentation rep1(ctx, nullptr, false), rep2(ctx, nullptr, false);
  rep1.det_name = rep2.det_name = "";          // nothing extracted by these objects: every list comes from AddRegions
  for (int l = 0; l < n_lists; l++) {
    const modsgpu_region_list& L = lists[l];
    if ((L.image != 1 && L.image != 2) || !L.det || !L.desc || L.n < 0 || (L.n > 0 && !L.f)) return MODSGPU_EINVAL;
    AffineRegionVector v(L.n);
    auto blk = std::make_shared<std::vector<float>>((size_t)L.n * 128);
    for (int i = 0; i < L.n; i++) {
      AffineKeypoint& k = v[i].reproj_kp;
      k.x = L.f[i].x; k.y = L.f[i].y; k.s = L.f[i].s; k.a11 = L.f[i].a11; k.a12 = L.f[i].a12; k.a21 = L.f[i].a21; k.a22 = L.f[i].a22;
      k.response = L.f[i].response; k.octave_number = L.f[i].octave; k.sub_type = L.f[i].type;
      v[i].det_kp = k; v[i].id = i;
      memcpy(blk->data() + (size_t)i * 128, L.f[i].desc, 512);
    }
    std::shared_ptr<const std::vector<float>> cblk = blk;
    for (int i = 0; i < L.n; i++) v[i].desc.view(cblk, (size_t)i * 128, 128);
    (L.image == 1 ? rep1 : rep2).AddRegions(v, L.det, L.desc);
  }
  WhatToMatch what;
  what.group_detectors = split(group_detectors); what.group_descriptors = split(group_descriptors);
  what.separate_detectors = split(separate_detectors); what.separate_descriptors = split(separate_descriptors);
  std::map<std::string, double> fginn;
  for (const std::string& kv : split(fginn_thresholds)) {
    const size_t eq = kv.find('=');
    if (eq == std::string::npos) return MODSGPU_EINVAL;
    fginn[kv.substr(0, eq)] = atof(kv.c_str() + eq + 1);
  }
  CorrespondenceBank bank(ctx);
  MatchPars mp;
  const int rc = bank.MatchImgReps(rep1, rep2, what, mp, fginn);
  if (rc) return rc;
  const TentativeCorrespListExt all = bank.GetCorresponcesVector();
  *n_out = (int)all.TCList.size();
  for (int i = 0; i < *n_out && i < capacity && out; i++) {
    const TentativeCorrespExt& c = all.TCList[i];
    double* o = out + 7 * (size_t)i;
    o[0] = c.first.reproj_kp.x; o[1] = c.first.reproj_kp.y; o[2] = c.second.reproj_kp.x; o[3] = c.second.reproj_kp.y;
    o[4] = c.d1; o[5] = c.d2; o[6] = c.ratio;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// C entry: the whole pair pipeline (what mods.cpp:202-356 does for one iteration of the deep config)
// ------------------------------------------------------------------------------------------------------
static void unpack_params(const modsgpu_pipeline_params* p, modsb200::DetectPars& dp, modsb200::MatchPars& mp, modsb200::RANSACPars& rp) {
  dp.pyr = p->pyr; dp.mrSize = p->mrSize; dp.patchSize = p->patchSize;
  mp.FGINNThreshold = p->fginn_threshold; mp.contradDist = p->contrad_dist; mp.nn = p->nn; mp.doubleFilteringRadius = p->dup_filter_radius;
  rp.err_threshold = p->err_threshold; rp.confidence = p->confidence; rp.max_samples = p->max_samples; rp.doSymmCheck = p->do_symm_check;
  rp.HLAFCoef = p->HLAFCoef; rp.LAFCoef = p->LAFCoef; rp.errorType = p->error_type; rp.justMarkOutliers = p->just_mark_outliers;
  rp.useF = p->use_F; rp.seed = p->seed;
}
extern "C" void modsgpu_default_pipeline_params(modsgpu_pipeline_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  const modsb200::DetectPars dp;
  const modsb200::MatchPars mp;
  const modsb200::RANSACPars rp;
  p->pyr = dp.pyr; p->mrSize = dp.mrSize; p->patchSize = dp.patchSize;
  p->fginn_threshold = mp.FGINNThreshold; p->contrad_dist = mp.contradDist; p->dup_filter_radius = mp.doubleFilteringRadius; p->nn = mp.nn;
  p->err_threshold = rp.err_threshold; p->confidence = rp.confidence; p->HLAFCoef = rp.HLAFCoef; p->LAFCoef = rp.LAFCoef;
  p->max_samples = rp.max_samples; p->do_symm_check = rp.doSymmCheck; p->error_type = rp.errorType;
  p->just_mark_outliers = rp.justMarkOutliers; p->use_F = rp.useF; p->seed = rp.seed;
}
extern "C" int modsgpu_pair_pipeline_images(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2,
                                            unsigned long long seed, modsgpu_pair_result* res,
                                            double* inlier_xy /* capacity*4 or NULL */, int capacity) {
  modsgpu_pipeline_params p;
  modsgpu_default_pipeline_params(&p);
  p.seed = seed;
  return modsgpu_pair_pipeline_images_ex(ctx, img1, img2, &p, res, inlier_xy, capacity);
}
extern "C" int modsgpu_pair_pipeline_images_ex(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2,
                                               const modsgpu_pipeline_params* pp, modsgpu_pair_result* res,
                                               double* inlier_xy /* capacity*4 or NULL */, int capacity) {
  using namespace modsb200;
  if (!ctx || !img1 || !img2 || !res || !pp) return MODSGPU_EINVAL;
  if (pp->patchSize != 32 || !(pp->err_threshold > 0) || pp->nn < 2 || pp->nn > 50) return MODSGPU_EINVAL;
  memset(res, 0, sizeof(*res));
  DetectPars dp;
  MatchPars mp;
  RANSACPars rp;
  unpack_params(pp, dp, mp, rp);
  // the pair-level call returns correspondences, never descriptors: they stay on the device between HardNet++ and the
  // matcher (MODSGPU_HOST_DESC=1, read per call, keeps the host round trip of the seam route -- the tests compare the two)
  {
    const char* e = getenv("MODSGPU_HOST_DESC");
    const char* seam = getenv("MODSGPU_SEAM_CHAIN");
    dp.desc_on_device = !(e && atoi(e) != 0) && !(seam && atoi(seam) != 0);
  }
  const bool hp = host_profile();
  double c[6] = {0, 0, 0, 0, 0, 0}, w[6] = {0, 0, 0, 0, 0, 0};
  auto mark = [&](int i) { if (hp) { c[i] = cpu_ms(); w[i] = now_ms(); } };
  mark(0);
  // mods.cpp:234-251 extracts the two images as concurrent OpenMP tasks; with pair overlap on, image 2 runs on the
  // sibling context from a helper thread
  modsgpu_ctx* sib = nullptr;
  if (modsgpu_get_pair_overlap(ctx) && !hp) {
    int rc = modsgpu_ctx_sibling(ctx, &sib);
    if (rc) return rc;
  }
  ImageRepresentation r1(ctx, img1, false), r2(sib ? sib : ctx, img2, false);
  int n1, n2;
  if (sib) {
    n2 = 0;
    std::thread t2([&] { n2 = r2.SynthDetectDescribeKeypoints(dp); });
    n1 = r1.SynthDetectDescribeKeypoints(dp);
    t2.join();
    modsgpu_ctx_sibling_join(ctx);
    if (n1 < 0) return n1;
    if (n2 < 0) return n2;
  } else {
    n1 = r1.SynthDetectDescribeKeypoints(dp);
    if (n1 < 0) return n1;
    mark(1);
    n2 = r2.SynthDetectDescribeKeypoints(dp);
    if (n2 < 0) return n2;
  }
  mark(2);
  res->keypoints[0] = r1.n_keypoints; res->keypoints[1] = r2.n_keypoints;
  res->regions[0] = r1.n_affine; res->regions[1] = r2.n_affine;
  res->descriptors[0] = n1; res->descriptors[1] = n2;
  TentativeCorrespListExt tent, verified;
  int nt, nu;
  const char* unfused_env = getenv("MODSGPU_UNFUSED_TAIL");      // read per call: the tests compare the two routes
  const bool fused_tail = dp.desc_on_device && !(unfused_env && atoi(unfused_env) != 0) &&
                          (int)r1.GetAffineRegionVector().size() <= 16384;
  if (fused_tail) {
    // matcher + duplicate filter in one device call: the tentative list stays on the device between them
    nu = MatchDedupDevice(ctx, r1, r2, tent, mp, &nt);
    if (nu < 0) return nu;
    mark(3);
    res->tentatives = nt;
  } else {
    nt = dp.desc_on_device ? MatchFlannFGINNDevice(ctx, r1, r2, tent, mp)
                           : MatchFlannFGINN(ctx, r1.GetAffineRegionVector(), r2.GetAffineRegionVector(), tent, mp);
    if (nt < 0) return nt;
    mark(3);
    res->tentatives = nt;
    nu = DuplicateFiltering(ctx, tent, mp.doubleFilteringRadius);
    if (nu < 0) return nu;
  }
  mark(4);
  res->unique_tentatives = nu;
  int ni = LORANSACFiltering(ctx, tent, verified, res->H, rp);
  if (ni < 0) return ni;
  mark(5);
  res->inliers = ni;
  if (hp)
    fprintf(stderr, "[modsgpu host] cpu ms (wall ms): image1 %.2f (%.2f) image2 %.2f (%.2f) match %.2f (%.2f) dup %.2f (%.2f) ransac %.2f (%.2f)\n",
            c[1] - c[0], w[1] - w[0], c[2] - c[1], w[2] - w[1], c[3] - c[2], w[3] - w[2], c[4] - c[3], w[4] - w[3],
            c[5] - c[4], w[5] - w[4]);
  for (int i = 0; i < ni && i < capacity && inlier_xy; i++) {
    const TentativeCorrespExt& c = verified.TCList[i];
    inlier_xy[4 * i + 0] = c.first.reproj_kp.x; inlier_xy[4 * i + 1] = c.first.reproj_kp.y;
    inlier_xy[4 * i + 2] = c.second.reproj_kp.x; inlier_xy[4 * i + 3] = c.second.reproj_kp.y;
  }
  return 0;
}

// config 1 (config_affori_classic.ini + iters_HessianSIFT.ini): the same pair loop with the classic per-image stages
extern "C" int modsgpu_pair_pipeline_classic_images(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2,
                                                    unsigned long long seed, modsgpu_pair_result* res, double* inlier_xy,
                                                    int capacity) {
  using namespace modsb200;
  if (!ctx || !img1 || !img2 || !res) return MODSGPU_EINVAL;
  memset(res, 0, sizeof(*res));
  DetectPars dp;
  MatchPars mp;
  RANSACPars rp;
  rp.seed = seed;
  ImageRepresentation r1(ctx, img1, false), r2(ctx, img2, false);
  int n1 = r1.SynthDetectDescribeKeypointsClassic(dp);
  if (n1 < 0) return n1;
  int n2 = r2.SynthDetectDescribeKeypointsClassic(dp);
  if (n2 < 0) return n2;
  res->keypoints[0] = r1.n_keypoints; res->keypoints[1] = r2.n_keypoints;
  res->regions[0] = r1.n_affine; res->regions[1] = r2.n_affine;
  res->descriptors[0] = n1; res->descriptors[1] = n2;
  TentativeCorrespListExt tent, verified;
  int nt = MatchFlannFGINN(ctx, r1.GetAffineRegionVector(), r2.GetAffineRegionVector(), tent, mp);
  if (nt < 0) return nt;
  res->tentatives = nt;
  int nu = DuplicateFiltering(ctx, tent, mp.doubleFilteringRadius);
  if (nu < 0) return nu;
  res->unique_tentatives = nu;
  int ni = LORANSACFiltering(ctx, tent, verified, res->H, rp);
  if (ni < 0) return ni;
  res->inliers = ni;
  for (int i = 0; i < ni && i < capacity && inlier_xy; i++) {
    const TentativeCorrespExt& c = verified.TCList[i];
    inlier_xy[4 * i + 0] = c.first.reproj_kp.x; inlier_xy[4 * i + 1] = c.first.reproj_kp.y;
    inlier_xy[4 * i + 2] = c.second.reproj_kp.x; inlier_xy[4 * i + 3] = c.second.reproj_kp.y;
  }
  return 0;
}

extern "C" int modsgpu_pair_pipeline(modsgpu_ctx* ctx, const uint8_t* bgr1, const uint8_t* bgr2, int w, int h,
                                     unsigned long long seed, modsgpu_pair_result* res, double* inlier_xy, int capacity) {
  if (!ctx || !bgr1 || !bgr2 || !res) return MODSGPU_EINVAL;
  modsgpu_image *i1 = nullptr, *i2 = nullptr;
  int rc = modsgpu_image_from_bgr8(ctx, bgr1, w, h, &i1);
  if (rc) return rc;
  rc = modsgpu_image_from_bgr8(ctx, bgr2, w, h, &i2);
  if (rc) { modsgpu_image_free(ctx, i1); return rc; }
  rc = modsgpu_pair_pipeline_images(ctx, i1, i2, seed, res, inlier_xy, capacity);
  modsgpu_image_free(ctx, i1);
  modsgpu_image_free(ctx, i2);
  return rc;
}


// ------------------------------------------------------------------------------------------------------
// C entries: one image -> described regions (extract_features_batch.cpp:128-139) and the OxAff writer
// ------------------------------------------------------------------------------------------------------
static int export_features(const modsb200::AffineRegionVector& v, int nd, modsgpu_feature** out, int* n);

extern "C" int modsgpu_view_schedule(const double* scale_set, int n_scales, const double* tilt_set, int n_tilts, double phi_base,
                                     double InitSigma, int doBlur, modsgpu_view* out, int cap) {
  using namespace modsb200;
  if (n_scales < 0 || n_tilts < 0 || (n_scales > 0 && !scale_set) || (n_tilts > 0 && !tilt_set) || (cap > 0 && !out)) return MODSGPU_EINVAL;
  std::vector<double> sc(scale_set, scale_set + n_scales), ti(tilt_set, tilt_set + n_tilts);
  std::vector<ViewSynthParameters> par, prev;
  const int nv = SetVSPars(sc, ti, phi_base, par, prev, InitSigma, doBlur);
  for (int i = 0; i < nv && i < cap; i++) {
    out[i].tilt = par[i].tilt; out[i].phi = par[i].phi; out[i].zoom = par[i].zoom; out[i].InitSigma = par[i].InitSigma;
    out[i].doBlur = par[i].doBlur; out[i]._pad = 0;
  }
  return nv;
}

extern "C" int modsgpu_extract_features_views(modsgpu_ctx* ctx, modsgpu_image* img, const modsgpu_view* views, int n_views,
                                              modsgpu_feature** out, int* n) {
  using namespace modsb200;
  if (!ctx || !img || !out || !n || n_views < 0 || (n_views > 0 && !views)) return MODSGPU_EINVAL;
  *out = nullptr; *n = 0;
  std::vector<ViewSynthParameters> vs((size_t)n_views);
  for (int i = 0; i < n_views; i++) {
    vs[i].tilt = views[i].tilt; vs[i].phi = views[i].phi; vs[i].zoom = views[i].zoom; vs[i].InitSigma = views[i].InitSigma;
    vs[i].doBlur = views[i].doBlur;
  }
  DetectPars dp;
  ImageRepresentation rep(ctx, img, false);
  const int nd = rep.SynthDetectDescribeKeypoints(vs, dp);
  if (nd < 0) return nd;
  return export_features(rep.GetAffineRegionVector(), nd, out, n);
}

extern "C" int modsgpu_extract_features(modsgpu_ctx* ctx, modsgpu_image* img, modsgpu_feature** out, int* n) {
  using namespace modsb200;
  if (!ctx || !img || !out || !n) return MODSGPU_EINVAL;
  *out = nullptr; *n = 0;
  DetectPars dp;
  ImageRepresentation rep(ctx, img, false);
  const int nd = rep.SynthDetectDescribeKeypoints(dp);
  if (nd < 0) return nd;
  return export_features(rep.GetAffineRegionVector(), nd, out, n);
}

static int export_features(const modsb200::AffineRegionVector& v, int nd, modsgpu_feature** out, int* n) {
  using namespace modsb200;
  modsgpu_feature* f = (modsgpu_feature*)malloc(sizeof(modsgpu_feature) * (size_t)(nd > 0 ? nd : 1));
  if (!f) return MODSGPU_EINVAL;
  for (int i = 0; i < nd; i++) {
    const AffineKeypoint& k = v[i].reproj_kp;
    f[i].x = k.x; f[i].y = k.y; f[i].s = k.s; f[i].a11 = k.a11; f[i].a12 = k.a12; f[i].a21 = k.a21; f[i].a22 = k.a22;
    f[i].response = k.response; f[i].octave = k.octave_number; f[i].type = k.sub_type;
    f[i].view = v[i].img_reproj_id; f[i]._pad = 0;
    for (int d = 0; d < 128; d++) f[i].desc[d] = d < (int)v[i].desc.size() ? v[i].desc[d] : 0.f;
  }
  *out = f; *n = nd;
  return 0;
}

extern "C" int modsgpu_write_oxaff(const char* path, const modsgpu_feature* f, int n) {
  using namespace modsb200;
  if (!path || n < 0 || (n > 0 && !f)) return MODSGPU_EINVAL;
  AffineRegionVector v((size_t)n);
  auto blk = std::make_shared<std::vector<float>>((size_t)n * 128);
  for (int i = 0; i < n; i++) memcpy(blk->data() + (size_t)i * 128, f[i].desc, 512);
  const std::shared_ptr<const std::vector<float>> cblk = blk;
  for (int i = 0; i < n; i++) {
    AffineKeypoint& k = v[i].reproj_kp;
    k.x = f[i].x; k.y = f[i].y; k.s = f[i].s; k.a11 = f[i].a11; k.a12 = f[i].a12; k.a21 = f[i].a21; k.a22 = f[i].a22;
    v[i].desc.view(cblk, (size_t)i * 128, 128);
  }
  return SaveRegionsMichal(v, path) ? MODSGPU_EIO : 0;
}


// imagerepresentation.cpp:1219-1255 SaveRegions (text) with saveAR / saveKPBench (:109-111, :196-203)
extern "C" int modsgpu_write_regions_text(const char* path, const modsgpu_feature* f, int n) {
  if (!path || n < 0 || (n > 0 && !f)) return MODSGPU_EINVAL;
  std::ofstream kpfile(path);
  if (!kpfile.is_open()) return MODSGPU_EIO;
  kpfile << 1 << std::endl;
  kpfile << "HessianAffine" << " " << 1 << std::endl;
  kpfile << "ZMQ" << " " << n << std::endl;
  if (n > 0) kpfile << 128 << std::endl;
  for (int i = 0; i < n; i++) {
    kpfile << f[i].x << " " << f[i].y << " " << f[i].s << " " << f[i].a11 << " " << f[i].a12 << " " << f[i].a21 << " " << f[i].a22;
    kpfile << " " << 128 << " ";
    for (int d = 0; d < 128; d++) kpfile << f[i].desc[d] << " ";
    kpfile << std::endl;
  }
  return kpfile.good() ? 0 : MODSGPU_EIO;
}

// imagerepresentation.cpp:1257-1316 SaveRegionsNPZ
extern "C" int modsgpu_write_regions_npz(const char* path, const modsgpu_feature* f, int n) {
  if (!path || n < 0 || (n > 0 && !f)) return MODSGPU_EINVAL;
  const size_t N = (size_t)n;
  std::vector<double> xy(2 * N), scales(N), responses(N), A(4 * N);
  std::vector<unsigned char> descs(128 * N);
  for (size_t i = 0; i < N; i++) {
    xy[2 * i] = f[i].x; xy[2 * i + 1] = f[i].y;
    scales[i] = f[i].s;
    A[4 * i] = f[i].a11; A[4 * i + 1] = f[i].a12; A[4 * i + 2] = f[i].a21; A[4 * i + 3] = f[i].a22;
    responses[i] = f[i].response;
    for (int d = 0; d < 128; d++) descs[i * 128 + d] = (unsigned char)f[i].desc[d];
  }
  NpzWriter w(path);
  if (!w.ok()) return MODSGPU_EIO;
  w.add("xy", "<f8", {N, 2}, xy.data(), xy.size() * 8);
  w.add("scales", "<f8", {N, 1}, scales.data(), scales.size() * 8);
  w.add("responses", "<f8", {N, 1}, responses.data(), responses.size() * 8);
  w.add("A", "<f8", {N, 4}, A.data(), A.size() * 8);
  w.add("descs", "|u1", {N, 128}, descs.data(), descs.size());
  return w.close() ? 0 : MODSGPU_EIO;
}


// imagerepresentation.cpp:1355-1507 PreLoadRegionsNPZ: xy [N,2], scales [N], responses [N], descs [N,D] and either
// A [N,4] (affine shape), angles [N] (degrees -> rotation) or neither (upright circular regions).  type = DET_READ.
static const int DET_READ = 4;    // structures.hpp:16-20 detector_type (ReadAffs)
extern "C" int modsgpu_read_regions_npz(const char* path, modsgpu_feature** out, int* n) {
  if (!path || !out || !n) return MODSGPU_EINVAL;
  *out = nullptr; *n = 0;
  std::map<std::string, NpzRaw> z;
  std::string err;
  if (!npz_load_raw(path, z, err)) return MODSGPU_EIO;
  for (const char* k : {"xy", "scales", "responses", "descs"})
    if (!z.count(k) || !z[k].numeric()) return MODSGPU_EIO;
  const NpzRaw &xy = z["xy"], &sc = z["scales"], &rs = z["responses"], &ds = z["descs"];
  if (xy.shape.size() != 2 || xy.shape[1] != 2 || ds.shape.size() != 2) return MODSGPU_EIO;
  const size_t N = (size_t)xy.shape[0];
  const int D = ds.shape[1];
  if (sc.count() != N || rs.count() != N || (size_t)ds.shape[0] != N || D < 1 || D > 128) return MODSGPU_EIO;
  const NpzRaw* A = z.count("A") ? &z["A"] : nullptr;
  const NpzRaw* ang = !A && z.count("angles") ? &z["angles"] : nullptr;
  if (A && (!A->numeric() || A->count() != 4 * N)) return MODSGPU_EIO;
  if (ang && (!ang->numeric() || ang->count() != N)) return MODSGPU_EIO;
  modsgpu_feature* f = (modsgpu_feature*)calloc(N > 0 ? N : 1, sizeof(modsgpu_feature));
  if (!f) return MODSGPU_EINVAL;
  for (size_t i = 0; i < N; i++) {
    f[i].x = xy.at(2 * i); f[i].y = xy.at(2 * i + 1);
    f[i].s = sc.at(i);
    if (A) { f[i].a11 = A->at(4 * i); f[i].a12 = A->at(4 * i + 1); f[i].a21 = A->at(4 * i + 2); f[i].a22 = A->at(4 * i + 3); }
    else {
      const double angle = ang ? ang->at(i) * M_PI / 180.0 : 0.0;
      f[i].a11 = cos(angle); f[i].a12 = sin(angle); f[i].a21 = -sin(angle); f[i].a22 = cos(angle);
    }
    f[i].response = rs.at(i);
    f[i].type = DET_READ;
    for (int d = 0; d < D; d++) f[i].desc[d] = (float)(unsigned char)ds.at(i * D + d);   // descs_ is read as uchar (:1381)
  }
  *out = f; *n = (int)N;
  return 0;
}

// imagerepresentation.cpp:1317-1354 LoadRegions (text) with loadAR / loadKP (:237-253): per region
//   id img_id img_reproj_id parent_id | det_kp: x y a11 a12 a21 a22 pyramid_scale octave s sub_type | reproj_kp: same |
//   desc_size d0 .. ;  every detector/descriptor list of the file is appended in file order.
extern "C" int modsgpu_read_regions_text(const char* path, modsgpu_feature** out, int* n) {
  if (!path || !out || !n) return MODSGPU_EINVAL;
  *out = nullptr; *n = 0;
  std::ifstream kpfile(path);
  if (!kpfile.is_open()) return MODSGPU_EIO;
  int numberOfDetectors = 0;
  kpfile >> numberOfDetectors;
  if (!kpfile || numberOfDetectors < 0) return MODSGPU_EIO;
  std::vector<modsgpu_feature> v;
  for (int det = 0; det < numberOfDetectors; det++) {
    std::string det_name, desc_name;
    int num_of_descs = 0;
    kpfile >> det_name >> num_of_descs;
    for (int desc = 0; desc < num_of_descs; desc++) {
      int num_of_kp = 0, desc_size = 0;
      kpfile >> desc_name >> num_of_kp;
      if (num_of_kp > 0) kpfile >> desc_size;          // SaveRegions writes the size line only for non-empty lists (:1236)
      if (!kpfile || num_of_kp < 0) return MODSGPU_EIO;
      for (int kp = 0; kp < num_of_kp; kp++) {
        modsgpu_feature f;
        memset(&f, 0, sizeof(f));
        int id, img_id, img_reproj_id, parent_id, size1 = 0;
        double d[10], r[10];
        kpfile >> id >> img_id >> img_reproj_id >> parent_id;
        for (int k = 0; k < 10; k++) kpfile >> d[k];
        for (int k = 0; k < 10; k++) kpfile >> r[k];
        kpfile >> size1;
        if (!kpfile || size1 < 0 || size1 > 128) return MODSGPU_EIO;
        for (int k = 0; k < size1; k++) kpfile >> f.desc[k];
        if (!kpfile) return MODSGPU_EIO;
        f.x = r[0]; f.y = r[1]; f.a11 = r[2]; f.a12 = r[3]; f.a21 = r[4]; f.a22 = r[5];
        f.octave = (int)r[7]; f.s = r[8]; f.type = (int)r[9];
        f.view = img_reproj_id;
        v.push_back(f);
      }
    }
  }
  modsgpu_feature* f = (modsgpu_feature*)malloc(sizeof(modsgpu_feature) * (v.size() ? v.size() : 1));
  if (!f) return MODSGPU_EINVAL;
  if (!v.empty()) memcpy(f, v.data(), sizeof(modsgpu_feature) * v.size());
  *out = f; *n = (int)v.size();
  return 0;
}

// mods.cpp:216-229 (`read_pre_extracted`) + :262-356: two pre-extracted region lists -> FGINN tentatives -> duplicate
// filter -> LO-RANSAC.  The same three calls MODSPair makes per step.
static void features_to_lists(const modsgpu_feature* f1, int n1, const modsgpu_feature* f2, int n2, int desc_dim,
                              modsb200::AffineRegionVector (&l)[2]) {
  using namespace modsb200;
  const modsgpu_feature* src[2] = {f1, f2};
  const int cnt[2] = {n1, n2};
  for (int k = 0; k < 2; k++) {
    l[k].resize((size_t)cnt[k]);
    // one descriptor block per list, the regions hold views into it (what DescribeView produces)
    auto blk = std::make_shared<std::vector<float>>((size_t)cnt[k] * desc_dim);
    for (int i = 0; i < cnt[k]; i++)
      for (int d = 0; d < desc_dim; d++) (*blk)[(size_t)i * desc_dim + d] = src[k][i].desc[d];
    const std::shared_ptr<const std::vector<float>> cblk = blk;
    for (int i = 0; i < cnt[k]; i++) {
      const modsgpu_feature& f = src[k][i];
      AffineKeypoint& kp = l[k][i].reproj_kp;
      kp.x = f.x; kp.y = f.y; kp.s = f.s; kp.a11 = f.a11; kp.a12 = f.a12; kp.a21 = f.a21; kp.a22 = f.a22;
      kp.response = f.response; kp.octave_number = f.octave; kp.sub_type = f.type;
      l[k][i].det_kp = kp;
      l[k][i].id = i; l[k][i].img_id = k; l[k][i].img_reproj_id = f.view; l[k][i].type = f.type;
      l[k][i].desc.view(cblk, (size_t)i * desc_dim, (size_t)desc_dim);
    }
  }
}
// DuplicateFiltering + LORANSACFiltering + result packing shared by modsgpu_match_features / modsgpu_verify_matches
static int verify_tentatives(modsgpu_ctx* ctx, modsb200::TentativeCorrespListExt& tent, int use_F, unsigned long long seed,
                             modsgpu_mods_result* res, double* inlier_xy, int capacity) {
  using namespace modsb200;
  MatchPars mp;
  TentativeCorrespListExt verified;
  int nu = DuplicateFiltering(ctx, tent, mp.doubleFilteringRadius);
  if (nu < 0) return nu;
  res->unique_tentatives = nu;
  RANSACPars rp;
  rp.seed = seed; rp.useF = use_F;
  int ni = LORANSACFiltering(ctx, tent, verified, res->model, rp);
  if (ni < 0) return ni;
  res->inliers = ni;
  for (int i = 0; i < ni && i < capacity && inlier_xy; i++) {
    const TentativeCorrespExt& c = verified.TCList[i];
    inlier_xy[4 * i + 0] = c.first.reproj_kp.x; inlier_xy[4 * i + 1] = c.first.reproj_kp.y;
    inlier_xy[4 * i + 2] = c.second.reproj_kp.x; inlier_xy[4 * i + 3] = c.second.reproj_kp.y;
  }
  return 0;
}
extern "C" int modsgpu_match_features(modsgpu_ctx* ctx, const modsgpu_feature* f1, int n1, const modsgpu_feature* f2, int n2,
                                      int desc_dim, double fginn_threshold, int use_F, unsigned long long seed,
                                      modsgpu_mods_result* res, double* inlier_xy, int capacity) {
  using namespace modsb200;
  if (!ctx || !res || n1 < 0 || n2 < 0 || (n1 > 0 && !f1) || (n2 > 0 && !f2) || desc_dim < 1 || desc_dim > 128 || capacity < 0)
    return MODSGPU_EINVAL;
  AffineRegionVector l[2];
  features_to_lists(f1, n1, f2, n2, desc_dim, l);
  memset(res, 0, sizeof(*res));
  res->steps_done = 1;
  res->regions[0] = n1; res->regions[1] = n2;
  MatchPars mp;
  mp.FGINNThreshold = fginn_threshold;
  TentativeCorrespListExt tent;
  int nt = MatchFlannFGINN(ctx, l[0], l[1], tent, mp);
  if (nt < 0) return nt;
  res->tentatives = nt;
  return verify_tentatives(ctx, tent, use_F, seed, res, inlier_xy, capacity);
}
// The second half of modsgpu_match_features on tentatives matched elsewhere -- e.g. by several GPUs, each running
// modsgpu_match_fginn on its slice of the query rows (mods_light_zmq_b200/mods_dist.py): m[k].qi / .ti index f1 / f2, the
// list must be in query order (the order MatchFlannFGINN appends in, matching.cpp:430-455).
extern "C" int modsgpu_verify_matches(modsgpu_ctx* ctx, const modsgpu_feature* f1, int n1, const modsgpu_feature* f2, int n2,
                                      const modsgpu_match* m, int nm, int use_F, unsigned long long seed, modsgpu_mods_result* res,
                                      double* inlier_xy, int capacity) {
  using namespace modsb200;
  if (!ctx || !res || n1 < 0 || n2 < 0 || nm < 0 || (n1 > 0 && !f1) || (n2 > 0 && !f2) || (nm > 0 && !m) || capacity < 0) return MODSGPU_EINVAL;
  memset(res, 0, sizeof(*res));
  res->steps_done = 1;
  res->regions[0] = n1; res->regions[1] = n2;
  // only the matched regions are materialised, and without their descriptors: nothing after the matcher reads them
  auto region_of = [](const modsgpu_feature& f, int id, int img) {
    AffineRegion r;
    AffineKeypoint& kp = r.reproj_kp;
    kp.x = f.x; kp.y = f.y; kp.s = f.s; kp.a11 = f.a11; kp.a12 = f.a12; kp.a21 = f.a21; kp.a22 = f.a22;
    kp.response = f.response; kp.octave_number = f.octave; kp.sub_type = f.type;
    r.det_kp = kp;
    r.id = id; r.img_id = img; r.img_reproj_id = f.view; r.type = f.type;
    return r;
  };
  TentativeCorrespListExt tent;
  tent.TCList.reserve(nm);
  for (int k = 0; k < nm; k++) {
    if (m[k].qi < 0 || m[k].qi >= n1 || m[k].ti < 0 || m[k].ti >= n2) return MODSGPU_EINVAL;
    TentativeCorrespExt tc;
    tc.first = region_of(f1[m[k].qi], m[k].qi, 0);
    tc.second = region_of(f2[m[k].ti], m[k].ti, 1);
    tc.secondbad_idx = m[k].tj_bad;
    tc.d1 = m[k].d1; tc.d2 = m[k].d2; tc.ratio = m[k].ratio;
    tent.TCList.push_back(tc);
  }
  res->tentatives = nm;
  return verify_tentatives(ctx, tent, use_F, seed, res, inlier_xy, capacity);
}

// MODS run over an iteration schedule (mods.cpp:202-356, HessianAffine steps)
extern "C" int modsgpu_mods_pair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const modsgpu_mods_step* steps,
                                 int n_steps, int min_matches, int use_F, unsigned long long seed, modsgpu_mods_result* res,
                                 double* inlier_xy, int capacity) {
  modsgpu_pipeline_params p;
  modsgpu_default_pipeline_params(&p);
  p.seed = seed; p.use_F = use_F;
  return modsgpu_mods_pair_ex(ctx, img1, img2, steps, n_steps, min_matches, &p, res, inlier_xy, capacity);
}
extern "C" int modsgpu_mods_pair_ex(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const modsgpu_mods_step* steps,
                                    int n_steps, int min_matches, const modsgpu_pipeline_params* pp, modsgpu_mods_result* res,
                                    double* inlier_xy, int capacity) {
  using namespace modsb200;
  if (!ctx || !img1 || !img2 || !res || !pp || n_steps < 0 || (n_steps > 0 && !steps)) return MODSGPU_EINVAL;
  std::vector<IterationStep> st((size_t)n_steps);
  for (int i = 0; i < n_steps; i++) {
    if (steps[i].n_scales < 0 || steps[i].n_scales > 8 || steps[i].n_tilts < 0 || steps[i].n_tilts > 8) return MODSGPU_EINVAL;
    st[i].ScaleSet.assign(steps[i].scale_set, steps[i].scale_set + steps[i].n_scales);
    st[i].TiltSet.assign(steps[i].tilt_set, steps[i].tilt_set + steps[i].n_tilts);
    st[i].Phi = steps[i].phi; st[i].initSigma = steps[i].init_sigma; st[i].FGINNThreshold = steps[i].fginn_threshold;
    st[i].doBlur = steps[i].do_blur;
  }
  DetectPars dp;
  MatchPars mp;
  RANSACPars rp;
  unpack_params(pp, dp, mp, rp);
  MODSResult r;
  TentativeCorrespListExt verified;
  int rc = MODSPair(ctx, img1, img2, st, min_matches, dp, mp, rp, r, verified);
  if (rc) return rc;
  res->steps_done = r.steps_done;
  for (int k = 0; k < 2; k++) { res->views[k] = r.views[k]; res->regions[k] = r.regions[k]; }
  res->tentatives = r.tentatives; res->unique_tentatives = r.unique_tentatives; res->inliers = r.inliers;
  for (int i = 0; i < 9; i++) res->model[i] = r.model[i];
  for (int i = 0; i < r.inliers && i < capacity && inlier_xy; i++) {
    const TentativeCorrespExt& c = verified.TCList[i];
    inlier_xy[4 * i + 0] = c.first.reproj_kp.x; inlier_xy[4 * i + 1] = c.first.reproj_kp.y;
    inlier_xy[4 * i + 2] = c.second.reproj_kp.x; inlier_xy[4 * i + 3] = c.second.reproj_kp.y;
  }
  return 0;
}
