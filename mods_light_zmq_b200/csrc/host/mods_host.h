// mods_host.h -- C++ host mirror of the reference's operator interface for the hot path, written on
// top of the C ABI (include/modsgpu.h).  Same names, argument meaning and return conventions as the
// reference so the call sites in mods.cpp / extract_features_batch.cpp read the same:
//   ImageRepresentation::SynthDetectDescribeKeypoints   imagerepresentation.cpp:686-1104 (identity view,
//       HessianAffine + AffNet + OriNet + HardNet++; the three DescribeWithZmq round trips
//       :800/:878/:995 become modsgpu_describe calls)
//   MatchFlannFGINN       matching.cpp:356-460      DuplicateFiltering  matching.cpp:2615-2679
//   LORANSACFiltering     matching.cpp:637-823      (H branch: NaiveHCheck :1014-1043, H_LAF_check :250-308)
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/modsgpu.h"

namespace modsb200 {

struct AffineKeypoint {            // structures.hpp:185-194
  double x = 0, y = 0, s = 0;
  double a11 = 1, a12 = 0, a21 = 0, a22 = 1;
  double response = 0;
  int octave_number = 0, sub_type = 0;
};
// descriptor.vec of a region.  Regions travel by value through the tentative / filtered / verified lists
// (TentativeCorrespExt holds two of them), so the vector is a VIEW into a shared block -- the n x 128 read-back of one
// modsgpu_describe call -- and a copy costs a reference count, not an allocation.
class DescVec {
 public:
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  const float* data() const { return blk_ ? blk_->data() + off_ : nullptr; }
  float operator[](size_t i) const { return (*blk_)[off_ + i]; }
  template <class It> void assign(It a, It b) {
    auto v = std::make_shared<std::vector<float>>(a, b);
    n_ = v->size(); off_ = 0; blk_ = std::move(v);
  }
  void view(const std::shared_ptr<const std::vector<float>>& blk, size_t off, size_t n) { blk_ = blk; off_ = off; n_ = n; }
  const std::vector<float>* block() const { return blk_.get(); }
  size_t offset() const { return off_; }
 private:
  std::shared_ptr<const std::vector<float>> blk_;
  size_t off_ = 0, n_ = 0;
};
struct AffineRegion {              // structures.hpp:218-229
  int img_id = 0, img_reproj_id = 0, id = 0, parent_id = 0, type = 0;
  AffineKeypoint det_kp, reproj_kp;
  DescVec desc;                    // descriptor.vec
};
typedef std::vector<AffineRegion> AffineRegionVector;

struct TentativeCorrespExt {       // matching.hpp:39-51
  AffineRegion first, second;
  int secondbad_idx = -1;
  double d1 = 0, d2 = 0, ratio = 0;
  int isTrue = 0;
};
struct TentativeCorrespListExt {
  std::vector<TentativeCorrespExt> TCList;
  double H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};

struct MatchPars {                 // matching.hpp:97-130, values of config_aff_ori_desc_zeromq.ini
  double FGINNThreshold = 0.8;     // iters_HessianZMQ.ini
  double contradDist = 10.0;
  int nn = 50;
  double doubleFilteringRadius = 2.0;   // mods.cpp:283
};
struct RANSACPars {                // matching.hpp:132-164
  double err_threshold = 4.0;
  double confidence = 0.99;
  int max_samples = 1000000;
  int doSymmCheck = 1;
  double HLAFCoef = 12.0;
  double LAFCoef = 2.0;            // F branch (matching.cpp:810)
  int errorType = 0;               // SAMPSON / SYMM_SUM / SYMM_MAX (matching.hpp:86-90) = MODSGPU_ERR_*
  int justMarkOutliers = 0;        // matching.cpp:751-762: keep every tentative in the output list, only flag isTrue
  int useF = 0;                    // LORANSACF (mods.cpp:325): exp_ransacFcustom instead of exp_ransacHcustom
  unsigned long long seed = 12345; // the reference seeds with time(NULL) (exp_ranH.c:823)
};
struct DetectPars {
  modsgpu_pyr_params pyr;          // [HessianAffine]
  double mrSize = 5.1962;          // AffNet / OriNet / desc patch extent
  int patchSize = 32;
  modsgpu_affshape_params aff;     // [HessianAffine] Baumberg block (classic configuration)
  int oriPatchSize = 32, maxAngles = 1;   // [DominantOrientation]
  double oriThreshold = 0.8;
  int siftPatchSize = 41, photoNorm = 1, rootSift = 1;   // [SIFTDescriptor] + iters_HessianSIFT.ini (RootSIFT)
  // identity-view extraction only: leave the descriptors on the device (ImageRepresentation::device_descriptors()) instead
  // of filling AffineRegion::desc -- for callers that only match them (MatchFlannFGINNDevice)
  bool desc_on_device = false;
  DetectPars() { modsgpu_default_pyr_params(&pyr); modsgpu_default_affshape_params(&aff); }
};

struct ViewSynthParameters {       // structures.hpp:196-209 (the fields the hot path reads)
  double zoom = 1, tilt = 1, phi = 0, InitSigma = 0.5;
  int doBlur = 1;
};
// synth-detection.cpp:191-322 SetVSPars: the view list of one iteration (ScaleSet x TiltSet x rotations with
// n_rot = floor(180 * tilt / Phi), a negative Phi meaning one vertical-tilt view); views already in prev_par are
// dropped and the new ones appended to it.
int SetVSPars(const std::vector<double>& scale_set, const std::vector<double>& tilt_set, double phi_base,
              std::vector<ViewSynthParameters>& par, std::vector<ViewSynthParameters>& prev_par, double InitSigma, int doBlur);

struct TimeLog {                   // structures.hpp:33-56 (device + host ms per stage)
  double SynthTime = 0, DetectTime = 0, OrientTime = 0, DescTime = 0, MatchingTime = 0, RANSACTime = 0;
};

typedef std::map<std::string, AffineRegionVector> AffineRegionVectorMap;   // descriptor name -> regions (structures.hpp:231)
struct WhatToMatch {               // structures.hpp:236-242: which (detector, descriptor) lists one iteration matches
  std::vector<std::string> group_detectors, group_descriptors, separate_detectors, separate_descriptors;
};

class ImageRepresentation {
 public:
  ImageRepresentation(modsgpu_ctx* ctx, modsgpu_image* img, bool owns_image);
  ~ImageRepresentation();
  ImageRepresentation(const ImageRepresentation&) = delete;              // owns device handles
  ImageRepresentation& operator=(const ImageRepresentation&) = delete;
  // the keyed region store of the reference (imagerepresentation.h:64, RegionVectorMap[detector][descriptor]).  The lists
  // this object extracts itself are kept in regions_ under (det_name, desc_name) = ("HessianAffine", "ZMQ" | "RootSIFT");
  // AddRegions (imagerepresentation.cpp:637-660) files further lists, e.g. pre-extracted ones, under their own keys.
  std::map<std::string, AffineRegionVectorMap> RegionVectorMap;
  std::string det_name = "HessianAffine", desc_name = "ZMQ";
  void AddRegions(const AffineRegionVector& RegionsToAdd, const std::string& det, const std::string& desc);
  // imagerepresentation.cpp:600-635 for a named detector (MatchImgReps never asks for "All")
  AffineRegionVector GetAffineRegionVector(const std::string& desc, const std::string& det) const;
  // returns the number of described regions, < 0 on error (message via modsgpu_last_error)
  int SynthDetectDescribeKeypoints(const DetectPars& par);
  // the same over a list of synthesised views (imagerepresentation.cpp:704-1102): every view is generated on the
  // device (modsgpu_synth_view), detected / described in view coordinates and reprojected to the original image
  // (ReprojectByH synth-detection.cpp:578-587).  Regions of a view are appended ONCE (the reference's AddRegions
  // loop sits inside the view loop and appends view k up to n-k times, SURVEY Q1 -- documented deviation).
  int SynthDetectDescribeKeypoints(const std::vector<ViewSynthParameters>& views, const DetectPars& par);
  int n_views = 0;
  const AffineRegionVector& GetAffineRegionVector() const { return regions_; }
  // regions of further views are appended (AddRegions, imagerepresentation.cpp:1098-1102)
  int AddViews(const std::vector<ViewSynthParameters>& views, const DetectPars& par);
  // the classic configuration (config_affori_classic.ini + iters_HessianSIFT.ini), identity view: Hessian-Affine with
  // the in-pyramid Baumberg iteration -> dominant orientation (DetectOrientation) -> RootSIFT (DescribeRegions)
  int SynthDetectDescribeKeypointsClassic(const DetectPars& par);
  int n_keypoints = 0, n_affine = 0;
  TimeLog TimeSpent;
  // descriptor rows of GetAffineRegionVector() in list order, on the device (DetectPars::desc_on_device); NULL otherwise
  const modsgpu_devdesc* device_descriptors() const { return devdesc_; }

 private:
  modsgpu_devdesc* devdesc_ = nullptr;
  modsgpu_ctx* ctx_;
  modsgpu_image* img_;
  bool owns_;
  AffineRegionVector regions_;
  int DescribeView(modsgpu_image* view, const double* H, int orig_w, int orig_h, const DetectPars& par, AffineRegionVector& out);
};

// helpers.cpp:524-549 / :401-410 / :504-515
bool interpolateCheckBorders(int orig_img_w, int orig_img_h, float ofsx, float ofsy, float a11, float a12, float a21,
                             float a22, int res_w, int res_h);
void rectifyAffineTransformationUpIsUp(double& a11, double& a12, double& a21, double& a22);
bool getEigenvalues(float a, float b, float c, float d, float& l1, float& l2);

// imagerepresentation.cpp:113-126 saveKP_KM_format + :205-211 saveAR_KM_format + :1187-1213 SaveRegionsMichal (text mode)
int SaveRegionsMichal(const AffineRegionVector& regions, const std::string& fname);
void OxAffEllipse(const AffineKeypoint& k, float& a, float& b, float& c);

int MatchFlannFGINN(modsgpu_ctx* ctx, const AffineRegionVector& list1, const AffineRegionVector& list2,
                    TentativeCorrespListExt& corresp, const MatchPars& par);
// the same over two images whose descriptors stayed on the device (DetectPars::desc_on_device)
int MatchFlannFGINNDevice(modsgpu_ctx* ctx, const ImageRepresentation& img1, const ImageRepresentation& img2,
                          TentativeCorrespListExt& corresp, const MatchPars& par);
// MatchFlannFGINNDevice + DuplicateFiltering in one device call (modsgpu_match_dedup_dev): `corresp` receives the filtered
// list, *n_tentatives the length before the filter; returns the filtered length
int MatchDedupDevice(modsgpu_ctx* ctx, const ImageRepresentation& img1, const ImageRepresentation& img2,
                     TentativeCorrespListExt& corresp, const MatchPars& par, int* n_tentatives);

// correspondencebank.h / .cpp: tentatives filed per (descriptor, detector | "Group").  MatchImgReps
// (correspondencebank.cpp:234-343): GROUPED -- for every group descriptor the regions of all group detectors are pooled
// (image 2 = train, image 1 = query) and matched once; SEPARATE -- every separate detector x separate descriptor is
// matched on its own.  fginn: FGINNThreshold per descriptor name (iters file); a descriptor without an entry, or with a
// threshold <= 0, is not matched (correspondencebank.cpp:262-277).
class CorrespondenceBank {
 public:
  explicit CorrespondenceBank(modsgpu_ctx* ctx) : ctx_(ctx) {}
  int MatchImgReps(const ImageRepresentation& imgrep1, const ImageRepresentation& imgrep2, const WhatToMatch& what,
                   const MatchPars& par, const std::map<std::string, double>& fginn);
  TentativeCorrespListExt GetCorresponcesVector(const std::string& desc = "All", const std::string& det = "All") const;
  int GetCorrespondencesNumber(const std::string& desc = "All", const std::string& det = "All") const;
  void AddCorrespondences(const TentativeCorrespListExt& CorrsToAdd, const std::string& det, const std::string& desc);
  void ClearCorrespondences(const std::string& det, const std::string& desc);
 private:
  modsgpu_ctx* ctx_;
  std::map<std::string, std::map<std::string, TentativeCorrespListExt>> CorrespondencesMapMap;   // [descriptor][detector]
};
int DuplicateFiltering(modsgpu_ctx* ctx, TentativeCorrespListExt& in_corresp, double r);
// one MODS run over a schedule of iterations (mods.cpp:202-356 for the HessianAffine steps): every step adds the
// regions of its new views to both images, matches ALL accumulated regions, filters duplicates and verifies with
// LO-RANSAC (H or F); the loop stops at the first step with >= minMatches verified correspondences.
struct IterationStep { std::vector<double> ScaleSet, TiltSet; double Phi = 360, initSigma = 0.2, FGINNThreshold = 0.8; int doBlur = 1; };
struct MODSResult { int steps_done = 0, views[2] = {0, 0}, regions[2] = {0, 0}, tentatives = 0, unique_tentatives = 0, inliers = 0; double model[9] = {0}; };
int MODSPair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const std::vector<IterationStep>& steps,
             int minMatches, const RANSACPars& rp, MODSResult& res, TentativeCorrespListExt& verified);
int MODSPair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const std::vector<IterationStep>& steps,
             int minMatches, const DetectPars& dp, const MatchPars& mp, const RANSACPars& rp, MODSResult& res,
             TentativeCorrespListExt& verified);

int LORANSACFiltering(modsgpu_ctx* ctx, TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& ransac_corresp,
                      double* H, const RANSACPars& pars);
// matching.cpp:917-1013 (verification against a known homography) and :574-633 (binary descriptors, Hamming 2-NN)
int HMatrixFiltering(TentativeCorrespListExt& in_corresp, TentativeCorrespListExt& true_corresp, const double* H, int isExtended,
                     const RANSACPars& pars);
int MatchFLANNDistance(modsgpu_ctx* ctx, const AffineRegionVector& list1, const AffineRegionVector& list2,
                       TentativeCorrespListExt& corresp, double matchDistanceThreshold);
// the empirical checks at the end of LORANSACFiltering (matching.cpp:764-820); Hloran = the degensac-convention model
int EmpiricalChecks(TentativeCorrespListExt& ransac_corresp, const double* Hloran, const RANSACPars& pars, double* H);

}  // namespace modsb200
