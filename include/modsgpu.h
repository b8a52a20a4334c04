/*
 * modsgpu.h -- C ABI of libmodsgpu.so: the B200 (sm_100a) implementation of the MODS
 * per-view hot path  detect -> (AffNet) -> orient (OriNet) -> describe (HardNet++) ->
 * match (FGINN) -> LO-RANSAC.
 *
 * Plain C: pointers and sizes only.  Every entry point names the reference interface it
 * replaces (file:line in ducha-aiki/mods-light-zmq @ 33c9ba2).  INTEGRATION.md shows the
 * C++ stubs a maintainer adds on the reference side to bind them.
 *
 * Conventions (mirroring the reference: integer returns, no exceptions):
 *   return 0 on success, a negative MODSGPU_E* code on failure; modsgpu_last_error() gives
 *   the message.  Host pointers unless a parameter is called "device".  A modsgpu_ctx owns
 *   one CUDA stream and its workspaces and is NOT thread-safe: use one ctx per calling
 *   thread (the reference calls the per-image pipeline from one OpenMP task per image,
 *   mods.cpp:234-251).  The library never falls back to a CPU path: if no sm_100 device is
 *   present modsgpu_create() fails with MODSGPU_ENODEV.
 */
#ifndef MODSGPU_H
#define MODSGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MODSGPU_OK        0
#define MODSGPU_ENODEV   -1  /* no CUDA device / not sm_100                        */
#define MODSGPU_ECUDA    -2  /* CUDA runtime error (message in modsgpu_last_error) */
#define MODSGPU_EINVAL   -3  /* bad argument                                       */
#define MODSGPU_EIO      -4  /* weight file unreadable / malformed                 */
#define MODSGPU_ESTATE   -5  /* e.g. describe before modsgpu_load_weights          */

typedef struct modsgpu_ctx modsgpu_ctx;
typedef struct modsgpu_image modsgpu_image;   /* device-resident fp32 gray image */

/* ---- context ------------------------------------------------------------------------- */
int  modsgpu_create(int device, modsgpu_ctx** out);
void modsgpu_destroy(modsgpu_ctx* ctx);
const char* modsgpu_last_error(const modsgpu_ctx* ctx);
const char* modsgpu_version(void);
/* device-side time (ms, CUDA events on the ctx stream) of the last entry point called */
float modsgpu_last_device_ms(const modsgpu_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
long long modsgpu_launch_count(const modsgpu_ctx* ctx);
/* raw stream handle (cudaStream_t) so callers can record their own events */
void* modsgpu_stream(const modsgpu_ctx* ctx);

/* ---- image (replaces ImageRepresentation::ImageRepresentation imagerepresentation.cpp:293-302
 *      + the gray conversion of GenerateSynthImageCorr synth-detection.cpp:344-354) -------- */
/* 8-bit interleaved BGR (cv::imread layout) -> fp32 gray = (B+G+R)/3 on the device */
int  modsgpu_image_from_bgr8(modsgpu_ctx* ctx, const uint8_t* bgr, int w, int h, modsgpu_image** out);
/* fp32 gray, `stride` in floats */
int  modsgpu_image_from_gray32f(modsgpu_ctx* ctx, const float* gray, int w, int h, int stride, modsgpu_image** out);
int  modsgpu_image_download(modsgpu_ctx* ctx, const modsgpu_image* img, float* gray /* w*h */);
void modsgpu_image_size(const modsgpu_image* img, int* w, int* h);
void modsgpu_image_free(modsgpu_ctx* ctx, modsgpu_image* img);

/* ---- view synthesis (replaces GenerateSynthImageCorr synth-detection.cpp:324-518: rotate by phi, anisotropic
 *      anti-aliasing blur, tilt / zoom; tilt < 0 = vertical tilt; phi in [0, pi); the identity view is a copy).
 *      H: 9 doubles, row-major, original -> view (SynthImage::H).  Arithmetic = OpenCV 4.x warpAffine /
 *      GaussianBlur for CV_32F, bit-exact against cv2 4.13 (tests/golden/synth_pins.npz). ----------------------- */
int  modsgpu_synth_geometry(int w, int h, double tilt, double phi, double zoom, int* ow, int* oh, double* H);
int  modsgpu_synth_view(modsgpu_ctx* ctx, const modsgpu_image* in, double tilt, double phi, double zoom,
                        double InitSigma, int doBlur, modsgpu_image** out, double* H);

/* ---- S1 detector (replaces DetectAffineKeypoints scale-space-detector.cpp:13-32 ->
 *      ScaleSpaceDetector::detectPyramidKeypoints pyramid.cpp:496-529, for DET_HESSIAN,
 *      FIXED_TH, doBaumberg = 0) ------------------------------------------------------------ */
typedef struct {            /* PyramidParams, structures.hpp:114-151 */
  int    numberOfScales;        /* 3    */
  float  initialSigma;          /* 1.6  */
  float  threshold;             /* 5.33 ([HessianAffine] threshold in the ini) */
  double edgeEigenValueRatio;   /* 10   */
  int    border;                /* 5    */
  /* detection mode (structures.hpp:10-14, prepareKeysForExport scale-space-detector.hpp:125-198).  Every mode but
   * FIXED_TH detects with all thresholds at 0 (pyramid.h:58-59) and truncates the |response|-sorted list:
   *   RELATIVE_TH keeps |r| > max|r| * rel_threshold;  FIXED_REG_NUMBER keeps reg_number;  RELATIVE_REG_NUMBER keeps
   *   floor(rel_reg_number * n);  NOT_LESS_THAN_REGIONS keeps max(reg_number, #{|r| > threshold}) (capped at n). */
  int    detectorMode;          /* MODSGPU_FIXED_TH */
  float  rel_threshold;         /* -1 */
  int    reg_number;            /* -1 */
  float  rel_reg_number;        /* -1 */
} modsgpu_pyr_params;
enum { MODSGPU_FIXED_TH = 0, MODSGPU_RELATIVE_TH = 1, MODSGPU_FIXED_REG_NUMBER = 2, MODSGPU_RELATIVE_REG_NUMBER = 3,
       MODSGPU_NOT_LESS_THAN_REGIONS = 4 };
/* DetectAffineKeypoints (scale-space-detector.cpp:13-32): reg_number shrinks on strongly tilted / zoomed-out views */
int  modsgpu_reg_number_for_view(int reg_number, double tilt, double zoom);

typedef struct {            /* AffineKeypoint structures.hpp:185-194 + provenance for parity tests */
  float x, y, s;                /* pyramid.cpp:392-402 */
  float response;
  int   type;                   /* 0 dark blob, 1 bright blob, 2 saddle (pyramid.cpp:65-124) */
  int   octave;                 /* 0,1,..  (pixelDistance = 2^octave)                        */
  int   level;                  /* 1..numberOfScales                                          */
  int   r0, c0;                 /* raster position of the NMS candidate (findLevelKeypoints)  */
  int   r, c;                   /* integer position after localizeKeypoint                    */
  int   seq;                    /* reserved (0)                                               */
} modsgpu_keypoint;

void modsgpu_default_pyr_params(modsgpu_pyr_params* p);
/* Output order = the reference's export order: |response| descending
 * (scale-space-detector.hpp:120-131), ties by (octave, level, r0, c0).
 * *out is malloc()ed by the library; release with modsgpu_free(). */
int  modsgpu_detect(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p,
                    modsgpu_keypoint** out, int* n);
void modsgpu_free(void* p);

/* the same detector with the in-pyramid affine shape adaptation of the classic configuration (doBaumberg = 1,
 * method SMM: AffineShape::findAffineShape affine.cpp:26-158, called from localizeKeypoint on the level below the
 * response level, pyramid.cpp:402).  Keypoints whose iteration does not converge are dropped.  *A: 4 floats per
 * keypoint (a11 a12 a21 a22 as handed to onAffineShapeFound); both arrays are malloc()ed, release with modsgpu_free. */
typedef struct {            /* AffineShapeParams, affine.h:26-68 */
  int   maxIterations;          /* 16   */
  float convergenceThreshold;   /* 0.05 */
  int   smmWindowSize;          /* 19   */
  float initialSigma;           /* 1.6  */
  int   doBaumberg;             /* 1    */
} modsgpu_affshape_params;
void modsgpu_default_affshape_params(modsgpu_affshape_params* a);
int  modsgpu_detect_affine(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_pyr_params* p,
                           const modsgpu_affshape_params* aff, modsgpu_keypoint** out, float** A, int* n);

/* pyramid internals exposed for the parity tests only (helpers.cpp:717-731 gaussianBlur,
 * pyramid.cpp:196-254 HessianResponse, pyramid.cpp:476 cv::resize 0.5) */
int  modsgpu_gaussian_blur(modsgpu_ctx* ctx, const float* in, float* out, int w, int h, float sigma);
int  modsgpu_hessian_response(modsgpu_ctx* ctx, const float* in, float* out, int w, int h, float norm);
int  modsgpu_half_image(modsgpu_ctx* ctx, const float* in, int w, int h, float* out /* round(w/2)*round(h/2) */);

/* ---- regions ---------------------------------------------------------------------------- */
typedef struct {            /* the AffineKeypoint fields the sampler reads, structures.hpp:185-194 */
  double x, y, s;
  double a11, a12, a21, a22;
} modsgpu_region;

/* ---- S5 patch sampler (replaces ExtractPatchesColumn synth-detection.cpp:38-132 with
 *      fast_extraction=false, photoNorm=false, followed by the float->u8 conversion of
 *      cv::imencode(".png") imagerepresentation.cpp:45).  out: n * patchSize*patchSize bytes. */
int  modsgpu_extract_patches(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                             double mrSize, int patchSize, uint8_t* out);

/* the same sampler without the u8 quantisation: n * patchSize*patchSize floats (DescribeRegions' patches) */
int  modsgpu_extract_patches_f32(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                 double mrSize, int patchSize, float* out);

/* ---- S2 CNNs (replace DescribeWithZmq imagerepresentation.cpp:21-103 and the three daemons
 *      build/affnet_server.py, orinet_server.py, desc_server.py) ------------------------------ */
typedef enum { MODSGPU_AFFNET = 0, MODSGPU_ORINET = 1, MODSGPU_HARDNET = 2 } modsgpu_net;
/* output floats per region: 3 (AffNet: a11, a21, a22 with +1 on a11,a22), 2 (OriNet: sin-like,
 * cos-like), 128 (HardNet++: uint8-valued floats, desc_server.py:42) */
int  modsgpu_net_out_dim(modsgpu_net net);
/* .npz written by tools/export_weights.py from build/{AffNet,OriNet,HardNet++}.pth */
int  modsgpu_load_weights(modsgpu_ctx* ctx, modsgpu_net net, const char* npz_path);
/* patches -> net, no 2000-region batching limit.  out: n * out_dim floats. */
int  modsgpu_describe(modsgpu_ctx* ctx, modsgpu_net net, const modsgpu_image* img, const modsgpu_region* regs,
                      int n, double mrSize, int patchSize, float* out);
/* the net alone on caller-supplied 32x32 u8 patches (what the daemons receive as PNG) */
int  modsgpu_net_forward_u8(modsgpu_ctx* ctx, modsgpu_net net, const uint8_t* patches, int n, float* out);

/* ---- classic per-region stages (config_affori_classic.ini) ----------------------------------
 * dominant orientation: replaces DetectOrientation (synth-detection.cpp:1039-1149, doHalfSIFT = 0, addUpRight = false).
 *   n_ang[i] = -1 when region i is dropped by the 2*3*sqrt(3)*s frame test, else the number of angles (<= maxAngles,
 *   taken in bin order) written to angles[i*maxAngles ..]; the caller rotates A by -angle like :1091-1101.
 * (Root)SIFT: replaces DescribeRegions<SIFTDescriptor> (synth-detection.hpp:170-263, FastPatchExtraction = false) with
 *   matching/siftdesc.cpp: out = n x 128 floats holding integers 0..255. */
int  modsgpu_dominant_orientation(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                                  double mrSize, int patchSize, int maxAngles, double th, int* n_ang, float* angles);
int  modsgpu_describe_sift(modsgpu_ctx* ctx, const modsgpu_image* img, const modsgpu_region* regs, int n,
                           double mrSize, int patchSize, int photoNorm, int rootSift, float* out);

/* ---- S3 matcher (replaces MatchFlannFGINN matching.cpp:356-460 with vector_matcher=linear:
 *      exact 50-NN + first-geometrically-inconsistent ratio test) ----------------------------- */
typedef struct {            /* TentativeCorrespExt fields filled at matching.cpp:437-449 */
  int    qi, ti, tj_bad;        /* query idx, 1st NN, the ratio-passing (geometrically inconsistent) NN */
  float  d1, d2;                /* squared L2 distances to ti and tj_bad */
  int    _pad;
  double ratio;                 /* sqrt(d1/d2) */
} modsgpu_match;
/* q: nq x dim, t: nt x dim row-major floats holding integers in [0,255] (both reference descriptors
 * do); txy: nt x 2 doubles (train keypoint positions).  out: capacity nq.
 * knn_idx/knn_dist (nq x nn, may be NULL) receive the ordered neighbour lists (dist asc, idx asc). */
int  modsgpu_match_fginn(modsgpu_ctx* ctx, const float* q, int nq, const float* t, const double* txy, int nt,
                         int dim, double ratio_thr, double contrad_dist, int nn,
                         modsgpu_match* out, int* nout, int* knn_idx, float* knn_dist);

/* ---- binary descriptors (replaces MatchFLANNDistance matching.cpp:574-633 with binary_matcher = linear): the two nearest
 *      train descriptors by Hamming distance over bytes = floor(entry); a match when d1 <= max_distance; ratio = d1 / d2
 *      (tj_bad = the second neighbour).  q: nq x dim, t: nt x dim floats, dim = descriptor length in BYTES (32 for ORB),
 *      a multiple of 4, <= 128.  Ties keep the lower train index (cvflann LinearIndex order). ------------------------ */
int  modsgpu_match_hamming(modsgpu_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim, double max_distance,
                           modsgpu_match* out, int* nout);

/* ---- duplicate filter (replaces DuplicateFiltering matching.cpp:2615-2679, mode bestFGINN;
 *      stable order on ties).  order_out: indices of the survivors in sorted order. ---------- */
int  modsgpu_duplicate_filter(modsgpu_ctx* ctx, const double* xy1, const double* xy2, const double* ratio,
                              int T, double r, int* order_out, int* nout);

/* ---- S4 LO-RANSAC (replaces exp_ransacHcustom degensac/exp_ranH.h:53-57 as called from
 *      LORANSACFiltering matching.cpp:731).  Batched, counter-based RNG -> reproducible. ------- */
typedef struct {
  double th;                    /* squared pixel threshold (matching.cpp:731 passes err_threshold^2) */
  double conf;                  /* 0.99 */
  int    max_samples;           /* 1e6; matching.cpp:644-645 clamps to 1000 when T <= 20 */
  int    do_sym_check;          /* matching.cpp:652-681 */
  uint64_t seed;                /* the reference seeds with time(NULL), exp_ranH.c:823 */
  int    error_type;            /* RANSACPars::errorType (matching.cpp:652-681): MODSGPU_ERR_SAMPSON (HDs / FDs),
                                 * MODSGPU_ERR_SYMM_MAX (HDsSymMax), MODSGPU_ERR_SYMM_SUM (HDsSym).  The F estimator
                                 * implements Sampson only and returns MODSGPU_EINVAL for the other two. */
  int    _pad;
} modsgpu_ransac_params;
enum { MODSGPU_ERR_SAMPSON = 0, MODSGPU_ERR_SYMM_MAX = 1, MODSGPU_ERR_SYMM_SUM = 2 };   /* RANSAC_error_t, matching.hpp:95 */
typedef struct {
  int    n_inliers;             /* Score.I */
  double J;                     /* Score.J (MSAC cost) */
  int    samples;               /* data_out[0] */
  int    lo_runs;               /* data_out[1] */
  int    oc_rejects;            /* data_out[2] */
  int    degen_runs;            /* F only: completed DEGENSAC (plane-and-parallax) passes, exp_ranF.c degen_cnt */
  int    h_inliers;             /* F only: *Ih, the largest plane consensus seen (exp_ranF.c:980) */
} modsgpu_ransac_result;
/* u: T x 6 doubles (x1,y1,1,x2,y2,1) as packed at matching.cpp:695-713.  H: 9 doubles in the
 * degensac convention (column-major / transposed, maps image 2 -> image 1).
 * inl: T bytes. */
int  modsgpu_ransac_H(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                      double* H, unsigned char* inl, modsgpu_ransac_result* res);
/* the same, also returning the error of every correspondence under H (the *resids array of exp_ransacHcustom) */
int  modsgpu_ransac_H_resid(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                            double* H, unsigned char* inl, modsgpu_ransac_result* res, double* resid /* T */);

/* the empirical checks LORANSACFiltering applies to degensac's inliers (matching.cpp:764-820): H = inv(Hloran^T),
 * NaiveHCheck (:1014-1043), H_LAF_check (:250-308, HDsSymMax on the three LAF points, 3*HLAFCoef*err_threshold) or
 * F_LAF_check (:192-249, FDs, LAFCoef*err_threshold); fewer than 8 survivors empty the list.  Host arithmetic only (no
 * device, no ctx).  kp1 / kp2: reproj_kp of the n RANSAC inliers; model: the degensac-convention H or F; laf_coef:
 * HLAFCoef (12) or LAFCoef (2); keep: n bytes; model_out: H row-major image 1 -> 2, or F unchanged. */
int  modsgpu_empirical_checks(const modsgpu_region* kp1, const modsgpu_region* kp2, int n, const double* model, int use_F,
                              double err_threshold, double laf_coef, unsigned char* keep, double* model_out, int* n_out);

/* verification against a KNOWN homography (replaces HMatrixFiltering matching.cpp:917-1013, ver_type GR_TRUTH of
 * mods.cpp:292-303): error of every tentative under H (HDs / HDsSymMax / HDsSym by error_type) against err_threshold^2.
 * The reference packs u = (image 2, image 1) here; H is read in the degensac layout for that order.  Host arithmetic only.
 * xy1 / xy2: n x 2 doubles; keep: n bytes; H_out (may be NULL): H transposed, as true_corresp.H receives it. */
int  modsgpu_hmatrix_filter(const double* xy1, const double* xy2, int n, const double* H, int error_type, double err_threshold,
                            unsigned char* keep, double* H_out, int* n_out);

/* replaces exp_ransacFcustom degensac/exp_ranF.h:71-73 as called from LORANSACFiltering matching.cpp:722
 * (7-point sample, oriented epipolar constraint, Sampson error, MSAC, symmetric check, LO).  F: 9 doubles with
 * u2^T M u1 = 0, M[k][l] = F[3k+l] (the degensac convention, Ftools.c:15-37).  res->oc_rejects counts models
 * discarded by the symmetric check.  Includes the DEGENSAC branch (plane-dominated samples -> inner H-RANSAC ->
 * plane-and-parallax, exp_ranF.c:963-1016); MODSGPU_NO_DEGENSAC=1 in the environment switches it off. */
int  modsgpu_ransac_F(modsgpu_ctx* ctx, const double* u, int T, const modsgpu_ransac_params* p,
                      double* F, unsigned char* inl, modsgpu_ransac_result* res);

/* ---- one VIEW from pixels to described regions on the device (the per-view body of
 *      ImageRepresentation::SynthDetectDescribeKeypoints, imagerepresentation.cpp:704-1006: DetectAffineRegions :739,
 *      DescribeWithZmq(AffNet) + rectify / eigen-ratio / frame tests :797-845, ReprojectRegionsAndRemoveTouchBoundary :868,
 *      DescribeWithZmq(OriNet) + rotation :876-899, ReprojectRegions :951, DescribeWithZmq(desc) :992-1006).
 *      Same results as modsgpu_detect + 3 x modsgpu_describe with the host arithmetic in between, but the region list
 *      never leaves the device between the stages: two host synchronisations per view instead of four, no host-side
 *      sampler preparation.  H: 9 doubles row-major original -> view (NULL = identity view); orig_w/h: the ORIGINAL image.
 *      *regions: n rows, *desc: n x 128 floats holding integers 0..255; both malloc()ed, release with modsgpu_free().
 *      counts (may be NULL): [0] raw keypoints, [1] regions after AffNet's tests, [2] described regions. ------------- */
typedef struct {
  modsgpu_region det;       /* det_kp: view coordinates (structures.hpp:218-229)   */
  modsgpu_region reproj;    /* reproj_kp: original image (ReprojectByH)            */
  double response;
  int    octave, type;
} modsgpu_view_region;
int  modsgpu_describe_view(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                           const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                           float** desc, int* n, int* counts);

/* The same with the descriptors left ON THE DEVICE (what a caller that only matches them wants: 0.5 MB instead of 2.7 MB
 * read back per 4k-keypoint view and nothing uploaded again).  *desc (NULL when the view has no keypoint) is released with
 * modsgpu_devdesc_free on the context that made it; modsgpu_match_fginn_dev = modsgpu_match_fginn over two such blocks
 * (any context of the same device); modsgpu_devdesc_download copies the n x 128 floats out (tests). */
typedef struct modsgpu_devdesc modsgpu_devdesc;
int  modsgpu_describe_view_dev(modsgpu_ctx* ctx, const modsgpu_image* view, const double* H, int orig_w, int orig_h,
                               const modsgpu_pyr_params* p, double mrSize, int patchSize, modsgpu_view_region** regions,
                               modsgpu_devdesc** desc, int* n, int* counts);
int  modsgpu_devdesc_size(const modsgpu_devdesc* desc);
int  modsgpu_devdesc_download(modsgpu_ctx* ctx, const modsgpu_devdesc* desc, float* out);
void modsgpu_devdesc_free(modsgpu_ctx* ctx, modsgpu_devdesc* desc);
int  modsgpu_match_fginn_dev(modsgpu_ctx* ctx, const modsgpu_devdesc* q, const modsgpu_devdesc* t, const double* txy,
                             double ratio_thr, double contrad_dist, int nn, modsgpu_match* out, int* nout);
/* MatchFlannFGINN + DuplicateFiltering (matching.cpp:356-460, :2615-2679, mode bestFGINN) in ONE call: the tentative list
 * stays on the device between the two, one read-back.  matches (capacity nq) = the tentatives in query order; order
 * (capacity nq) = indices into matches of the survivors, in the order DuplicateFiltering returns them.  qxy: nq x 2
 * doubles (query positions), dup_radius <= 0 disables the filter.  Equal to modsgpu_match_fginn_dev followed by
 * modsgpu_duplicate_filter; refuses nq > 16384 (use the two calls). */
int  modsgpu_match_dedup_dev(modsgpu_ctx* ctx, const modsgpu_devdesc* q, const modsgpu_devdesc* t, const double* qxy,
                             const double* txy, double ratio_thr, double contrad_dist, int nn, double dup_radius,
                             modsgpu_match* matches, int* n_matches, int* order, int* n_unique);
/* test-only: the two device post-processing steps of the chain on caller-supplied net outputs (survivors, order kept) */
int  modsgpu_debug_affnet_post(modsgpu_ctx* ctx, const modsgpu_view_region* regs, const float* aff, int n, int w, int h,
                               int orig_w, int orig_h, double mrSize, const double* H, modsgpu_view_region* out,
                               int* n_affine, int* n_out);
int  modsgpu_debug_orinet_post(modsgpu_ctx* ctx, const modsgpu_view_region* regs, const float* ori, int n, int orig_w,
                               int orig_h, const double* H, modsgpu_view_region* out, int* n_out);

/* ---- whole pair (what one iteration of mods.cpp:202-356 does for the deep configuration
 *      config_aff_ori_desc_zeromq.ini + iters_HessianZMQ.ini, vector_matcher = linear):
 *      SynthDetectDescribeKeypoints x 2 -> MatchFlannFGINN -> DuplicateFiltering -> LORANSACFiltering.
 *      Implemented by the C++ host mirror of the reference operators (csrc/host/mods_host.h). ---------- */
typedef struct {
  int    keypoints[2];          /* raw Hessian keypoints per image                     */
  int    regions[2];            /* after the AffNet eigen-ratio / border filters       */
  int    descriptors[2];        /* after ReprojectRegions (described regions)          */
  int    tentatives, unique_tentatives, inliers;
  double H[9];                  /* row-major, image 1 -> image 2 (matching.cpp:767-784) */
} modsgpu_pair_result;
/* images already resident on the device */
int  modsgpu_pair_pipeline_images(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, unsigned long long seed,
                                  modsgpu_pair_result* res, double* inlier_xy /* capacity x (x1,y1,x2,y2) or NULL */,
                                  int capacity);
/* the same pair loop for the CLASSIC configuration (config_affori_classic.ini + iters_HessianSIFT.ini, BASELINE config 1):
 * Hessian-Affine with Baumberg -> dominant orientation -> RootSIFT -> FGINN -> duplicate filter -> LO-RANSAC(H) */
int  modsgpu_pair_pipeline_classic_images(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, unsigned long long seed,
                                          modsgpu_pair_result* res, double* inlier_xy, int capacity);
/* host BGR images (cv::imread layout): upload + pipeline */
int  modsgpu_pair_pipeline(modsgpu_ctx* ctx, const uint8_t* bgr1, const uint8_t* bgr2, int w, int h,
                           unsigned long long seed, modsgpu_pair_result* res, double* inlier_xy, int capacity);
/* ---- the parameter block of the pair-level entry points: what the reference reads from its ini files into
 *      PyramidParams (structures.hpp:114-151, [HessianAffine]), MatchPars (matching.hpp:97-130, [Matching] + the
 *      FGINNThreshold of the iters file) and RANSACPars (matching.hpp:132-164, [RANSAC]).  modsgpu_default_pipeline_params
 *      fills the values of build/config_aff_ori_desc_zeromq.ini + iters_HessianZMQ.ini; the entry points above without a
 *      parameter block use exactly those. ------------------------------------------------------------------------------ */
typedef struct {
  modsgpu_pyr_params pyr;          /* [HessianAffine] */
  double mrSize;                   /* 5.1962 : patch extent of AffNet / OriNet / descriptor */
  int    patchSize, _pad0;         /* 32 */
  /* MatchPars */
  double fginn_threshold;          /* 0.8  FGINNThreshold (iters file, per descriptor) */
  double contrad_dist;             /* 10   contradDist */
  double dup_filter_radius;        /* 2    DuplicateFiltering radius (mods.cpp:283) */
  int    nn, _pad1;                /* 50   neighbours per query */
  /* RANSACPars */
  double err_threshold;            /* 4    pixels (degensac gets its square) */
  double confidence;               /* 0.99 */
  double HLAFCoef, LAFCoef;        /* 12, 2 */
  int    max_samples;              /* 1e6 */
  int    do_symm_check;            /* 1 */
  int    error_type;               /* MODSGPU_ERR_SAMPSON */
  int    just_mark_outliers;       /* 0 */
  int    use_F, _pad2;             /* 0: LORANSAC (H), 1: LORANSACF */
  uint64_t seed;                   /* the reference seeds with time(NULL) */
} modsgpu_pipeline_params;
void modsgpu_default_pipeline_params(modsgpu_pipeline_params* p);
/* modsgpu_pair_pipeline_images / modsgpu_mods_pair with every threshold of the run passed by the caller */
int  modsgpu_pair_pipeline_images_ex(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const modsgpu_pipeline_params* p,
                                     modsgpu_pair_result* res, double* inlier_xy, int capacity);

/* The two images of a pair side by side, as the reference's OpenMP tasks do (mods.cpp:234-251): with pair overlap on, the
 * pair-level entry points (modsgpu_pair_pipeline*) extract image 2 on a sibling context (same device, own stream and
 * workspaces, shared nets; created on first use, destroyed with ctx) from a helper thread while the calling thread
 * extracts image 1.  Results are identical.  Off by default (MODSGPU_PAIR_OVERLAP=1 turns it on for every new context):
 * it shortens ONE pair (7.2 -> ~5 ms); a process that already keeps the GPU full with many contexts gains nothing. */
int  modsgpu_set_pair_overlap(modsgpu_ctx* ctx, int on);
int  modsgpu_get_pair_overlap(const modsgpu_ctx* ctx);
int  modsgpu_ctx_sibling(modsgpu_ctx* ctx, modsgpu_ctx** sibling);   /* borrowed: never destroy it yourself */
int  modsgpu_ctx_sibling_join(modsgpu_ctx* ctx);                      /* fold the sibling's launch count / error into ctx */

/* ---- one image -> described regions (what extract_features_batch.cpp:128-139 does per image for the deep
 *      configuration: ImageRepresentation::SynthDetectDescribeKeypoints, identity view) and the OxAff writer
 *      (ImageRepresentation::SaveRegionsMichal text mode -> saveAR_KM_format, imagerepresentation.cpp:113-126,
 *      :205-211, :1187-1213: "128\nN\n" then `x y a b c d0..d127` per region, (a b; b c) = (A A^T)^-1 / (3 sqrt3 s)^2). */
typedef struct {
  double x, y, s, a11, a12, a21, a22;   /* reproj_kp (== det_kp for the identity view) */
  double response;
  int    octave, type;
  int    view, _pad;                    /* index of the synthesised view the region was detected in */
  float  desc[128];
} modsgpu_feature;
/* *out is malloc()ed by the library; release with modsgpu_free() */
int  modsgpu_extract_features(modsgpu_ctx* ctx, modsgpu_image* img, modsgpu_feature** out, int* n);
int  modsgpu_write_oxaff(const char* path, const modsgpu_feature* f, int n);
/* the two other region formats of extract_features_batch.cpp:147-155:
 *   text (ImageRepresentation::SaveRegions imagerepresentation.cpp:1219-1255 + saveAR :196-203):
 *        "1\nHessianAffine 1\nZMQ N\n128\n" then `x y s a11 a12 a21 a22 128 d0 .. d127 ` per region
 *   npz  (SaveRegionsNPZ :1257-1316 via cnpy): xy [N,2], scales [N,1], responses [N,1], A [N,4] float64, descs [N,128] uint8 */
int  modsgpu_write_regions_text(const char* path, const modsgpu_feature* f, int n);
int  modsgpu_write_regions_npz(const char* path, const modsgpu_feature* f, int n);

/* ---- one image over a list of synthesised views (ImageRepresentation::SynthDetectDescribeKeypoints,
 *      imagerepresentation.cpp:686-1104, HessianAffine + AffNet + OriNet + HardNet++): each view is generated on the
 *      device, detected and described in view coordinates, reprojected by H^-1 (ReprojectByH synth-detection.cpp:578-587)
 *      and filtered against the ORIGINAL image frame.  modsgpu_view_schedule = SetVSPars (synth-detection.cpp:191-322)
 *      for one iteration with an empty history: returns the number of views (<= cap written). */
typedef struct { double tilt, phi, zoom, InitSigma; int doBlur, _pad; } modsgpu_view;   /* ViewSynthParameters structures.hpp:196-209 */
int  modsgpu_view_schedule(const double* scale_set, int n_scales, const double* tilt_set, int n_tilts, double phi_base,
                           double InitSigma, int doBlur, modsgpu_view* out, int cap);
int  modsgpu_extract_features_views(modsgpu_ctx* ctx, modsgpu_image* img, const modsgpu_view* views, int n_views,
                                    modsgpu_feature** out, int* n);

/* ---- MODS run over an iteration schedule (the main loop of mods.cpp:202-356 for its HessianAffine steps, e.g.
 *      steps 2 and 3 of build/iters_MODS_ZMQ.ini): step k synthesises the views SetVSPars yields for it (views of
 *      earlier steps are not repeated), adds their regions to both images, matches all accumulated regions (FGINN),
 *      filters duplicates and verifies with LO-RANSAC -- homography (use_F = 0, mods.cpp LORANSAC) or fundamental
 *      matrix (use_F = 1, LORANSACF); stops at the first step with >= min_matches verified correspondences.
 *      model: H row-major image 1 -> 2 (use_F = 0) or F in the degensac layout (use_F = 1). ---------------------- */
typedef struct {
  double scale_set[8]; int n_scales;     /* ScaleSet */
  double tilt_set[8];  int n_tilts;      /* TiltSet  */
  double phi;                            /* Phi (rotation density, degrees; negative = vertical tilt) */
  double init_sigma;                     /* initSigma */
  double fginn_threshold;                /* FGINNThreshold */
  int    do_blur, _pad;
} modsgpu_mods_step;
typedef struct {
  int    steps_done, views[2], regions[2], tentatives, unique_tentatives, inliers;
  double model[9];
} modsgpu_mods_result;
int  modsgpu_mods_pair(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const modsgpu_mods_step* steps,
                       int n_steps, int min_matches, int use_F, unsigned long long seed, modsgpu_mods_result* res,
                       double* inlier_xy, int capacity);
/* the same with the run's parameter block (detector, matcher and RANSAC settings; p->use_F selects LORANSACF) */
int  modsgpu_mods_pair_ex(modsgpu_ctx* ctx, modsgpu_image* img1, modsgpu_image* img2, const modsgpu_mods_step* steps, int n_steps,
                          int min_matches, const modsgpu_pipeline_params* p, modsgpu_mods_result* res, double* inlier_xy, int capacity);

/* ---- pre-extracted regions (`read_pre_extracted`, mods.cpp:216-229): the reference re-loads region files instead of
 *      detecting, then matches them.  Readers (host only, *out malloc()ed -> modsgpu_free):
 *   npz  = ImageRepresentation::PreLoadRegionsNPZ / LoadRegionsNPZ (imagerepresentation.cpp:1355-1512): members xy [N,2],
 *          scales [N], responses [N], descs [N,D<=128] (taken as uchar) + A [N,4] | angles [N] (degrees) | neither
 *          (upright);  type = DET_READ (structures.hpp:20).  Any numeric dtype is converted like cnpy's typed views
 *          would be for the dtypes SaveRegionsNPZ writes (float64 / uint8).
 *   text = ImageRepresentation::LoadRegions (:1317-1354) with loadAR / loadKP (:237-253).  NB the reference's own
 *          SaveRegions (:1219) writes saveAR records, which LoadRegions cannot read back (reference quirk, kept);
 *          this reader follows LoadRegions.
 *   modsgpu_match_features = the matching half of a mods.cpp step on two region lists: MatchFlannFGINN (linear) ->
 *          DuplicateFiltering -> LORANSACFiltering (H, or F when use_F).  res->views are 0. */
int  modsgpu_read_regions_npz(const char* path, modsgpu_feature** out, int* n);
int  modsgpu_read_regions_text(const char* path, modsgpu_feature** out, int* n);
int  modsgpu_match_features(modsgpu_ctx* ctx, const modsgpu_feature* f1, int n1, const modsgpu_feature* f2, int n2,
                            int desc_dim, double fginn_threshold, int use_F, unsigned long long seed,
                            modsgpu_mods_result* res, double* inlier_xy, int capacity);

/* ---- CorrespondenceBank::MatchImgReps (correspondencebank.cpp:234-343) over region lists filed per (detector, descriptor):
 *      GROUPED -- for every group descriptor the lists of all group detectors are pooled (image 2 = train, image 1 = query)
 *      and matched once (FGINN, that descriptor's threshold), filed under ("Group", descriptor); SEPARATE -- every separate
 *      detector x separate descriptor pair is matched on its own.  A descriptor without a threshold (or <= 0) is skipped.
 *      Name sets / thresholds are comma-separated ("HessianAffine,MSER", "ZMQ=0.8,RootSIFT=0.85").  out: 7 doubles per
 *      tentative (x1 y1 x2 y2 d1 d2 ratio) in the order of CorrespondenceBank::GetCorresponcesVector("All", "All"). -------- */
typedef struct { int image /* 1 | 2 */, n; const char* det; const char* desc; const modsgpu_feature* f; } modsgpu_region_list;
int  modsgpu_match_imgreps(modsgpu_ctx* ctx, const modsgpu_region_list* lists, int n_lists, const char* group_detectors,
                           const char* group_descriptors, const char* separate_detectors, const char* separate_descriptors,
                           const char* fginn_thresholds, double* out, int capacity, int* n_out);

/* DuplicateFiltering + LORANSACFiltering (the second half of modsgpu_match_features) on tentatives matched elsewhere, e.g. by
 * several GPUs that each ran modsgpu_match_fginn on a slice of the query rows (config 5, mods_dist.py).  m in query order. */
int  modsgpu_verify_matches(modsgpu_ctx* ctx, const modsgpu_feature* f1, int n1, const modsgpu_feature* f2, int n2,
                            const modsgpu_match* m, int nm, int use_F, unsigned long long seed, modsgpu_mods_result* res,
                            double* inlier_xy, int capacity);

/* test-only: one 128x32x64 GEMM through the tcgen05 descriptor conventions of the dense kernels */
int  modsgpu_debug_umma_probe(modsgpu_ctx* ctx, const float* A, const float* B, float* D, int swap_lbo_sbo);
/* test / profiling only: cycles per tcgen05.mma (M128 K16) for a given operand-descriptor configuration;
 * cfg = {n, a_off_bytes, a_lbo, a_sbo, b_lbo, b_sbo, layout_type, reps, n_accumulators, a_step_bytes, grid} */
int  modsgpu_debug_umma_pace(modsgpu_ctx* ctx, const int* cfg, double* cycles_per_mma);

/* ---- measurement helpers used by bench.py ---------------------------------------------------------------- */
/* per-launch CUDA-event timing of every kernel (aggregated by kernel name); report is a JSON object */
int  modsgpu_profile_enable(modsgpu_ctx* ctx, int on);
int  modsgpu_profile_report(modsgpu_ctx* ctx, char* buf, int cap);
/* CUDA events on the ctx stream around an arbitrary host-side region */
int  modsgpu_timer_start(modsgpu_ctx* ctx);
int  modsgpu_timer_stop(modsgpu_ctx* ctx, float* ms);
/* overwrite a 256 MB scratch buffer (> the 126 MB L2) on the ctx stream */
int  modsgpu_flush_l2(modsgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
