/*
 * mods_oracle.h -- CPU restatement of the MODS hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the parity ORACLE for the B200 kernels.  It is never linked
 * into, imported by or called from the product library (libmodsgpu.so); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * Every function cites the reference file:line (relative to the upstream tree,
 * ducha-aiki/mods-light-zmq @ 33c9ba2) it restates.  OpenCV calls whose
 * arithmetic lives outside the reference tree (cv::GaussianBlur, cv::resize)
 * are restated op-for-op and PINNED bit-exactly against cv2 4.13.0
 * (tests/test_oracle_cv2_pin.py + tests/golden/).
 */
#ifndef MODS_ORACLE_H
#define MODS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int   numberOfScales;       /* structures.hpp:119  (3)    */
  float initialSigma;         /* structures.hpp:121  (1.6)  */
  float threshold;            /* structures.hpp:123  (5.33 from the ini) */
  double edgeEigenValueRatio; /* structures.hpp:128  (10)   */
  int   border;               /* structures.hpp:130  (5)    */
} orc_pyr_params;

/* One detected keypoint, in the order the reference exports them
 * (sorted by |response| descending, scale-space-detector.hpp:120-131). */
typedef struct {
  float x, y, s;      /* image coordinates / scale, pyramid.cpp:401-402 */
  float response;     /* localized peak value                            */
  int   type;         /* 0 dark, 1 bright, 2 saddle (pyramid.h:30-34)    */
  int   octave;       /* 0,1,2..                                          */
  int   level;        /* 1..numberOfScales: index of `cur` in the octave  */
  int   r0, c0;       /* raster position of the NMS candidate             */
  int   r, c;         /* final integer position after localisation        */
  int   seq;          /* detection sequence number (push order)           */
} orc_keypoint;

/* Region handed to the patch sampler / describers (AffineKeypoint, structures.hpp:185-194) */
typedef struct {
  double x, y, s;
  double a11, a12, a21, a22;
} orc_region;

/* ---- image primitives ---------------------------------------------------- */
/* gray = (B+G+R)/3.0 as cv::Mat expr (synth-detection.cpp:344-351) */
void orc_gray_from_bgr(const uint8_t* bgr, int w, int h, float* gray);
/* helpers.cpp:717-731 gaussianBlur(): cv::GaussianBlur, ksize=(int)(6s+1)|1, BORDER_REPLICATE */
void orc_gaussian_blur(const float* in, float* out, int w, int h, float sigma);
int  orc_gaussian_kernel(float sigma, float* taps /* >= 6*sigma+3 floats */);
/* pyramid.cpp:196-254 HessianResponse (borders := 0) */
void orc_hessian_response(const float* in, float* out, int w, int h, float norm);
/* pyramid.cpp:476 cv::resize(.., 0.5, 0.5, INTER_LINEAR) */
void orc_half_size(int w, int h, int* ow, int* oh);
void orc_half_image(const float* in, int w, int h, float* out);

/* ---- classic stages (config_affori_classic.ini): dominant orientation and (Root)SIFT ---- */
/* helpers.cpp:30-72 ATAN_LUT as doubles (round(atan(i/255),1e-10) with the reference's 3 typo entries) */
void orc_atan_lut(double* lut256);
/* helpers.cpp:442-459 computeCircularGaussMask (sigma == 0 -> 0.9*r^2) */
void orc_circular_gauss_mask(float* mask, int size, float sigma);
/* synth-detection.cpp:1039-1149 DetectOrientation + :836-929 EstimateDominantAnglesFunctor (doHalfSIFT = 0,
 * addUpRight = false).  n_ang[i] = -1: region dropped by the k_sigma*s frame test; else the number of angles
 * (<= maxAngles, taken in bin order -- SURVEY Q13) written to angles[i*maxAngles ..]. */
void orc_dominant_orientation(const float* img, int w, int h, const orc_region* regs, int n, double mrSize, int patchSize,
                              int maxAngles, double th, int* n_ang, float* angles);
/* synth-detection.hpp:170-263 DescribeRegions<SIFTDescriptor> (FastPatchExtraction = false) + siftdesc.cpp:
 * patchSize x patchSize float patch (3-step sampler), optional photometricallyNormalize (helpers.cpp:666-716),
 * gradients, 4x4x8 trilinear histogram with the circular Gauss mask, (Root)SIFT normalisation -> 128 integers 0..255 */
void orc_describe_sift(const float* img, int w, int h, const orc_region* regs, int n, double mrSize, int patchSize,
                       int photoNorm, int rootSift, float* desc);

/* ---- view synthesis (synth-detection.cpp:324-518 GenerateSynthImageCorr) ---- */
/* cv::warpAffine(src, dst, M (2x3 double, forward map), Size(ow,oh), INTER_LINEAR, BORDER_CONSTANT, border) on CV_32F */
void orc_warp_affine(const float* in, int w, int h, const double* M, float* out, int ow, int oh, float border);
/* cv::GaussianBlur(img, img, Size(kx,ky), sigma_x, sigma_y) with the default BORDER_REFLECT_101 */
void orc_gaussian_blur_xy(const float* in, float* out, int w, int h, int kx, int ky, double sigma_x, double sigma_y);
/* output size + H (row-major 3x3) of GenerateSynthImageCorr; returns 1 for the identity view (pixels = input) */
int  orc_synth_geometry(int w, int h, double tilt, double phi, double zoom, int* ow, int* oh, double* H);
/* the whole function; out must hold ow*oh floats (from orc_synth_geometry) */
void orc_synth_view(const float* gray, int w, int h, double tilt, double phi, double zoom, double InitSigma, int doBlur,
                    float* out);

/* ---- detector (pyramid.cpp:428-529, scale-space-detector.hpp:47-198) ------ */
int orc_detect_hessian(const float* gray, int w, int h, const orc_pyr_params* p,
                       orc_keypoint* out, int cap);
/* synth-detection.hpp:79-112 glue for doBaumberg=0: region = (x,y,s,I) */

/* AffineShapeParams (affine.h:26-68) read by the in-pyramid Baumberg iteration */
typedef struct {
  int   maxIterations;         /* 16   */
  float convergenceThreshold;  /* 0.05 */
  int   smmWindowSize;         /* 19   */
  float initialSigma;          /* 1.6  */
} orc_affshape_params;
/* the same detector with doBaumberg = 1, method SMM (affine.cpp:26-158, called from localizeKeypoint on prevBlur,
 * pyramid.cpp:402): keypoints whose iteration does not converge are dropped; A (4 floats per kept
 * keypoint: a11 a12 a21 a22, as AffineShape hands them to onAffineShapeFound) */
int orc_detect_hessian_affine(const float* gray, int w, int h, const orc_pyr_params* p, const orc_affshape_params* a,
                              orc_keypoint* out, float* A, int cap);
/* helpers.cpp:413-440 computeGaussMask */
void orc_gauss_mask(float* mask, int size);

/* ---- patch sampler (synth-detection.cpp:38-132, helpers.cpp:524-626) ------ */
int  orc_interpolate_check_borders(int w, int h, float ofsx, float ofsy, float a11, float a12,
                                   float a21, float a22, int res_w, int res_h);
int  orc_interpolate(const float* im, int w, int h, float ofsx, float ofsy, float a11, float a12,
                     float a21, float a22, float* res, int res_w, int res_h);
void orc_extract_patches(const float* img, int w, int h, const orc_region* regs, int n,
                         double mrSize, int patchSize, float* out /* n*ps*ps */);
/* cv::imencode(".png", CV_32F) -> 8 bit: saturate_cast<uchar>(float) = round-half-even, clamp
 * (imagerepresentation.cpp:45) */
void orc_quantize_u8(const float* in, uint8_t* out, long n);

/* ---- region filters -------------------------------------------------------- */
/* imagerepresentation.cpp:803-845: AffNet output -> A, eig-ratio and border filters.
 * in: n regions (x,y,s) + n x 3 AffNet outputs; out: surviving regions, keep[i]=index or -1 */
int orc_affnet_postprocess(const orc_region* in, const float* aff3, int n, int w, int h,
                           double mrSize, orc_region* out, int* src_index);
/* imagerepresentation.cpp:881-899: OriNet output -> A*R(angle) */
void orc_orinet_postprocess(const orc_region* in, const float* ori2, int n, orc_region* out);
/* synth-detection.cpp:631-706 ReprojectRegions with H = I: centre inside + 2*3*sqrt(3)*s frame */
int orc_reproject_filter(const orc_region* in, int n, int w, int h, orc_region* out, int* src_index);

/* ---- matching (matching.cpp:356-460 with vector_matcher=linear) ----------- */
typedef struct {
  int   qi, ti, tj_bad;   /* query idx, 1st NN idx, first ratio-passing NN idx */
  float d1, d2;           /* squared L2 to 1st NN, to the ratio-passing NN     */
  double ratio;           /* sqrt(d1/d2) (float division, then double sqrt)    */
} orc_match;
/* top-nn (dist asc, index asc) exact linear k-NN, cvflann LinearIndex + KNNSimpleResultSet */
void orc_knn_linear(const float* q, int nq, const float* t, int nt, int dim, int nn,
                    int* idx /* nq*nn */, float* dist /* nq*nn */);
int orc_match_fginn(const float* q, const double* qxy, int nq, const float* t, const double* txy,
                    int nt, int dim, double ratio_thr, double contrad_dist, int nn,
                    orc_match* out /* cap nq */);
/* matching.cpp:2615-2679 DuplicateFiltering mode bestFGINN, made deterministic by a stable sort.
 * xy1/xy2: T x 2 doubles; order_out: indices of survivors in sorted order. returns count. */
int orc_duplicate_filter(const double* xy1, const double* xy2, const double* ratio, int T,
                         double r, int* order_out);

#ifdef __cplusplus
}
#endif
#endif
