"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (libmodsgpu.so + its Python binding) never does.

liboracle.so      : our restatement (oracle/mods_oracle.cpp)
_ref/libdegensac_ref.so : the reference's own degensac C code (oracle/Makefile `ref`)
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class PyrParams(C.Structure):
    _fields_ = [("numberOfScales", C.c_int), ("initialSigma", C.c_float), ("threshold", C.c_float),
                ("edgeEigenValueRatio", C.c_double), ("border", C.c_int)]


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("s", C.c_float), ("response", C.c_float),
                ("type", C.c_int), ("octave", C.c_int), ("level", C.c_int), ("r0", C.c_int), ("c0", C.c_int),
                ("r", C.c_int), ("c", C.c_int), ("seq", C.c_int)]


KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("s", "f4"), ("response", "f4"), ("type", "i4"), ("octave", "i4"),
                     ("level", "i4"), ("r0", "i4"), ("c0", "i4"), ("r", "i4"), ("c", "i4"), ("seq", "i4")])
REGION_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("s", "f8"), ("a11", "f8"), ("a12", "f8"), ("a21", "f8"), ("a22", "f8")])
MATCH_DTYPE = np.dtype([("qi", "i4"), ("ti", "i4"), ("tj_bad", "i4"), ("d1", "f4"), ("d2", "f4"), ("_pad", "i4"), ("ratio", "f8")])

_lib = None


def default_params():
    """[HessianAffine] section of build/config_aff_ori_desc_zeromq.ini:41-55."""
    return PyrParams(3, 1.6, 5.33, 10.0, 5)


def build():
    subprocess.check_call(["make", "-C", HERE, "-s", "all"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        assert MATCH_DTYPE.itemsize == 32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gray_from_bgr(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.float32)
    lib().orc_gray_from_bgr(_p(bgr), w, h, _p(out))
    return out


def gaussian_blur(img, sigma):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur(_p(img), _p(out), w, h, C.c_float(sigma))
    return out


class AffShapeParams(C.Structure):
    _fields_ = [("maxIterations", C.c_int), ("convergenceThreshold", C.c_float), ("smmWindowSize", C.c_int),
                ("initialSigma", C.c_float)]


def default_affshape_params():
    """[HessianAffine] of build/config_affori_classic.ini:42-49"""
    return AffShapeParams(16, 0.05, 19, 1.6)


def detect_hessian_affine(gray, params=None, aff=None, cap=1 << 18):
    """Hessian-Affine with the in-pyramid Baumberg iteration: (keypoints, A [n x 4])."""
    gray = np.ascontiguousarray(gray, np.float32)
    h, w = gray.shape
    params = params or default_params()
    aff = aff or default_affshape_params()
    out = np.zeros(cap, KP_DTYPE)
    A = np.zeros((cap, 4), np.float32)
    n = lib().orc_detect_hessian_affine(_p(gray), w, h, C.byref(params), C.byref(aff), _p(out), _p(A), cap)
    assert n <= cap
    return out[:n].copy(), A[:n].copy()


def dominant_orientation(img, regs, mr_size=5.1962, patch_size=32, max_angles=1, th=0.8):
    """DetectOrientation (synth-detection.cpp:1039-1149): returns (n_ang [n] with -1 = dropped, angles [n x max_angles])."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    regs = np.ascontiguousarray(regs, REGION_DTYPE)
    n_ang = np.zeros(len(regs), np.int32)
    ang = np.zeros((len(regs), max(max_angles, 1)), np.float32)
    lib().orc_dominant_orientation(_p(img), w, h, _p(regs), len(regs), C.c_double(mr_size), patch_size, max_angles,
                                   C.c_double(th), _p(n_ang), _p(ang))
    return n_ang, ang


def apply_orientations(regs, n_ang, ang):
    """The region list DetectOrientation returns (addUpRight = false): one rotated copy per accepted angle."""
    out = []
    for r, n, a in zip(regs, n_ang, ang):
        for j in range(max(n, 0)):
            ci, si = np.cos(-float(a[j])), np.sin(-float(a[j]))
            t = r.copy()
            t["a11"] = r["a11"] * ci - r["a12"] * si
            t["a12"] = r["a11"] * si + r["a12"] * ci
            t["a21"] = r["a21"] * ci - r["a22"] * si
            t["a22"] = r["a21"] * si + r["a22"] * ci
            out.append(t)
    return np.array(out, REGION_DTYPE) if out else np.zeros(0, REGION_DTYPE)


def describe_sift(img, regs, mr_size=5.1962, patch_size=41, photo_norm=1, root_sift=1):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    regs = np.ascontiguousarray(regs, REGION_DTYPE)
    out = np.zeros((len(regs), 128), np.float32)
    lib().orc_describe_sift(_p(img), w, h, _p(regs), len(regs), C.c_double(mr_size), patch_size, int(photo_norm), int(root_sift), _p(out))
    return out


def atan_lut():
    t = np.zeros(256, np.float64)
    lib().orc_atan_lut(_p(t))
    return t


def warp_affine(img, M, ow, oh, border=128.0):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    out = np.empty((oh, ow), np.float32)
    lib().orc_warp_affine(_p(img), w, h, _p(M), _p(out), ow, oh, C.c_float(border))
    return out


def gaussian_blur_xy(img, kx, ky, sigma_x, sigma_y):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur_xy(_p(img), _p(out), w, h, kx, ky, C.c_double(sigma_x), C.c_double(sigma_y))
    return out


def synth_geometry(w, h, tilt, phi, zoom):
    """(ow, oh, H[3x3], identity?) of GenerateSynthImageCorr (synth-detection.cpp:356-431)."""
    ow, oh = C.c_int(), C.c_int()
    H = np.zeros(9, np.float64)
    ident = lib().orc_synth_geometry(w, h, C.c_double(tilt), C.c_double(phi), C.c_double(zoom), C.byref(ow), C.byref(oh), _p(H))
    return ow.value, oh.value, H.reshape(3, 3), bool(ident)


def synth_view(gray, tilt, phi, zoom, init_sigma, do_blur=1):
    gray = np.ascontiguousarray(gray, np.float32)
    h, w = gray.shape
    ow, oh, H, _ = synth_geometry(w, h, tilt, phi, zoom)
    out = np.empty((oh, ow), np.float32)
    lib().orc_synth_view(_p(gray), w, h, C.c_double(tilt), C.c_double(phi), C.c_double(zoom), C.c_double(init_sigma),
                         int(do_blur), _p(out))
    return out, H


def gaussian_kernel(sigma):
    k = np.zeros(4096, np.float32)
    n = lib().orc_gaussian_kernel(C.c_float(sigma), _p(k))
    return k[:n].copy()


def hessian_response(img, norm):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_hessian_response(_p(img), _p(out), w, h, C.c_float(norm))
    return out


def half_image(img):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    ow, oh = C.c_int(), C.c_int()
    lib().orc_half_size(w, h, C.byref(ow), C.byref(oh))
    out = np.empty((oh.value, ow.value), np.float32)
    lib().orc_half_image(_p(img), w, h, _p(out))
    return out


def detect_hessian(gray, params=None, cap=200000):
    gray = np.ascontiguousarray(gray, np.float32)
    h, w = gray.shape
    params = params or default_params()
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_detect_hessian(_p(gray), w, h, C.byref(params), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


FIXED_TH, RELATIVE_TH, FIXED_REG_NUMBER, RELATIVE_REG_NUMBER, NOT_LESS_THAN_REGIONS = range(5)


def detect_hessian_mode(gray, mode, threshold=5.33, rel_threshold=-1.0, reg_number=-1, rel_reg_number=-1.0, affine=False):
    """The detection modes of prepareKeysForExport (scale-space-detector.hpp:125-198): every mode but FIXED_TH runs the
    pyramid with all thresholds at 0 (pyramid.h:58-59) and truncates the |response|-descending key list.  Numpy
    restatement of the truncation rules on top of the C oracle's detector."""
    par = default_params()
    par.threshold = threshold if mode == FIXED_TH else 0.0
    if affine:
        kps, A = detect_hessian_affine(gray, par, cap=1 << 20)
    else:
        kps, A = detect_hessian(gray, par, cap=1 << 20), None
    n = len(kps)
    if n == 0 or mode == FIXED_TH:
        return (kps, A) if affine else kps
    mag = np.abs(kps["response"].astype(np.float64))
    keep = n
    if mode == RELATIVE_TH:
        eff = np.float32(mag[0] * np.float64(np.float32(rel_threshold)))       # float member effectiveThreshold (:145)
        keep = int(np.sum(mag > abs(float(eff))))                             # lower_bound on the sorted list (:149-150)
    elif mode == FIXED_REG_NUMBER:
        nr = int(np.floor(3.0 * reg_number)) if affine else reg_number          # doBaumberg triples it (:156-157) ...
        if 0 <= nr < n:
            keep = nr
        keep = min(keep, reg_number)                                            # ... and :194-195 cuts it back
    elif mode == RELATIVE_REG_NUMBER:
        keep = int(np.floor(np.float64(np.float32(rel_reg_number)) * n))
    elif mode == NOT_LESS_THAN_REGIONS:
        fixed = int(np.sum(mag > float(np.float32(threshold))))                 # un-squared threshold (:174)
        keep = min(reg_number, n) if fixed < reg_number else min(fixed, n)
    keep = max(0, min(keep, n))
    return (kps[:keep], A[:keep]) if affine else kps[:keep]


def regions_from_keypoints(kps):
    """synth-detection.hpp:79-112 with doBaumberg=0: A = I, s unchanged (sqrt|det I| = 1)."""
    r = np.zeros(len(kps), REGION_DTYPE)
    r["x"], r["y"], r["s"] = kps["x"], kps["y"], kps["s"]
    r["a11"] = 1.0
    r["a22"] = 1.0
    return r


def interpolate(img, ofsx, ofsy, a11, a12, a21, a22, res_w, res_h):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((res_h, res_w), np.float32)
    f = C.c_float
    t = lib().orc_interpolate(_p(img), w, h, f(ofsx), f(ofsy), f(a11), f(a12), f(a21), f(a22), _p(out), res_w, res_h)
    return out, bool(t)


def extract_patches(img, regions, mrSize=5.1962, patchSize=32):
    img = np.ascontiguousarray(img, np.float32)
    regions = np.ascontiguousarray(regions, REGION_DTYPE)
    h, w = img.shape
    out = np.empty((len(regions), patchSize, patchSize), np.float32)
    lib().orc_extract_patches(_p(img), w, h, _p(regions), len(regions), C.c_double(mrSize), patchSize, _p(out))
    return out


def quantize_u8(a):
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty(a.shape, np.uint8)
    lib().orc_quantize_u8(_p(a), _p(out), C.c_long(a.size))
    return out


def affnet_postprocess(regions, aff3, w, h, mrSize=5.1962):
    regions = np.ascontiguousarray(regions, REGION_DTYPE)
    aff3 = np.ascontiguousarray(aff3, np.float32)
    out = np.zeros(len(regions), REGION_DTYPE)
    src = np.zeros(len(regions), np.int32)
    m = lib().orc_affnet_postprocess(_p(regions), _p(aff3), len(regions), w, h, C.c_double(mrSize), _p(out), _p(src))
    return out[:m].copy(), src[:m].copy()


def orinet_postprocess(regions, ori2):
    regions = np.ascontiguousarray(regions, REGION_DTYPE)
    ori2 = np.ascontiguousarray(ori2, np.float32)
    out = np.zeros(len(regions), REGION_DTYPE)
    lib().orc_orinet_postprocess(_p(regions), _p(ori2), len(regions), _p(out))
    return out


def reproject_filter(regions, w, h):
    regions = np.ascontiguousarray(regions, REGION_DTYPE)
    out = np.zeros(len(regions), REGION_DTYPE)
    src = np.zeros(len(regions), np.int32)
    m = lib().orc_reproject_filter(_p(regions), len(regions), w, h, _p(out), _p(src))
    return out[:m].copy(), src[:m].copy()


def knn_linear(q, t, nn=50):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    idx = np.empty((len(q), nn), np.int32)
    dist = np.empty((len(q), nn), np.float32)
    lib().orc_knn_linear(_p(q), len(q), _p(t), len(t), q.shape[1], nn, _p(idx), _p(dist))
    return idx, dist


def match_fginn(q, qxy, t, txy, ratio=0.8, contrad=10.0, nn=50):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    qxy = np.ascontiguousarray(qxy, np.float64)
    txy = np.ascontiguousarray(txy, np.float64)
    out = np.zeros(max(len(q), 1), MATCH_DTYPE)
    m = lib().orc_match_fginn(_p(q), _p(qxy), len(q), _p(t), _p(txy), len(t), q.shape[1] if len(q) else 128,
                              C.c_double(ratio), C.c_double(contrad), nn, _p(out))
    return out[:m].copy()


def duplicate_filter(xy1, xy2, ratio, r=2.0):
    xy1 = np.ascontiguousarray(xy1, np.float64)
    xy2 = np.ascontiguousarray(xy2, np.float64)
    ratio = np.ascontiguousarray(ratio, np.float64)
    T = len(ratio)
    order = np.zeros(max(T, 1), np.int32)
    m = lib().orc_duplicate_filter(_p(xy1), _p(xy2), _p(ratio), T, C.c_double(r), _p(order))
    return order[:m].copy()


# --------------------------------------------------------------------------- reference degensac
_ref = None


class Score(C.Structure):
    _fields_ = [("I", C.c_uint), ("J", C.c_double)]


def _load_lapack():
    """dsyev_/dgesvd_ come from the OpenBLAS bundled with the cv2 wheel of this image
    (SURVEY §8c.1).  Loaded RTLD_GLOBAL so libdegensac_ref.so resolves against it."""
    import importlib.util
    try:
        import cv2  # noqa: F401  -- pulls in the wheel's bundled libgfortran/libquadmath in the right order
    except ImportError:
        pass
    spec = importlib.util.find_spec("cv2")
    base = os.path.dirname(os.path.dirname(spec.origin))
    cands = glob.glob(os.path.join(base, "opencv_python_headless.libs", "libopenblas*.so*")) + \
        glob.glob(os.path.join(base, "opencv_python.libs", "libopenblas*.so*")) + \
        glob.glob(os.path.join(base, "scipy.libs", "libscipy_openblas*.so*"))
    # gfortran runtime deps of that OpenBLAS live in the same directory
    for c in cands:
        d = os.path.dirname(c)
        for dep in sorted(glob.glob(os.path.join(d, "libquadmath*.so*"))) + sorted(glob.glob(os.path.join(d, "libgfortran*.so*"))):
            try:
                C.CDLL(dep, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
        try:
            return C.CDLL(c, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    raise OSError("no LAPACK (OpenBLAS) found for the reference degensac")


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libdegensac_ref.so"))


def ref():
    global _ref
    if _ref is None:
        _load_lapack()
        _ref = C.CDLL(os.path.join(HERE, "_ref", "libdegensac_ref.so"))
        _ref.exp_ransacHcustom.restype = Score
    return _ref


def ref_ransac_H(u, th=16.0, conf=0.99, max_sam=1000000, seed_time=12345, sym_check=1, error="sampson"):
    """Calls the reference's exp_ransacHcustom exactly as matching.cpp:731 does.
    u: T x 6 doubles (x1,y1,1,x2,y2,1).  Returns dict(H, inl, samples, lo_count, oc_rejects, I, J)."""
    L = ref()
    u = np.ascontiguousarray(u, np.float64)
    T = len(u)
    H = np.zeros(9, np.float64)
    inl = np.zeros(T, np.uint8)
    data_out = np.zeros(T * 18 + 8, np.int32)
    resids = C.c_void_p()
    L.orc_ref_set_time(C.c_long(seed_time))
    names = {"sampson": ("HDs", "HDsi", "HDsidx"), "symm_sum": ("HDsSym", "HDsiSym", "HDsSymidx"),
             "symm_max": ("HDsSymMax", "HDsiSymMax", "HDsSymidxMax")}[error]
    f = [C.cast(getattr(L, n), C.c_void_p) for n in names]
    if T <= 20:
        max_sam = 1000  # matching.cpp:644-645
    S = L.exp_ransacHcustom(_p(u), T, C.c_double(th), C.c_double(conf), max_sam, _p(H), _p(inl), 4, _p(data_out),
                            1, C.c_uint(0), C.byref(resids), f[0], f[1], f[2], sym_check)
    C.CDLL(None).free(resids)
    return dict(H=H, inl=inl, samples=int(data_out[0]), lo_count=int(data_out[1]), oc_rejects=int(data_out[2]),
                I=int(S.I), J=float(S.J))


def ref_ransac_F(u, th=16.0, conf=0.99, max_sam=1000000, seed_time=12345, sym_check=1):
    """Calls the reference's exp_ransacFcustom exactly as matching.cpp:722 does (do_lo = 1, inlLimit = 0,
    Sampson error functions exFDs / FDs).  Returns dict(F, inl, samples, lo_count, I, Ih)."""
    L = ref()
    u = np.ascontiguousarray(u, np.float64)
    T = len(u)
    F = np.zeros(9, np.float64)
    Hin = np.zeros(9, np.float64)
    inl = np.zeros(T, np.uint8)
    data_out = np.zeros(T * 18 + 8, np.int32)   # exp_ranF.c:1029 increments data_out[LmaxI + 2]
    resids = C.c_void_p()
    Ih = C.c_int(0)
    L.orc_ref_set_time(C.c_long(seed_time))
    if T <= 20:
        max_sam = 1000  # matching.cpp:644-645
    L.exp_ransacFcustom.restype = C.c_int
    I = L.exp_ransacFcustom(_p(u), T, C.c_double(th), C.c_double(conf), max_sam, _p(F), _p(inl), _p(data_out), 1,
                            C.c_uint(0), C.byref(resids), _p(Hin), C.byref(Ih), C.cast(L.exFDs, C.c_void_p),
                            C.cast(L.FDs, C.c_void_p), sym_check)
    C.CDLL(None).free(resids)
    return dict(F=F, inl=inl, samples=int(data_out[0]), lo_count=int(data_out[1]), I=int(I), Ih=int(Ih.value))


def sampson_F(F, u):
    """Ftools.c:83-101 FDs (numpy): squared Sampson error of every correspondence under F (degensac layout)."""
    F = np.asarray(F, np.float64).ravel()
    u = np.asarray(u, np.float64)
    rxc = F[0] * u[:, 3] + F[3] * u[:, 4] + F[6]
    ryc = F[1] * u[:, 3] + F[4] * u[:, 4] + F[7]
    rwc = F[2] * u[:, 3] + F[5] * u[:, 4] + F[8]
    r = u[:, 0] * rxc + u[:, 1] * ryc + rwc
    rx = F[0] * u[:, 0] + F[1] * u[:, 1] + F[2]
    ry = F[3] * u[:, 0] + F[4] * u[:, 1] + F[5]
    return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry)


def classic_regions(gray, mr_size=5.1962):
    """The per-image chain of config_affori_classic.ini + iters_HessianSIFT.ini (identity view), on the CPU:
    Hessian-Affine with Baumberg -> DetectAffineRegions glue (synth-detection.hpp:79-112) -> centre-inside ->
    DetectOrientation -> ReprojectRegions frame test -> RootSIFT.  Returns (n_keypoints, regions, descriptors)."""
    h, w = gray.shape
    kp, A = detect_hessian_affine(gray)
    regs = np.zeros(len(kp), REGION_DTYPE)
    a11, a12, a21, a22 = [A[:, i].astype(np.float64) for i in range(4)]
    regs["s"] = kp["s"].astype(np.float64) * np.sqrt(np.abs(a11 * a22 - a12 * a21))
    det = np.sqrt(np.abs(a11 * a22 - a12 * a21))          # rectifyTransformation, synth-detection.cpp:134-143
    b2a2 = np.sqrt(a12 * a12 + a11 * a11)
    regs["a11"], regs["a12"] = b2a2 / det, 0.0
    regs["a21"], regs["a22"] = (a22 * a12 + a21 * a11) / (b2a2 * det), det / b2a2
    regs["x"], regs["y"] = kp["x"], kp["y"]
    regs = regs[(regs["x"] < w) & (regs["y"] < h) & (regs["x"] > 0) & (regs["y"] > 0)]
    n_ang, ang = dominant_orientation(gray, regs, mr_size)
    r2 = apply_orientations(regs, n_ang, ang)
    r3, _ = reproject_filter(r2, w, h)
    return len(kp), r3, describe_sift(gray, r3, mr_size)


# --------------------------------------------------------------------------- reprojection (H != I)
def reproject_by_H(regions, Hinv):
    """ReprojectByH (synth-detection.cpp:578-587) on every region: the affine part of Hinv (row-major 3x3, view ->
    original) applied to the centre and to A; s and the other fields are copied.  numpy evaluates each product and
    sum as a separate IEEE double operation, left to right, like the reference's expression."""
    Hm = np.asarray(Hinv, np.float64).ravel()
    out = regions.copy()
    x, y = regions["x"], regions["y"]
    out["x"] = (Hm[0] * x + Hm[1] * y + Hm[2])
    out["y"] = (Hm[3] * x + Hm[4] * y + Hm[5])
    out["a11"] = (Hm[0] * regions["a11"] + Hm[1] * regions["a21"])
    out["a12"] = (Hm[0] * regions["a12"] + Hm[1] * regions["a22"])
    out["a21"] = (Hm[3] * regions["a11"] + Hm[4] * regions["a21"])
    out["a22"] = (Hm[3] * regions["a12"] + Hm[4] * regions["a22"])
    return out


def invert3(H):
    """cofactor inverse, the operation order of the host mirror's invert3h (cv::invert's arithmetic is third-party)"""
    A = np.asarray(H, np.float64).ravel()
    c0 = A[4] * A[8] - A[5] * A[7]
    c1 = A[5] * A[6] - A[3] * A[8]
    c2 = A[3] * A[7] - A[4] * A[6]
    det = A[0] * c0 + A[1] * c1 + A[2] * c2
    i = 1.0 / det
    return np.array([c0 * i, (A[2] * A[7] - A[1] * A[8]) * i, (A[1] * A[5] - A[2] * A[4]) * i,
                     c1 * i, (A[0] * A[8] - A[2] * A[6]) * i, (A[2] * A[3] - A[0] * A[5]) * i,
                     c2 * i, (A[1] * A[6] - A[0] * A[7]) * i, (A[0] * A[4] - A[1] * A[3]) * i])


def centre_inside(regions, w, h):
    """ReprojectRegionsAndRemoveTouchBoundary with dontRemove (synth-detection.cpp:151-190): centre strictly inside"""
    return (regions["x"] < w) & (regions["y"] < h) & (regions["x"] > 0) & (regions["y"] > 0)


def describe_view_chain(gray_view, H, orig_w, orig_h, nets, mrSize=5.1962):
    """The per-view chain of imagerepresentation.cpp:704-1006 on the CPU: detect -> AffNet + tests -> centre test ->
    OriNet + rotation -> ReprojectRegions -> HardNet++.  `nets` = (affnet, orinet, hardnet) callables on u8 patches.
    Returns dict(det, reproj, desc, counts)."""
    h, w = gray_view.shape
    eye = H is None
    Hinv = None if eye else invert3(H)
    kp = detect_hessian(gray_view)
    regs = regions_from_keypoints(kp)
    aff = nets[0](quantize_u8(extract_patches(gray_view, regs, mrSize)))
    r2, _ = affnet_postprocess(regs, aff, w, h, mrSize)
    n_affine = len(r2)
    rp = r2 if eye else reproject_by_H(r2, Hinv)
    r2 = r2[centre_inside(rp, orig_w, orig_h)]
    ori = nets[1](quantize_u8(extract_patches(gray_view, r2, mrSize)))
    r3 = orinet_postprocess(r2, ori)
    rp = r3 if eye else reproject_by_H(r3, Hinv)
    _, src = reproject_filter(rp, orig_w, orig_h)
    r4, rp4 = r3[src], rp[src]
    d = nets[2](quantize_u8(extract_patches(gray_view, r4, mrSize)))
    return dict(det=r4, reproj=rp4, desc=d, counts=[len(kp), n_affine, len(r4)], aff=aff, ori=ori)


# --------------------------------------------------------------------------- batched LO-RANSAC(H) schedule (ransac_batched.c)
_rb = None


def batched_ransac_H(u, th=16.0, conf=0.99, max_samples=1000000, sym_check=1, seed=12345, error_type=0):
    """The device's batched LO-RANSAC(H) schedule restated on the CPU (oracle/ransac_batched.c): same counter-based
    generator, batch sizes and reduction orders -> the device's inlier mask and H, bit for bit."""
    global _rb
    if _rb is None:
        path = os.path.join(HERE, "libransac_batched.so")
        if not os.path.exists(path):
            build()
        _rb = C.CDLL(path)
    u = np.ascontiguousarray(u, np.float64)
    T = len(u)
    if T <= 20:
        max_samples = 1000  # matching.cpp:644-645
    H = np.zeros(9, np.float64)
    inl = np.zeros(max(T, 1), np.uint8)
    resid = np.zeros(max(T, 1), np.float64)
    stats = np.zeros(4, np.int32)
    J = C.c_double()
    _rb.orb_ransac_H(_p(u), T, C.c_double(th), C.c_double(conf), int(max_samples), int(sym_check), C.c_uint64(seed), int(error_type),
                     _p(H), _p(inl), _p(stats), C.byref(J), _p(resid))
    return dict(H=H, inl=inl[:T], I=int(stats[0]), samples=int(stats[1]), lo_count=int(stats[2]), oc_rejects=int(stats[3]),
                J=J.value, resid=resid[:T])


# --------------------------------------------------------------------------- LORANSACFiltering's empirical checks
K_SIGMA = 2 * 3.0 * np.sqrt(3.0)     # matching.cpp (k_sigma of the LAF points)


def laf_points(kp1, kp2):
    """The 18 doubles per correspondence that H_LAF_check / F_LAF_check build (matching.cpp:209-235 / :265-290): the two
    centres and the two pairs of LAF points x + k_sigma * (a12, a22) * s and x + k_sigma * (a11, a21) * s."""
    n = len(kp1)
    u = np.ones((n, 18), np.float64)
    u[:, 0], u[:, 1] = kp1["x"], kp1["y"]
    u[:, 3], u[:, 4] = kp2["x"], kp2["y"]
    u[:, 6] = u[:, 0] + K_SIGMA * kp1["a12"] * kp1["s"]
    u[:, 7] = u[:, 1] + K_SIGMA * kp1["a22"] * kp1["s"]
    u[:, 9] = u[:, 3] + K_SIGMA * kp2["a12"] * kp2["s"]
    u[:, 10] = u[:, 4] + K_SIGMA * kp2["a22"] * kp2["s"]
    u[:, 12] = u[:, 0] + K_SIGMA * kp1["a11"] * kp1["s"]
    u[:, 13] = u[:, 1] + K_SIGMA * kp1["a21"] * kp1["s"]
    u[:, 15] = u[:, 3] + K_SIGMA * kp2["a11"] * kp2["s"]
    u[:, 16] = u[:, 4] + K_SIGMA * kp2["a21"] * kp2["s"]
    return u


def ref_H_LAF_check(kp1, kp2, Hloran, thresh):
    """H_LAF_check (matching.cpp:250-308) with the REFERENCE's own HDsSymMax (oracle/_ref): keep mask."""
    L = ref()
    u = laf_points(kp1, kp2)
    Hl = np.ascontiguousarray(Hloran, np.float64)
    keep = np.ones(len(u), bool)
    err = np.zeros(3, np.float64)
    lin = np.zeros(6 * max(len(u), 1), np.float64)
    if thresh > 0:
        for i in range(len(u)):
            row = np.ascontiguousarray(u[i])
            L.HDsSymMax(_p(lin), _p(row), _p(Hl), _p(err), 3)
            keep[i] = not (np.sqrt(err[0] + err[1] + err[2]) > thresh)
    return keep


def ref_F_LAF_check(kp1, kp2, F, thresh):
    """F_LAF_check (matching.cpp:192-249) with the REFERENCE's own FDs (oracle/_ref): keep mask."""
    L = ref()
    u = laf_points(kp1, kp2)
    Fm = np.ascontiguousarray(F, np.float64)
    keep = np.ones(len(u), bool)
    err = np.zeros(3, np.float64)
    if thresh > 0:
        for i in range(len(u)):
            row = np.ascontiguousarray(u[i])
            L.FDs(_p(row), _p(Fm), _p(err), 3)
            keep[i] = not (np.sqrt(err[0]) + np.sqrt(err[1]) + np.sqrt(err[2]) > thresh)
    return keep


def naive_H_check(kp1, kp2, H, error=10.0):
    """NaiveHCheck (matching.cpp:1014-1043): number of correspondences within `error` px under H and under inv(H).
    cv::invert(DECOMP_LU) is third-party; numpy's LU inverse stands in (the decision margin is pixels, not ulps)."""
    Hm = np.asarray(H, np.float64).reshape(3, 3)
    Hi = np.linalg.inv(Hm)
    x1, y1, x2, y2 = kp1["x"], kp1["y"], kp2["x"], kp2["y"]
    den = Hm[2, 0] * x1 + Hm[2, 1] * y1 + Hm[2, 2]
    xa, ya = (Hm[0, 0] * x1 + Hm[0, 1] * y1 + Hm[0, 2]) / den, (Hm[1, 0] * x1 + Hm[1, 1] * y1 + Hm[1, 2]) / den
    d1 = (x2 - xa) ** 2 + (y2 - ya) ** 2
    den = Hi[2, 0] * x2 + Hi[2, 1] * y2 + Hi[2, 2]
    xa, ya = (Hi[0, 0] * x2 + Hi[0, 1] * y2 + Hi[0, 2]) / den, (Hi[1, 0] * x2 + Hi[1, 1] * y2 + Hi[1, 2]) / den
    d2 = (x1 - xa) ** 2 + (y1 - ya) ** 2
    return int(((d1 <= error * error) & (d2 <= error * error)).sum())


def empirical_checks(kp1, kp2, model, use_F, err_threshold=4.0, laf_coef=None):
    """The tail of LORANSACFiltering (matching.cpp:764-820) on the RANSAC inliers: returns (keep mask, H or F).
    H branch: H = inv(Hloran^T); NaiveHCheck < 8 empties the list; H_LAF_check at 3*HLAFCoef*err_threshold; < 8
    survivors empty the list.  F branch: F_LAF_check at LAFCoef*err_threshold."""
    n = len(kp1)
    if use_F:
        keep = ref_F_LAF_check(kp1, kp2, model, (2.0 if laf_coef is None else laf_coef) * err_threshold)
        if keep.sum() < 8:
            keep[:] = False
        return keep, np.asarray(model, np.float64).copy()
    Hl = np.asarray(model, np.float64).reshape(3, 3)
    Hout = np.linalg.inv(Hl.T)
    keep = np.ones(n, bool)
    if naive_H_check(kp1, kp2, Hout, 10.0) < 8:
        keep[:] = False
    k2 = ref_H_LAF_check(kp1, kp2, model, 3.0 * (12.0 if laf_coef is None else laf_coef) * err_threshold)
    keep &= k2
    if keep.sum() < 8:
        keep[:] = False
    return keep, Hout.ravel()


# --------------------------------------------------------------------------- f4 rows: Hamming matcher, ground-truth H filter
def match_hamming(q, t, max_distance):
    """MatchFLANNDistance (matching.cpp:574-633) with an exact linear index: bytes = floor(entry), the two nearest train rows
    by Hamming distance (ties: lower index first, the LinearIndex / KNNSimpleResultSet order), a match when d1 <= (int)
    max_distance, ratio = d1 / d2.  Returns MATCH_DTYPE rows in query order."""
    qb = np.floor(np.asarray(q, np.float32)).astype(np.uint8)
    tb = np.floor(np.asarray(t, np.float32)).astype(np.uint8)
    out = []
    md = int(np.float32(max_distance))
    for i in range(len(qb)):
        d = np.unpackbits(qb[i][None, :] ^ tb, axis=1).sum(1).astype(np.int64)
        order = np.argsort(d, kind="stable")
        if len(order) and d[order[0]] <= md:
            d1 = float(d[order[0]])
            d2 = float(d[order[1]]) if len(order) > 1 else 0.0
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = float(np.float64(d1) / np.float64(d2))
            out.append((i, int(order[0]), int(order[1]) if len(order) > 1 else -1, d1, d2, 0, ratio))
    return np.array(out, MATCH_DTYPE) if out else np.zeros(0, MATCH_DTYPE)


def ref_hmatrix_filter(xy1, xy2, H, error="sampson", err_threshold=4.0):
    """HMatrixFiltering (matching.cpp:917-1013) with the REFERENCE's own HDs / HDsSym / HDsSymMax (oracle/_ref): u is packed
    (image 2, image 1) as the reference does; returns (keep mask, errors)."""
    L = ref()
    xy1 = np.asarray(xy1, np.float64)
    xy2 = np.asarray(xy2, np.float64)
    T = len(xy1)
    u = np.ascontiguousarray(np.c_[xy2, np.ones(T), xy1, np.ones(T)])
    Hm = np.ascontiguousarray(H, np.float64)
    d = np.zeros(max(T, 1), np.float64)
    Z = np.zeros(max(T, 1) * 18, np.float64)
    pidx = np.arange(max(T, 1), dtype=np.int32)
    fn = {"sampson": L.HDs, "symm_max": L.HDsSymMax, "symm_sum": L.HDsSym}[error]
    if T:
        L.lin_hg(_p(u), _p(Z), _p(pidx), T)                   # matching.cpp:969: the linearised rows HDs reads
        fn(_p(Z), _p(u), _p(Hm), _p(d), T)
    th = np.float32(err_threshold * err_threshold)           # float th in the reference (:968)
    return d[:T] <= np.float64(th), d[:T]
