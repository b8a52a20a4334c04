"""mods_light_zmq_b200 -- ctypes binding of libmodsgpu.so (include/modsgpu.h).

The library is the product; this module only marshals numpy arrays across the C ABI so that
tests/ and bench.py can call it.  There is no CPU fallback: importing works without a GPU (the
not-gpu tests check the exported symbols), but `ModsGpu()` raises unless an sm_100 device and the
built library are present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmodsgpu.so")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # see modsgpu_create (api.cu)
WEIGHTS_DIR = os.path.join(os.path.dirname(HERE), "weights")

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("s", "f4"), ("response", "f4"), ("type", "i4"), ("octave", "i4"),
                     ("level", "i4"), ("r0", "i4"), ("c0", "i4"), ("r", "i4"), ("c", "i4"), ("seq", "i4")])
REGION_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("s", "f8"), ("a11", "f8"), ("a12", "f8"), ("a21", "f8"), ("a22", "f8")])
MATCH_DTYPE = np.dtype([("qi", "i4"), ("ti", "i4"), ("tj_bad", "i4"), ("d1", "f4"), ("d2", "f4"), ("_pad", "i4"), ("ratio", "f8")])

FEATURE_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("s", "f8"), ("a11", "f8"), ("a12", "f8"), ("a21", "f8"), ("a22", "f8"),
                          ("response", "f8"), ("octave", "i4"), ("type", "i4"), ("view", "i4"), ("_pad", "i4"),
                          ("desc", "f4", (128,))])
_R7 = [("x", "f8"), ("y", "f8"), ("s", "f8"), ("a11", "f8"), ("a12", "f8"), ("a21", "f8"), ("a22", "f8")]
VIEW_REGION_DTYPE = np.dtype([("det", _R7), ("reproj", _R7), ("response", "f8"), ("octave", "i4"), ("type", "i4")])
VIEW_DTYPE = np.dtype([("tilt", "f8"), ("phi", "f8"), ("zoom", "f8"), ("InitSigma", "f8"), ("doBlur", "i4"), ("_pad", "i4")])

AFFNET, ORINET, HARDNET = 0, 1, 2
NET_FILES = {AFFNET: "affnet.npz", ORINET: "orinet.npz", HARDNET: "hardnet.npz"}
NET_DIM = {AFFNET: 3, ORINET: 2, HARDNET: 128}


class PyrParams(C.Structure):
    _fields_ = [("numberOfScales", C.c_int), ("initialSigma", C.c_float), ("threshold", C.c_float),
                ("edgeEigenValueRatio", C.c_double), ("border", C.c_int),
                ("detectorMode", C.c_int), ("rel_threshold", C.c_float), ("reg_number", C.c_int), ("rel_reg_number", C.c_float)]


FIXED_TH, RELATIVE_TH, FIXED_REG_NUMBER, RELATIVE_REG_NUMBER, NOT_LESS_THAN_REGIONS = range(5)


class AffShapeParams(C.Structure):
    _fields_ = [("maxIterations", C.c_int), ("convergenceThreshold", C.c_float), ("smmWindowSize", C.c_int),
                ("initialSigma", C.c_float), ("doBaumberg", C.c_int)]


class RansacParams(C.Structure):
    _fields_ = [("th", C.c_double), ("conf", C.c_double), ("max_samples", C.c_int), ("do_sym_check", C.c_int),
                ("seed", C.c_uint64), ("error_type", C.c_int), ("_pad", C.c_int)]


ERR_SAMPSON, ERR_SYMM_MAX, ERR_SYMM_SUM = 0, 1, 2      # RANSAC_error_t, matching.hpp:95


class PairResult(C.Structure):
    _fields_ = [("keypoints", C.c_int * 2), ("regions", C.c_int * 2), ("descriptors", C.c_int * 2),
                ("tentatives", C.c_int), ("unique_tentatives", C.c_int), ("inliers", C.c_int), ("H", C.c_double * 9)]


class PipelineParams(C.Structure):
    """modsgpu_pipeline_params: PyramidParams + MatchPars + RANSACPars of one run"""
    _fields_ = [("pyr", PyrParams), ("mrSize", C.c_double), ("patchSize", C.c_int), ("_pad0", C.c_int),
                ("fginn_threshold", C.c_double), ("contrad_dist", C.c_double), ("dup_filter_radius", C.c_double),
                ("nn", C.c_int), ("_pad1", C.c_int),
                ("err_threshold", C.c_double), ("confidence", C.c_double), ("HLAFCoef", C.c_double), ("LAFCoef", C.c_double),
                ("max_samples", C.c_int), ("do_symm_check", C.c_int), ("error_type", C.c_int), ("just_mark_outliers", C.c_int),
                ("use_F", C.c_int), ("_pad2", C.c_int), ("seed", C.c_uint64)]


def default_pipeline_params():
    p = PipelineParams()
    load_library().modsgpu_default_pipeline_params(C.byref(p))
    return p


class ModsStep(C.Structure):
    _fields_ = [("scale_set", C.c_double * 8), ("n_scales", C.c_int), ("tilt_set", C.c_double * 8), ("n_tilts", C.c_int),
                ("phi", C.c_double), ("init_sigma", C.c_double), ("fginn_threshold", C.c_double), ("do_blur", C.c_int),
                ("_pad", C.c_int)]


class ModsResult(C.Structure):
    _fields_ = [("steps_done", C.c_int), ("views", C.c_int * 2), ("regions", C.c_int * 2), ("tentatives", C.c_int),
                ("unique_tentatives", C.c_int), ("inliers", C.c_int), ("model", C.c_double * 9)]


class RansacResult(C.Structure):
    _fields_ = [("n_inliers", C.c_int), ("J", C.c_double), ("samples", C.c_int), ("lo_runs", C.c_int),
                ("oc_rejects", C.c_int), ("degen_runs", C.c_int), ("h_inliers", C.c_int)]


_lib = None


def load_library():
    """dlopen libmodsgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmodsgpu.so is not built: run `make` (or __graft_entry__.build()) first")
        _lib = C.CDLL(LIB_PATH)
        _lib.modsgpu_last_error.restype = C.c_char_p
        _lib.modsgpu_version.restype = C.c_char_p
        _lib.modsgpu_last_device_ms.restype = C.c_float
        _lib.modsgpu_launch_count.restype = C.c_longlong
        _lib.modsgpu_stream.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ModsGpuError(RuntimeError):
    pass


class Image:
    def __init__(self, owner, handle, w, h):
        self.owner, self.handle, self.w, self.h = owner, handle, w, h

    def free(self):
        if self.handle:
            self.owner.lib.modsgpu_image_free(self.owner.ctx, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ModsGpu:
    """One modsgpu_ctx (one CUDA stream).  Method names follow the C ABI."""

    def __init__(self, device=0, load_nets=False):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.modsgpu_create(int(device), C.byref(self.ctx))
        if rc != 0:
            raise ModsGpuError("modsgpu_create failed (%d): no sm_100 CUDA device -- there is no CPU path" % rc)
        if load_nets:
            for net in (AFFNET, ORINET, HARDNET):
                self.load_weights(net)

    def close(self):
        if self.ctx:
            self.lib.modsgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def _check(self, rc):
        if rc != 0:
            raise ModsGpuError("modsgpu error %d: %s" % (rc, self.lib.modsgpu_last_error(self.ctx).decode()))

    @property
    def last_device_ms(self):
        return float(self.lib.modsgpu_last_device_ms(self.ctx))

    @property
    def launch_count(self):
        return int(self.lib.modsgpu_launch_count(self.ctx))

    # ---- images
    def image_from_bgr8(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        hd = C.c_void_p()
        self._check(self.lib.modsgpu_image_from_bgr8(self.ctx, _p(bgr), w, h, C.byref(hd)))
        return Image(self, hd, w, h)

    def image_from_gray32f(self, gray):
        gray = np.ascontiguousarray(gray, np.float32)
        h, w = gray.shape
        hd = C.c_void_p()
        self._check(self.lib.modsgpu_image_from_gray32f(self.ctx, _p(gray), w, h, w, C.byref(hd)))
        return Image(self, hd, w, h)

    def image_download(self, img):
        out = np.empty((img.h, img.w), np.float32)
        self._check(self.lib.modsgpu_image_download(self.ctx, img.handle, _p(out)))
        return out

    def detect_affine(self, img, params=None, aff=None):
        """modsgpu_detect_affine: Hessian-Affine with the in-pyramid Baumberg iteration -> (keypoints, A [n x 4])"""
        if params is None:
            params = PyrParams()
            self.lib.modsgpu_default_pyr_params(C.byref(params))
        if aff is None:
            aff = AffShapeParams()
            self.lib.modsgpu_default_affshape_params(C.byref(aff))
        out, A = C.c_void_p(), C.c_void_p()
        n = C.c_int()
        self._check(self.lib.modsgpu_detect_affine(self.ctx, img.handle, C.byref(params), C.byref(aff), C.byref(out), C.byref(A), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, KP_DTYPE), np.zeros((0, 4), np.float32)
            kb = (C.c_char * (n.value * KP_DTYPE.itemsize)).from_address(out.value)
            ab = (C.c_char * (n.value * 16)).from_address(A.value)
            return np.frombuffer(kb, KP_DTYPE).copy(), np.frombuffer(ab, np.float32).reshape(-1, 4).copy()
        finally:
            self.lib.modsgpu_free(out)
            self.lib.modsgpu_free(A)

    # ---- view synthesis
    def synth_view(self, img, tilt, phi, zoom, init_sigma, do_blur=1):
        """modsgpu_synth_view -> (Image, H[3x3])"""
        hd = C.c_void_p()
        H = np.zeros(9, np.float64)
        self._check(self.lib.modsgpu_synth_view(self.ctx, img.handle, C.c_double(tilt), C.c_double(phi), C.c_double(zoom),
                                                C.c_double(init_sigma), int(do_blur), C.byref(hd), _p(H)))
        w, h = C.c_int(), C.c_int()
        self.lib.modsgpu_image_size(hd, C.byref(w), C.byref(h))
        return Image(self, hd, w.value, h.value), H.reshape(3, 3)

    # ---- detector
    def detect(self, img, params=None):
        if params is None:
            params = PyrParams()
            self.lib.modsgpu_default_pyr_params(C.byref(params))
        out = C.c_void_p()
        n = C.c_int()
        self._check(self.lib.modsgpu_detect(self.ctx, img.handle, C.byref(params), C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, KP_DTYPE)
            buf = (C.c_char * (n.value * KP_DTYPE.itemsize)).from_address(out.value)
            return np.frombuffer(buf, KP_DTYPE).copy()
        finally:
            self.lib.modsgpu_free(out)

    def gaussian_blur(self, img, sigma):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        out = np.empty_like(img)
        self._check(self.lib.modsgpu_gaussian_blur(self.ctx, _p(img), _p(out), w, h, C.c_float(sigma)))
        return out

    def hessian_response(self, img, norm):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        out = np.empty_like(img)
        self._check(self.lib.modsgpu_hessian_response(self.ctx, _p(img), _p(out), w, h, C.c_float(norm)))
        return out

    def half_image(self, img):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        ow, oh = int(np.rint(w * 0.5)), int(np.rint(h * 0.5))
        out = np.empty((oh, ow), np.float32)
        self._check(self.lib.modsgpu_half_image(self.ctx, _p(img), w, h, _p(out)))
        return out

    # ---- sampler / nets
    def extract_patches(self, img, regions, mrSize=5.1962, patchSize=32):
        regions = np.ascontiguousarray(regions, REGION_DTYPE)
        out = np.empty((len(regions), patchSize, patchSize), np.uint8)
        self._check(self.lib.modsgpu_extract_patches(self.ctx, img.handle, _p(regions), len(regions), C.c_double(mrSize),
                                                     patchSize, _p(out)))
        return out

    def load_weights(self, net, path=None):
        path = path or os.path.join(WEIGHTS_DIR, NET_FILES[net])
        self._check(self.lib.modsgpu_load_weights(self.ctx, net, path.encode()))

    def describe(self, net, img, regions, mrSize=5.1962, patchSize=32):
        regions = np.ascontiguousarray(regions, REGION_DTYPE)
        out = np.empty((len(regions), NET_DIM[net]), np.float32)
        self._check(self.lib.modsgpu_describe(self.ctx, net, img.handle, _p(regions), len(regions), C.c_double(mrSize),
                                              patchSize, _p(out)))
        return out

    def net_forward_u8(self, net, patches):
        patches = np.ascontiguousarray(patches, np.uint8).reshape(-1, 32, 32)
        out = np.empty((len(patches), NET_DIM[net]), np.float32)
        self._check(self.lib.modsgpu_net_forward_u8(self.ctx, net, _p(patches), len(patches), _p(out)))
        return out

    def debug_umma_probe(self, A, B, swap=0):
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        D = np.empty((128, 32), np.float32)
        self._check(self.lib.modsgpu_debug_umma_probe(self.ctx, _p(A), _p(B), _p(D), int(swap)))
        return D

    # ---- matching
    def match_fginn(self, q, t, txy, ratio=0.8, contrad=10.0, nn=50, want_knn=False):
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        txy = np.ascontiguousarray(txy, np.float64)
        nq, nt = len(q), len(t)
        dim = q.shape[1] if q.ndim == 2 and nq else (t.shape[1] if t.ndim == 2 and nt else 128)
        out = np.zeros(max(nq, 1), MATCH_DTYPE)
        n = C.c_int()
        ki = np.empty((nq, nn), np.int32) if want_knn else None
        kd = np.empty((nq, nn), np.float32) if want_knn else None
        self._check(self.lib.modsgpu_match_fginn(self.ctx, _p(q), nq, _p(t), _p(txy), nt, dim, C.c_double(ratio),
                                                 C.c_double(contrad), nn, _p(out), C.byref(n),
                                                 _p(ki) if want_knn else None, _p(kd) if want_knn else None))
        m = out[:n.value].copy()
        return (m, ki, kd) if want_knn else m

    def match_hamming(self, q, t, max_distance):
        """modsgpu_match_hamming = MatchFLANNDistance (binary descriptors as floats holding bytes)"""
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        nq, nt = len(q), len(t)
        dim = q.shape[1] if q.ndim == 2 and q.shape[1] else (t.shape[1] if t.ndim == 2 else 32)
        out = np.zeros(max(nq, 1), MATCH_DTYPE)
        n = C.c_int()
        self._check(self.lib.modsgpu_match_hamming(self.ctx, _p(q), nq, _p(t), nt, int(dim), C.c_double(max_distance), _p(out), C.byref(n)))
        return out[:n.value].copy()

    def duplicate_filter(self, xy1, xy2, ratio, r=2.0):
        xy1 = np.ascontiguousarray(xy1, np.float64)
        xy2 = np.ascontiguousarray(xy2, np.float64)
        ratio = np.ascontiguousarray(ratio, np.float64)
        T = len(ratio)
        order = np.zeros(max(T, 1), np.int32)
        n = C.c_int()
        self._check(self.lib.modsgpu_duplicate_filter(self.ctx, _p(xy1), _p(xy2), _p(ratio), T, C.c_double(r), _p(order),
                                                      C.byref(n)))
        return order[:n.value].copy()

    def ransac_H(self, u, th=16.0, conf=0.99, max_samples=1000000, sym_check=1, seed=12345, error_type=0):
        u = np.ascontiguousarray(u, np.float64)
        T = len(u)
        p = RansacParams(th, conf, 1000 if T <= 20 else max_samples, sym_check, seed, error_type, 0)
        H = np.zeros(9, np.float64)
        inl = np.zeros(max(T, 1), np.uint8)
        resid = np.zeros(max(T, 1), np.float64)
        res = RansacResult()
        self._check(self.lib.modsgpu_ransac_H_resid(self.ctx, _p(u), T, C.byref(p), _p(H), _p(inl), C.byref(res), _p(resid)))
        return dict(H=H, inl=inl[:T], I=res.n_inliers, J=res.J, samples=res.samples, lo_count=res.lo_runs,
                    oc_rejects=res.oc_rejects, resid=resid[:T])

    # ---- classic stages
    def dominant_orientation(self, img, regs, mr_size=5.1962, patch_size=32, max_angles=1, th=0.8):
        regs = np.ascontiguousarray(regs, REGION_DTYPE)
        n_ang = np.zeros(max(len(regs), 1), np.int32)
        ang = np.zeros((max(len(regs), 1), max(max_angles, 1)), np.float32)
        self._check(self.lib.modsgpu_dominant_orientation(self.ctx, img.handle, _p(regs), len(regs), C.c_double(mr_size),
                                                          patch_size, max_angles, C.c_double(th), _p(n_ang), _p(ang)))
        return n_ang[:len(regs)], ang[:len(regs)]

    def describe_sift(self, img, regs, mr_size=5.1962, patch_size=41, photo_norm=1, root_sift=1):
        regs = np.ascontiguousarray(regs, REGION_DTYPE)
        out = np.zeros((max(len(regs), 1), 128), np.float32)
        self._check(self.lib.modsgpu_describe_sift(self.ctx, img.handle, _p(regs), len(regs), C.c_double(mr_size), patch_size,
                                                   int(photo_norm), int(root_sift), _p(out)))
        return out[:len(regs)]

    def extract_patches_f32(self, img, regs, mr_size=5.1962, patch_size=41):
        """float patches of the 3-step sampler (what DescribeRegions hands to the SIFT descriptor)"""
        regs = np.ascontiguousarray(regs, REGION_DTYPE)
        out = np.zeros((max(len(regs), 1), patch_size, patch_size), np.float32)
        self._check(self.lib.modsgpu_extract_patches_f32(self.ctx, img.handle, _p(regs), len(regs), C.c_double(mr_size),
                                                         patch_size, _p(out)))
        return out[:len(regs)]

    # ---- one view on the device: detect -> AffNet -> OriNet -> HardNet++ with the region list resident (chain.cu)
    def describe_view(self, img, H=None, orig_w=None, orig_h=None, params=None, mrSize=5.1962, patchSize=32):
        """modsgpu_describe_view.  Returns (regions [VIEW_REGION_DTYPE], desc [n,128] float32, counts [3])."""
        p = params or self.default_params()
        rows, desc, n = C.c_void_p(), C.c_void_p(), C.c_int()
        counts = (C.c_int * 3)()
        Hp = None if H is None else _p(np.ascontiguousarray(H, np.float64))
        self._check(self.lib.modsgpu_describe_view(self.ctx, img.handle, Hp, int(orig_w or img.w), int(orig_h or img.h), C.byref(p),
                                                   C.c_double(mrSize), int(patchSize), C.byref(rows), C.byref(desc), C.byref(n), counts))
        try:
            m = n.value
            if m == 0:
                return np.zeros(0, VIEW_REGION_DTYPE), np.zeros((0, 128), np.float32), list(counts)
            r = np.frombuffer((C.c_char * (m * VIEW_REGION_DTYPE.itemsize)).from_address(rows.value), VIEW_REGION_DTYPE).copy()
            d = np.frombuffer((C.c_char * (m * 512)).from_address(desc.value), np.float32).reshape(m, 128).copy()
            return r, d, list(counts)
        finally:
            self.lib.modsgpu_free(rows)
            self.lib.modsgpu_free(desc)

    def default_params(self):
        p = PyrParams()
        self.lib.modsgpu_default_pyr_params(C.byref(p))
        return p

    def debug_affnet_post(self, regs, aff, w, h, orig_w, orig_h, mrSize=5.1962, H=None):
        """the chain's AffNet post-processing kernel on caller-supplied net outputs: (survivors, n_affine)"""
        rows = np.zeros(len(regs), VIEW_REGION_DTYPE)
        for f in REGION_DTYPE.names:
            rows["det"][f] = regs[f]
        aff = np.ascontiguousarray(aff, np.float32)
        out = np.zeros(max(len(regs), 1), VIEW_REGION_DTYPE)
        na, no = C.c_int(), C.c_int()
        Hp = None if H is None else _p(np.ascontiguousarray(H, np.float64))
        self._check(self.lib.modsgpu_debug_affnet_post(self.ctx, _p(rows), _p(aff), len(rows), int(w), int(h), int(orig_w), int(orig_h),
                                                       C.c_double(mrSize), Hp, _p(out), C.byref(na), C.byref(no)))
        return out[:no.value].copy(), na.value

    def debug_orinet_post(self, rows, ori, orig_w, orig_h, H=None):
        rows = np.ascontiguousarray(rows, VIEW_REGION_DTYPE)
        ori = np.ascontiguousarray(ori, np.float32)
        out = np.zeros(max(len(rows), 1), VIEW_REGION_DTYPE)
        no = C.c_int()
        Hp = None if H is None else _p(np.ascontiguousarray(H, np.float64))
        self._check(self.lib.modsgpu_debug_orinet_post(self.ctx, _p(rows), _p(ori), len(rows), int(orig_w), int(orig_h), Hp, _p(out),
                                                       C.byref(no)))
        return out[:no.value].copy()

    # ---- one image -> described regions; OxAff writer (extract_features_batch)
    def extract_features(self, img):
        out = C.c_void_p()
        n = C.c_int()
        self._check(self.lib.modsgpu_extract_features(self.ctx, img.handle, C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, FEATURE_DTYPE)
            buf = (C.c_char * (n.value * FEATURE_DTYPE.itemsize)).from_address(out.value)
            return np.frombuffer(buf, FEATURE_DTYPE).copy()
        finally:
            self.lib.modsgpu_free(out)

    def extract_features_views(self, img, views):
        views = np.ascontiguousarray(views, VIEW_DTYPE)
        out = C.c_void_p()
        n = C.c_int()
        self._check(self.lib.modsgpu_extract_features_views(self.ctx, img.handle, _p(views), len(views), C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, FEATURE_DTYPE)
            buf = (C.c_char * (n.value * FEATURE_DTYPE.itemsize)).from_address(out.value)
            return np.frombuffer(buf, FEATURE_DTYPE).copy()
        finally:
            self.lib.modsgpu_free(out)

    def match_imgreps(self, lists, group_dets=(), group_descs=(), sep_dets=(), sep_descs=(), fginn=None, capacity=1 << 16):
        """modsgpu_match_imgreps = CorrespondenceBank::MatchImgReps.  lists: [(image 1|2, det, desc, FEATURE_DTYPE array)].
        Returns an [n, 7] array (x1 y1 x2 y2 d1 d2 ratio)."""
        class RegionList(C.Structure):
            _fields_ = [("image", C.c_int), ("n", C.c_int), ("det", C.c_char_p), ("desc", C.c_char_p), ("f", C.c_void_p)]
        keep = [np.ascontiguousarray(f, FEATURE_DTYPE) for _, _, _, f in lists]
        arr = (RegionList * max(len(lists), 1))()
        for a, (img, det, desc, _), f in zip(arr, lists, keep):
            a.image, a.n, a.det, a.desc, a.f = int(img), len(f), det.encode(), desc.encode(), f.ctypes.data
        out = np.zeros((capacity, 7), np.float64)
        n = C.c_int()
        thr = ",".join("%s=%r" % kv for kv in (fginn or {}).items())
        self._check(self.lib.modsgpu_match_imgreps(self.ctx, arr, len(lists), ",".join(group_dets).encode(), ",".join(group_descs).encode(),
                                                   ",".join(sep_dets).encode(), ",".join(sep_descs).encode(), thr.encode(), _p(out), capacity,
                                                   C.byref(n)))
        return out[:min(n.value, capacity)].copy()

    def mods_pair(self, img1, img2, steps, min_matches=10, use_F=False, seed=12345, capacity=8192):
        """modsgpu_mods_pair.  steps: list of dicts(scales, tilts, phi, init_sigma, fginn) -- one per iteration of an
        iters_*.ini schedule."""
        arr = (ModsStep * len(steps))()
        for a, st in zip(arr, steps):
            sc, ti = list(st.get("scales", [1.0])), list(st.get("tilts", [1.0]))
            a.n_scales, a.n_tilts = len(sc), len(ti)
            for i, v in enumerate(sc):
                a.scale_set[i] = v
            for i, v in enumerate(ti):
                a.tilt_set[i] = v
            a.phi, a.init_sigma = st.get("phi", 360.0), st.get("init_sigma", 0.2)
            a.fginn_threshold, a.do_blur = st.get("fginn", 0.8), st.get("do_blur", 1)
        res = ModsResult()
        xy = np.zeros((capacity, 4), np.float64)
        self._check(self.lib.modsgpu_mods_pair(self.ctx, img1.handle, img2.handle, arr, len(steps), int(min_matches),
                                               int(bool(use_F)), C.c_ulonglong(seed), C.byref(res), _p(xy), capacity))
        return dict(steps_done=res.steps_done, views=list(res.views), regions=list(res.regions), tentatives=res.tentatives,
                    unique_tentatives=res.unique_tentatives, inliers=res.inliers, model=np.array(list(res.model)),
                    inlier_xy=xy[:min(res.inliers, capacity)].copy())

    def match_features(self, f1, f2, desc_dim=128, fginn=0.8, use_F=False, seed=12345, capacity=8192):
        """modsgpu_match_features: the `read_pre_extracted` path of mods.cpp:216-229 + :262-356 on two region lists."""
        f1 = np.ascontiguousarray(f1, FEATURE_DTYPE)
        f2 = np.ascontiguousarray(f2, FEATURE_DTYPE)
        res = ModsResult()
        xy = np.zeros((capacity, 4), np.float64)
        self._check(self.lib.modsgpu_match_features(self.ctx, _p(f1), len(f1), _p(f2), len(f2), int(desc_dim), C.c_double(fginn),
                                                    int(bool(use_F)), C.c_ulonglong(seed), C.byref(res), _p(xy), capacity))
        return dict(regions=list(res.regions), tentatives=res.tentatives, unique_tentatives=res.unique_tentatives,
                    inliers=res.inliers, model=np.array(list(res.model)), inlier_xy=xy[:min(res.inliers, capacity)].copy())

    def verify_matches(self, f1, f2, matches, use_F=False, seed=12345, capacity=8192):
        """modsgpu_verify_matches: duplicate filter + LO-RANSAC + empirical checks on tentatives matched elsewhere"""
        f1 = np.ascontiguousarray(f1, FEATURE_DTYPE)
        f2 = np.ascontiguousarray(f2, FEATURE_DTYPE)
        matches = np.ascontiguousarray(matches, MATCH_DTYPE)
        res = ModsResult()
        xy = np.zeros((capacity, 4), np.float64)
        self._check(self.lib.modsgpu_verify_matches(self.ctx, _p(f1), len(f1), _p(f2), len(f2), _p(matches), len(matches),
                                                    int(bool(use_F)), C.c_ulonglong(seed), C.byref(res), _p(xy), capacity))
        return dict(regions=list(res.regions), tentatives=res.tentatives, unique_tentatives=res.unique_tentatives,
                    inliers=res.inliers, model=np.array(list(res.model)), inlier_xy=xy[:min(res.inliers, capacity)].copy())

    def ransac_F(self, u, th=16.0, conf=0.99, max_samples=1000000, sym_check=1, seed=12345):
        """modsgpu_ransac_F: LO-RANSAC for a fundamental matrix (exp_ransacFcustom, matching.cpp:722)."""
        u = np.ascontiguousarray(u, np.float64)
        T = len(u)
        p = RansacParams(th, conf, 1000 if T <= 20 else max_samples, sym_check, seed, 0, 0)
        F = np.zeros(9, np.float64)
        inl = np.zeros(max(T, 1), np.uint8)
        res = RansacResult()
        self._check(self.lib.modsgpu_ransac_F(self.ctx, _p(u), T, C.byref(p), _p(F), _p(inl), C.byref(res)))
        return dict(F=F, inl=inl[:T], I=res.n_inliers, J=res.J, samples=res.samples, lo_count=res.lo_runs,
                    sym_rejects=res.oc_rejects, degen_runs=res.degen_runs, h_inliers=res.h_inliers)


def view_schedule(scale_set, tilt_set, phi_base, init_sigma, do_blur=1, cap=4096):
    """modsgpu_view_schedule = SetVSPars (synth-detection.cpp:191-322); needs no GPU."""
    sc = np.ascontiguousarray(scale_set, np.float64)
    ti = np.ascontiguousarray(tilt_set, np.float64)
    out = np.zeros(cap, VIEW_DTYPE)
    n = load_library().modsgpu_view_schedule(_p(sc), len(sc), _p(ti), len(ti), C.c_double(phi_base), C.c_double(init_sigma),
                                             int(do_blur), _p(out), cap)
    if n < 0:
        raise ModsGpuError("modsgpu_view_schedule failed (%d)" % n)
    return out[:min(n, cap)].copy()


def write_oxaff(path, feats):
    """modsgpu_write_oxaff: SaveRegionsMichal text format (needs no GPU)."""
    feats = np.ascontiguousarray(feats, FEATURE_DTYPE)
    rc = load_library().modsgpu_write_oxaff(str(path).encode(), _p(feats), len(feats))
    if rc != 0:
        raise ModsGpuError("modsgpu_write_oxaff failed (%d)" % rc)


def write_regions(path, feats, fmt=None):
    """The three region formats of extract_features_batch.cpp:147-155: 'oxaff' (outputMikFormat), 'npz' (file name
    ending in .npz) or the native 'text' format."""
    if fmt is None:
        fmt = "npz" if str(path).endswith(".npz") else "oxaff"
    feats = np.ascontiguousarray(feats, FEATURE_DTYPE)
    fn = {"oxaff": "modsgpu_write_oxaff", "text": "modsgpu_write_regions_text", "npz": "modsgpu_write_regions_npz"}[fmt]
    rc = getattr(load_library(), fn)(str(path).encode(), _p(feats), len(feats))
    if rc != 0:
        raise ModsGpuError("%s failed (%d)" % (fn, rc))


def read_regions(path, fmt=None):
    """LoadRegionsNPZ (file name ending in .npz) / LoadRegions (text), imagerepresentation.cpp:1317-1512; needs no GPU."""
    if fmt is None:
        fmt = "npz" if str(path).endswith(".npz") else "text"
    fn = {"npz": "modsgpu_read_regions_npz", "text": "modsgpu_read_regions_text"}[fmt]
    lib = load_library()
    out, n = C.c_void_p(), C.c_int()
    rc = getattr(lib, fn)(str(path).encode(), C.byref(out), C.byref(n))
    if rc != 0:
        raise ModsGpuError("%s failed (%d)" % (fn, rc))
    try:
        if n.value == 0:
            return np.zeros(0, FEATURE_DTYPE)
        buf = (C.c_char * (n.value * FEATURE_DTYPE.itemsize)).from_address(out.value)
        return np.frombuffer(buf, FEATURE_DTYPE).copy()
    finally:
        lib.modsgpu_free(out)


def _pair_dict(res, xy):
    return dict(keypoints=list(res.keypoints), regions=list(res.regions), descriptors=list(res.descriptors),
                tentatives=res.tentatives, unique_tentatives=res.unique_tentatives, inliers=res.inliers,
                H=np.array(list(res.H)).reshape(3, 3), inlier_xy=xy[:res.inliers].copy())


def _pair_pipeline_images(self, img1, img2, seed=12345, capacity=4096):
    res = PairResult()
    xy = np.zeros((capacity, 4), np.float64)
    self._check(self.lib.modsgpu_pair_pipeline_images(self.ctx, img1.handle, img2.handle, C.c_ulonglong(seed),
                                                      C.byref(res), _p(xy), capacity))
    return _pair_dict(res, xy)


def _pair_pipeline_images_ex(self, img1, img2, params, capacity=4096):
    """modsgpu_pair_pipeline_images_ex: the deep pair pipeline with the run's parameter block"""
    res = PairResult()
    xy = np.zeros((capacity, 4), np.float64)
    self._check(self.lib.modsgpu_pair_pipeline_images_ex(self.ctx, img1.handle, img2.handle, C.byref(params), C.byref(res), _p(xy), capacity))
    return _pair_dict(res, xy)


def _pair_pipeline_classic_images(self, img1, img2, seed=12345, capacity=4096):
    res = PairResult()
    xy = np.zeros((capacity, 4), np.float64)
    self._check(self.lib.modsgpu_pair_pipeline_classic_images(self.ctx, img1.handle, img2.handle, C.c_ulonglong(seed),
                                                              C.byref(res), _p(xy), capacity))
    return _pair_dict(res, xy)


def _pair_pipeline(self, bgr1, bgr2, seed=12345, capacity=4096):
    """host BGR u8 images -> upload -> detect/describe x2 -> match -> dedup -> LO-RANSAC."""
    h, w, _ = bgr1.shape
    res = PairResult()
    xy = np.zeros((capacity, 4), np.float64)
    self._check(self.lib.modsgpu_pair_pipeline(self.ctx, _p(bgr1), _p(bgr2), w, h, C.c_ulonglong(seed),
                                               C.byref(res), _p(xy), capacity))
    return _pair_dict(res, xy)


def _set_pair_overlap(self, on=True):
    """modsgpu_set_pair_overlap: the two images of a pair side by side (image 2 on a sibling context), mods.cpp:234-251"""
    self._check(self.lib.modsgpu_set_pair_overlap(self.ctx, 1 if on else 0))


ModsGpu.set_pair_overlap = _set_pair_overlap
ModsGpu.pair_pipeline_images = _pair_pipeline_images
ModsGpu.pair_pipeline_classic_images = _pair_pipeline_classic_images
ModsGpu.pair_pipeline_images_ex = _pair_pipeline_images_ex
ModsGpu.pair_pipeline = _pair_pipeline


def regions_from_keypoints(kps):
    """DetectAffineRegions glue (synth-detection.hpp:79-112) for doBaumberg=0: A = I."""
    r = np.zeros(len(kps), REGION_DTYPE)
    r["x"], r["y"], r["s"] = kps["x"], kps["y"], kps["s"]
    r["a11"] = 1.0
    r["a22"] = 1.0
    return r
