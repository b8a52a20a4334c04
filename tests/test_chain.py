"""The per-view device chain (modsgpu_describe_view, csrc/chain.cu) against the oracle and against the seam-by-seam
route.  Decisions (which regions survive, in which order) and descriptors must be identical; the region doubles are
bit-equal up to the OriNet rotation, whose atan2 / cos / sin come from CUDA's libm instead of glibc (<= 2 ulp each,
<= 5 ulp in a rotated entry of A)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
F7 = ("x", "y", "s", "a11", "a12", "a21", "a22")
ULP = 1e-14     # a rotated entry is a11*cos - a12*sin with cos / sin each within 2 ulp of glibc's; the reprojection by
                # H^-1 of a tilted view (entries up to the tilt) adds products of those with cancellation: <= 45 ulp of max(1, |a|)


def _gray(oracle, u8):
    from mods_light_zmq_b200 import synth
    return oracle.gray_from_bgr(synth.gray_to_bgr(u8))


def _close(a, b, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = ULP * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    assert (err <= tol).all(), (what, float(err.max()))


def _view(mg, oracle, u8, tilt, phi, zoom=1.0):
    """device view + the same view on the CPU (both bit-identical, test_synth_view_bit_exact) + H"""
    from mods_light_zmq_b200 import synth
    img = mg.image_from_bgr8(synth.gray_to_bgr(u8))
    view, H = mg.synth_view(img, tilt, phi, zoom, 0.5)
    return img, view, np.asarray(H, np.float64).reshape(3, 3), mg.image_download(view)


@pytest.mark.parametrize("tilt,phi", [(1.0, 0.0), (2.0, 0.6), (4.0, 2.2)])
def test_post_kernels_fed_oracle_net_outputs(mg, oracle, tilt, phi):
    """Rows a14-a16 in isolation: the device post-processing kernels get the ORACLE's net outputs (torch CPU) and must
    make the oracle's decisions and produce its numbers -- AffNet step bit-equal, OriNet step within 2 ulp."""
    from mods_light_zmq_b200 import synth
    from oracle import cnn_oracle as CN
    u8 = synth.blob_image(seed=77, w=480, h=360, n_blobs=700)
    img, view, H, gv = _view(mg, oracle, u8, tilt, phi)
    eye = tilt == 1.0
    h, w = gv.shape
    Hinv = None if eye else oracle.invert3(H)
    kp = oracle.detect_hessian(gv)
    regs = oracle.regions_from_keypoints(kp)
    assert len(regs) > 150
    aff = CN.affnet(oracle.quantize_u8(oracle.extract_patches(gv, regs)))
    # make sure both kinds of rejection occur in the sample
    aff[::17, 1] *= 9.0
    r2, _ = oracle.affnet_postprocess(regs, aff, w, h)
    rp = r2 if eye else oracle.reproject_by_H(r2, Hinv)
    keep = oracle.centre_inside(rp, 480, 360)
    got, n_affine = mg.debug_affnet_post(regs, aff, w, h, 480, 360, H=None if eye else H)
    assert n_affine == len(r2) and len(got) == int(keep.sum()) and 0 < len(got) < len(regs)
    for f in F7:
        assert np.array_equal(got["det"][f], r2[keep][f]), f
        assert np.array_equal(got["reproj"][f], rp[keep][f]), f
    # OriNet step on the survivors
    r2k = r2[keep]
    ori = CN.orinet(oracle.quantize_u8(oracle.extract_patches(gv, r2k)))
    r3 = oracle.orinet_postprocess(r2k, ori)
    rp3 = r3 if eye else oracle.reproject_by_H(r3, Hinv)
    _, src = oracle.reproject_filter(rp3, 480, 360)
    got3 = mg.debug_orinet_post(got, ori, 480, 360, H=None if eye else H)
    assert len(got3) == len(src) and 0 < len(src) < len(r2k)
    for f in F7:
        _close(got3["det"][f], r3[src][f], "det." + f)
        _close(got3["reproj"][f], rp3[src][f], "reproj." + f)
    view.free(); img.free()


@pytest.mark.parametrize("tilt,phi", [(1.0, 0.0), (2.0, 1.1)])
def test_describe_view_equals_oracle_chain(mg, oracle, tilt, phi):
    """The whole chain on a small view.  The device nets differ from torch in the last bits (fp16 operands), so the CPU
    chain is walked with the DEVICE net outputs taken through the seam call modsgpu_describe; lists, order, counts and
    descriptors must then be identical."""
    import mods_light_zmq_b200 as M
    from mods_light_zmq_b200 import synth
    u8 = synth.blob_image(seed=5, w=400, h=300, n_blobs=500)
    img, view, H, gv = _view(mg, oracle, u8, tilt, phi)
    eye = tilt == 1.0
    h, w = gv.shape
    Hinv = None if eye else oracle.invert3(H)
    # CPU chain with the device nets evaluated through the seam call modsgpu_describe
    kp = oracle.detect_hessian(gv)
    regs = oracle.regions_from_keypoints(kp)
    aff = mg.describe(M.AFFNET, view, regs)
    r2, _ = oracle.affnet_postprocess(regs, aff, w, h)
    n_affine = len(r2)
    rp = r2 if eye else oracle.reproject_by_H(r2, Hinv)
    r2 = r2[oracle.centre_inside(rp, 400, 300)]
    ori = mg.describe(M.ORINET, view, r2)
    r3 = oracle.orinet_postprocess(r2, ori)
    rp3 = r3 if eye else oracle.reproject_by_H(r3, Hinv)
    _, src = oracle.reproject_filter(rp3, 400, 300)
    r4, rp4 = r3[src], rp3[src]
    d = mg.describe(M.HARDNET, view, r4)
    rows, desc, counts = mg.describe_view(view, None if eye else H, 400, 300)
    assert counts == [len(kp), n_affine, len(r4)] and len(rows) == len(r4) > 50
    for f in F7:
        _close(rows["det"][f], r4[f], "det." + f)
        _close(rows["reproj"][f], rp4[f], "reproj." + f)
    assert np.array_equal(desc, d)
    assert rows["octave"].min() >= 0 and set(np.unique(rows["type"])) <= {0, 1, 2}
    view.free(); img.free()


def _features_equal(a, b):
    assert len(a) == len(b) and len(a) > 0, (len(a), len(b))
    for f in F7 + ("response",):
        _close(a[f], b[f], f)
    for f in ("octave", "type", "view"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["desc"], b["desc"])


def test_chain_equals_seam_route_full_size(mg, synth_pair):
    """config 3's image A (1024x768, ~4.4k keypoints, large-window regions included): the device chain returns what
    the seam-by-seam route (modsgpu_detect + 3 x modsgpu_describe + host arithmetic) returns, identity view and a
    tilted / rotated view schedule."""
    import mods_light_zmq_b200 as M
    from mods_light_zmq_b200 import synth
    a, _, _ = synth_pair
    img = mg.image_from_bgr8(synth.gray_to_bgr(a))
    views = M.view_schedule([1.0], [1.0, 3.0], 120.0, 0.5)
    assert len(views) >= 3
    try:
        os.environ["MODSGPU_SEAM_CHAIN"] = "1"
        seam = mg.extract_features(img)
        seam_v = mg.extract_features_views(img, views)
    finally:
        os.environ.pop("MODSGPU_SEAM_CHAIN", None)
    l0 = mg.launch_count
    chain = mg.extract_features(img)
    chain_launches = mg.launch_count - l0
    chain_v = mg.extract_features_views(img, views)
    _features_equal(chain, seam)
    _features_equal(chain_v, seam_v)
    assert len(chain) > 3000 and chain_launches > 20
    # twice the same answer (graph replay, recycled workspaces)
    again = mg.extract_features(img)
    assert again.tobytes() == chain.tobytes()
    img.free()


def test_chain_other_detector_modes_and_empty(mg):
    """FixedRegNumber truncates the sorted list before the chain (prepareKeysForExport); a flat image yields nothing."""
    import mods_light_zmq_b200 as M
    from mods_light_zmq_b200 import synth
    u8 = synth.blob_image(seed=9, w=320, h=240, n_blobs=300)
    img = mg.image_from_bgr8(synth.gray_to_bgr(u8))
    p = mg.default_params()
    p.detectorMode = M.FIXED_REG_NUMBER
    p.reg_number = 60
    rows, desc, counts = mg.describe_view(img, params=p)
    assert counts[0] == 60 and len(rows) == counts[2] <= counts[1] <= 60 and len(desc) == len(rows)
    kp = mg.detect(img, p)
    assert len(kp) == 60
    flat = mg.image_from_bgr8(np.full((240, 320, 3), 128, np.uint8))
    rows, desc, counts = mg.describe_view(flat)
    assert len(rows) == 0 and counts == [0, 0, 0]
    flat.free(); img.free()


@pytest.mark.gpu
def test_pair_overlap_equals_sequential(mg):
    """modsgpu_set_pair_overlap: image 2 extracted on the sibling context from a helper thread (mods.cpp:234-251 runs the
    two images as concurrent OpenMP tasks) -- counts, H and the verified correspondences equal the one-stream run's."""
    from mods_light_zmq_b200 import synth
    pairs = []
    for seed, (w, h) in ((77, (640, 480)), (78, (1024, 768))):
        a, b, _ = synth.image_pair(seed=seed, w=w, h=h)
        A, B = synth.gray_to_bgr(a), synth.gray_to_bgr(b)
        pairs.append((A, B, mg.pair_pipeline(A, B, seed=9)))
    mg.set_pair_overlap(True)
    try:
        for it in range(6):     # from the second call on the sibling exists and its detector graph is captured
            A, B, ref = pairs[it % 2]
            got = mg.pair_pipeline(A, B, seed=9)
            for k in ("keypoints", "regions", "descriptors", "tentatives", "unique_tentatives", "inliers"):
                assert got[k] == ref[k], (it, k, got[k], ref[k])
            assert np.array_equal(got["H"], ref["H"]) and np.array_equal(got["inlier_xy"], ref["inlier_xy"]), it
    finally:
        mg.set_pair_overlap(False)


@pytest.mark.gpu
def test_detector_generic_blur_path_equals_specialised(mg, tmp_path):
    """The ksize-specialised blur (k_blur3, response of the source tile and half-size image fused in) and the generic route
    (k_blur_resp2 + k_response + k_half, taken for ksizes outside 7..23; forced here with MODSGPU_NO_BLUR3=1 in a child
    process -- the switch is read once per process) give byte-equal keypoint lists."""
    import subprocess
    import sys
    from mods_light_zmq_b200 import synth
    root = os.path.dirname(HERE)
    a, _, _ = synth.image_pair(seed=5, w=800, h=640)            # 800 -> 400 -> 200 -> 100 -> 50 -> 25: ragged widths too
    kp = mg.detect(mg.image_from_bgr8(synth.gray_to_bgr(a)))
    out = str(tmp_path / "kp.npy")
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import mods_light_zmq_b200 as M; from mods_light_zmq_b200 import synth; "
            "mg = M.ModsGpu(0, load_nets=False); a, _, _ = synth.image_pair(seed=5, w=800, h=640); "
            "np.save(%r, mg.detect(mg.image_from_bgr8(synth.gray_to_bgr(a))))" % (root, out))
    env = dict(os.environ, MODSGPU_NO_BLUR3="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
    ref = np.load(out)
    assert len(kp) == len(ref) > 500 and kp.tobytes() == ref.tobytes()


@pytest.mark.gpu
def test_pair_pipeline_device_descriptors_equal_host_route(mg):
    """The pair-level call keeps the HardNet++ descriptors on the device between the net and the matcher
    (modsgpu_describe_view_dev + modsgpu_match_fginn_dev); MODSGPU_HOST_DESC=1 (read per call) restores the read-back /
    re-upload of the seam route.  Same tentatives, same verified set, same H."""
    from mods_light_zmq_b200 import synth
    a, b, _ = synth.image_pair(seed=31, w=800, h=600)
    A, B = synth.gray_to_bgr(a), synth.gray_to_bgr(b)
    dev = mg.pair_pipeline(A, B, seed=4)           # descriptors on the device, matcher + duplicate filter in one call
    assert dev["tentatives"] > 100 and dev["unique_tentatives"] < dev["tentatives"]
    for env in ("MODSGPU_UNFUSED_TAIL", "MODSGPU_HOST_DESC"):      # two device calls; the host round trip of the seam route
        os.environ[env] = "1"
        try:
            other = mg.pair_pipeline(A, B, seed=4)
        finally:
            del os.environ[env]
        for k in ("keypoints", "regions", "descriptors", "tentatives", "unique_tentatives", "inliers"):
            assert dev[k] == other[k], (env, k)
        assert np.array_equal(dev["H"], other["H"]) and np.array_equal(dev["inlier_xy"], other["inlier_xy"]), env


@pytest.mark.gpu
def test_plain_c_example_equals_binding(mg, tmp_path):
    """examples/mods_pair.c (plain C11 over include/modsgpu.h, built by `make`): two PGM files in, verified
    correspondences out -- the same counts, H and correspondences as the ctypes binding gives for the same pair."""
    import subprocess
    from mods_light_zmq_b200 import synth
    root = os.path.dirname(HERE)
    exe = os.path.join(root, "examples", "mods_pair")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", root, "example"], check=True, timeout=300)
    a, b, _ = synth.image_pair(seed=12, w=640, h=480)
    paths = []
    for name, u8 in (("a.pgm", a), ("b.pgm", b)):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(b"P5\n# synthetic\n%d %d\n255\n" % (u8.shape[1], u8.shape[0]))
            f.write(np.ascontiguousarray(u8, np.uint8).tobytes())
        paths.append(p)
    ref = mg.pair_pipeline(synth.gray_to_bgr(a), synth.gray_to_bgr(b), seed=12345)
    for overlap in ("0", "1"):
        r = subprocess.run([exe, paths[0], paths[1], os.path.join(root, "weights"), overlap], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        xy = np.array([[float(v) for v in line.split()] for line in r.stdout.splitlines() if line.strip()])
        assert "verified %d" % ref["inliers"] in r.stderr and "tentatives %d," % ref["tentatives"] in r.stderr, r.stderr
        assert xy.shape == (ref["inliers"], 4) and np.allclose(xy, ref["inlier_xy"], rtol=0, atol=1e-6)
