"""CPU suite (-m "not gpu"): pins the oracle against golden vectors and checks the C-ABI library's
surface.  No GPU compute is attempted here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def test_oracle_blur_matches_cv2_bit_exact(oracle):
    """helpers.cpp:717-731 -> cv::GaussianBlur: the restated op order equals cv2 4.13 bit for bit."""
    z = np.load(os.path.join(GOLD, "cv2_pins.npz"))
    sig = z["sigmas"]
    n = 0
    for i in range(5):
        img = z["img%d" % i]
        for j, s in enumerate(sig):
            got = oracle.gaussian_blur(img, float(s))
            assert np.array_equal(got, z["blur%d_%d" % (i, j)]), (i, j, float(np.abs(got - z["blur%d_%d" % (i, j)]).max()))
            n += 1
    assert n == 35


def test_oracle_half_matches_cv2_bit_exact(oracle):
    """pyramid.cpp:476 cv::resize(0.5, INTER_LINEAR)."""
    z = np.load(os.path.join(GOLD, "cv2_pins.npz"))
    for i in range(5):
        got = oracle.half_image(z["img%d" % i])
        ref = z["half%d" % i]
        assert got.shape == ref.shape
        h, w = z["img%d" % i].shape
        # interior (both source neighbours exist) is bit-exact; where an odd source size clamps the +1
        # neighbour the oracle defines the arithmetic (cv2's IPP edge code differs by <= 1 ulp there)
        ih, iw = h // 2, w // 2
        assert np.array_equal(got[:ih, :iw], ref[:ih, :iw]), i
        assert np.abs(got - ref).max() <= 4e-6 * 255, i


def test_oracle_gaussian_kernel_rule(oracle):
    k = oracle.gaussian_kernel(1.5199)
    assert len(k) == 11 and abs(k.sum() - 1) < 1e-6 and np.allclose(k, k[::-1])
    assert len(oracle.gaussian_kernel(0.75)) == 5
    assert len(oracle.gaussian_kernel(2.4525)) == 15


def test_cnn_oracle_matches_original_checkpoints():
    """Folded-BN restatement vs the daemons' own nn.Sequential on the original .pth (golden)."""
    from oracle import cnn_oracle as CN
    z = np.load(os.path.join(GOLD, "cnn_golden.npz"))
    p = z["patches"]
    assert np.abs(CN.affnet(p) - z["affnet"]).max() < 2e-4
    assert np.abs(CN.orinet(p) - z["orinet"]).max() < 2e-4
    assert np.abs(CN.hardnet_raw(p) - z["hardnet_raw"]).max() < 2e-5
    d = np.abs(CN.hardnet(p) - z["hardnet"])
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_graf_counts_match_readme():
    """README.md:47-61 known answers: regions 3731/4527, descriptors 3358/4118 (+-2)."""
    r = json.load(open(os.path.join(GOLD, "graf_counts.json")))
    for i, name in enumerate(("graf1", "graf6")):
        assert abs(r["oracle"][name]["regions"] - r["readme"]["regions"][i]) <= 2
        assert abs(r["oracle"][name]["descriptors"] - r["readme"]["descriptors"][i]) <= 2


def test_graf_classic_counts_match_readme():
    """README.md:77-105 known answers of the CLASSIC configuration (Hessian-Affine + Baumberg, dominant orientation,
    RootSIFT): regions 2665/3287, descriptors 2331/2912 (+-2), ~21 RANSAC inliers out of ~74 tentatives (the README
    run uses the randomised kd-tree; the oracle's exact linear matcher finds 63)."""
    r = json.load(open(os.path.join(GOLD, "graf_counts.json")))
    for i, name in enumerate(("graf1", "graf6")):
        assert abs(r["oracle_classic"][name]["keypoints"] - r["readme_classic"]["regions"][i]) <= 2
        assert abs(r["oracle_classic"][name]["descriptors"] - r["readme_classic"]["descriptors"][i]) <= 2
    p = r["oracle_classic"]["pair"]
    assert abs(p["inliers_ref_degensac"] - r["readme_classic"]["inliers"]) <= 4
    assert 0.7 * r["readme_classic"]["unique"] <= p["unique"] <= 1.2 * r["readme_classic"]["unique"]


@pytest.mark.skipif(not os.path.exists("/root/reference/build/imgs/graf1.png"), reason="reference tree absent")
def test_graf1_classic_count_live(oracle):
    import cv2
    r = json.load(open(os.path.join(GOLD, "graf_counts.json")))
    g = oracle.gray_from_bgr(cv2.imread("/root/reference/build/imgs/graf1.png", cv2.IMREAD_COLOR))
    nk, regs, d = oracle.classic_regions(g)
    assert nk == r["oracle_classic"]["graf1"]["keypoints"] and len(regs) == r["oracle_classic"]["graf1"]["descriptors"]
    assert d.shape == (len(regs), 128) and d.min() >= 0 and d.max() <= 255 and np.all(d == np.round(d))
    assert np.abs(np.linalg.norm(d, axis=1) - 512).max() < 8          # RootSIFT: |d| = 512 up to the integer rounding


@pytest.mark.skipif(not os.path.exists("/root/reference/build/imgs/graf1.png"), reason="reference tree absent")
def test_graf1_detector_count_live(oracle):
    import cv2
    bgr = cv2.imread("/root/reference/build/imgs/graf1.png", cv2.IMREAD_COLOR)
    k = oracle.detect_hessian(oracle.gray_from_bgr(bgr))
    r = json.load(open(os.path.join(GOLD, "graf_counts.json")))
    assert len(k) == r["oracle"]["graf1"]["keypoints"]
    assert np.all(np.diff(np.abs(k["response"])) <= 0)


def test_reference_degensac_fixture(oracle):
    """The reference's own exp_ransacHcustom (oracle/_ref) reproduces the committed fixture."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libdegensac_ref.so not built")
    z = np.load(os.path.join(GOLD, "ransac_ref.npz"))
    r = oracle.ref_ransac_H(z["u"], th=16.0, seed_time=12345)
    assert r["I"] == int(z["I"]) and np.array_equal(r["inl"], z["inl"])
    assert np.allclose(r["H"] / r["H"][8], z["H"] / z["H"][8], rtol=1e-9, atol=1e-12)
    assert r["inl"][:150].sum() >= 145 and r["inl"][150:].sum() <= 3


def test_reference_degensac_F_fixture(oracle):
    """The reference's own exp_ransacFcustom (oracle/_ref) on the committed fixture: its model is the scene's
    epipolar geometry and the stored result lies in the band the reference itself produces."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libdegensac_ref.so not built")
    z = np.load(os.path.join(GOLD, "ransac_F_ref.npz"))
    mask = z["mask"]
    # NB the reference reads uninitialised heap in exp_ransacFcustom (errs[4] = errs[3] before any write,
    # exp_ranF.c:870-872; data_out histogram :1029): with time() pinned its result still differs between the
    # first and later calls of one process and between processes (I = 81..172 observed on this fixture, 170 true
    # inliers).  So the fixture pins a band over several calls, not a value.
    best = 0
    for _ in range(5):
        r = oracle.ref_ransac_F(z["u"], th=16.0, seed_time=12345)
        assert r["I"] <= int(mask.sum()) + 8, r["I"]
        d = oracle.sampson_F(r["F"], z["u"])
        assert ((d <= 16.0) == r["inl"].astype(bool)).mean() > 0.9
        assert (r["inl"].astype(bool) & mask).sum() >= 0.3 * mask.sum()      # I = 81 has been observed
        best = max(best, int((r["inl"].astype(bool) & mask).sum()))
    assert best >= 0.3 * mask.sum()      # a whole process can stay in the weak state, so no tighter bound holds
    d = oracle.sampson_F(z["F"], z["u"])
    assert np.array_equal(d <= 16.0, z["inl"].astype(bool)) and int(z["I"]) == int(z["inl"].sum())
    # the seeded scene's true F explains its inliers (checks sampson_F's layout convention too)
    assert np.median(oracle.sampson_F(z["F_true"].ravel(), z["u"])[z["mask"]]) < 1.0


def test_degensac_shim_exports_reference_symbols():
    import mods_light_zmq_b200 as M
    path = os.path.join(os.path.dirname(M.LIB_PATH), "libmodsgpu_degensac.so")
    assert os.path.exists(path), "run make"
    lib = ctypes.CDLL(path)
    for n in ("exp_ransacHcustom", "exp_ransacFcustom", "modsgpu_ransac_set_seed"):
        assert hasattr(lib, n), n


def test_oracle_view_synthesis_matches_cv2_bit_exact(oracle):
    """GenerateSynthImageCorr (synth-detection.cpp:324-518): the oracle's restatement of cv::warpAffine and the
    anisotropic cv::GaussianBlur is bit-identical to cv2 4.13 on the committed fixture."""
    z = np.load(os.path.join(GOLD, "synth_pins.npz"))
    img = z["img"]
    for i, (tilt, phi, zoom, isg) in enumerate(z["cases"]):
        ref = z["view%d" % i]
        got, H = oracle.synth_view(img, tilt, phi, zoom, isg)
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, tilt, phi, zoom)
        ow, oh, H2, ident = oracle.synth_geometry(img.shape[1], img.shape[0], tilt, phi, zoom)
        assert (oh, ow) == ref.shape and np.array_equal(H, H2)
        if ident:
            assert np.array_equal(H, np.eye(3))
        else:   # H maps the image centre into the view (SynthImage::H, original -> view)
            p = H @ np.array([img.shape[1] / 2, img.shape[0] / 2, 1.0])
            assert -2 <= p[0] <= ow + 2 and -2 <= p[1] <= oh + 2
    assert np.array_equal(oracle.gaussian_blur_xy(img[:, :61], 3, 5, 0.4, 0.8), z["blur_a"])
    assert np.array_equal(oracle.gaussian_blur_xy(img[:37, :50], 7, 3, 1.2, 0.1), z["blur_b"])


def test_view_schedule_matches_SetVSPars():
    """modsgpu_view_schedule vs a transcription of SetVSPars (synth-detection.cpp:191-322) on the schedules of
    build/iters_MODS_ZMQ.ini (HessianAffine steps: TiltSet=1,2,4,6,8; Phi=360 / 120) and a vertical-tilt case."""
    import mods_light_zmq_b200 as M

    def ref(scales, tilts, phi_base):
        out = []
        for sc in scales:
            for t in tilts:
                if abs(t - 1) > 0.01:
                    n_rot = int(np.floor(180.0 * t / phi_base))
                    dphi = np.pi / n_rot
                    if n_rot < 0:
                        n_rot, dphi = 1, 0.0
                        out.append((-t, 0.0, sc))
                    out += [(t, dphi * r, sc) for r in range(n_rot)]
                else:
                    out.append((t, 0.0, sc))
        return out
    for scales, tilts, phi in [([1.0], [1, 2, 4, 6, 8], 360.0), ([1.0], [1, 2, 4, 6, 8], 120.0), ([1, 0.25, 0.125], [1], 360.0),
                               ([1, 0.25], [1, 3, 6], 360.0), ([1.0], [1, 2], -360.0)]:
        v = M.view_schedule(scales, tilts, phi, 0.2)
        r = ref(scales, tilts, phi)
        assert len(v) == len(r)
        for a, b in zip(v, r):
            assert a["tilt"] == b[0] and abs(a["phi"] - b[1]) < 1e-15 and a["zoom"] == b[2] and a["doBlur"] == 1
    assert len(M.view_schedule([1.0], [1, 2, 4, 6, 8], 360.0, 0.2)) == 11


def test_oxaff_writer_matches_cv2_golden(tmp_path):
    """modsgpu_write_oxaff (SaveRegionsMichal text mode) against ellipse entries computed with cv2.SVDecomp the way
    saveKP_KM_format does (imagerepresentation.cpp:113-126); the file prints 6 significant digits."""
    import mods_light_zmq_b200 as M
    z = np.load(os.path.join(GOLD, "oxaff_golden.npz"))
    regs, abc = z["regs"], z["abc"]
    f = np.zeros(len(regs), M.FEATURE_DTYPE)
    for k in ("x", "y", "s", "a11", "a12", "a21", "a22"):
        f[k] = regs[k]
    rng = np.random.RandomState(0)
    f["desc"] = rng.randint(0, 256, (len(regs), 128))
    path = str(tmp_path / "img.oxaff")
    M.write_oxaff(path, f)
    lines = open(path).read().split("\n")
    assert lines[0] == "128" and int(lines[1]) == len(regs)
    rows = [l.split() for l in lines[2:] if l.strip()]
    assert len(rows) == len(regs) and all(len(r) == 5 + 128 for r in rows)
    got = np.array([[float(v) for v in r[:5]] for r in rows])
    assert np.allclose(got[:, 0], regs["x"], rtol=1e-5) and np.allclose(got[:, 1], regs["y"], rtol=1e-5)
    assert np.allclose(got[:, 2:], abc, rtol=3e-5, atol=1e-12), np.abs(got[:, 2:] / abc - 1).max()
    assert [int(v) for v in rows[3][5:]] == [int(v) for v in f["desc"][3]]
    M.write_oxaff(path, f[:0])
    assert open(path).read().split() == ["128", "0"]


def test_region_text_and_npz_writers(tmp_path):
    """SaveRegions text (imagerepresentation.cpp:1219-1255) and SaveRegionsNPZ (:1257-1316) layouts."""
    import mods_light_zmq_b200 as M
    rng = np.random.RandomState(3)
    n = 17
    f = np.zeros(n, M.FEATURE_DTYPE)
    for k in ("x", "y", "s", "a11", "a12", "a21", "a22", "response"):
        f[k] = rng.uniform(0.5, 900, n)
    f["desc"] = rng.randint(0, 256, (n, 128))
    p = str(tmp_path / "r.txt")
    M.write_regions(p, f, "text")
    lines = open(p).read().split("\n")
    assert lines[:4] == ["1", "HessianAffine 1", "ZMQ %d" % n, "128"]
    row = lines[4 + 5].split()
    assert len(row) == 7 + 1 + 128 and int(row[7]) == 128
    assert np.allclose([float(v) for v in row[:7]], [f[k][5] for k in ("x", "y", "s", "a11", "a12", "a21", "a22")], rtol=1e-5)
    assert [int(v) for v in row[8:]] == [int(v) for v in f["desc"][5]]
    p = str(tmp_path / "r.npz")
    M.write_regions(p, f)
    z = np.load(p)
    assert sorted(z.files) == ["A", "descs", "responses", "scales", "xy"]
    assert z["xy"].dtype == np.float64 and z["descs"].dtype == np.uint8 and z["descs"].shape == (n, 128)
    assert np.array_equal(z["xy"], np.c_[f["x"], f["y"]]) and np.array_equal(z["scales"][:, 0], f["s"])
    assert np.array_equal(z["A"], np.c_[f["a11"], f["a12"], f["a21"], f["a22"]]) and np.array_equal(z["responses"][:, 0], f["response"])
    assert np.array_equal(z["descs"], f["desc"].astype(np.uint8))
    M.write_regions(p, f[:0])
    assert np.load(p)["xy"].shape == (0, 2)


def test_region_readers(tmp_path):
    """LoadRegionsNPZ / PreLoadRegionsNPZ (imagerepresentation.cpp:1355-1512: A | angles | upright variants, descs as
    uchar) and LoadRegions text with the loadAR record layout (:237-253, :1317-1354)."""
    import mods_light_zmq_b200 as M
    rng = np.random.RandomState(4)
    n = 23
    f = np.zeros(n, M.FEATURE_DTYPE)
    for k in ("x", "y", "s", "a11", "a12", "a21", "a22", "response"):
        f[k] = rng.uniform(0.5, 900, n)
    f["desc"] = rng.randint(0, 256, (n, 128))
    p = str(tmp_path / "r.npz")
    M.write_regions(p, f)                                     # SaveRegionsNPZ -> LoadRegionsNPZ round trip, exact
    g = M.read_regions(p)
    for k in ("x", "y", "s", "a11", "a12", "a21", "a22", "response"):
        assert np.array_equal(g[k], f[k]), k
    assert np.array_equal(g["desc"], f["desc"]) and np.all(g["type"] == 4)          # DET_READ
    # numpy-written files (compressed, angles in degrees, 64-dim descriptors, mixed dtypes)
    ang = rng.uniform(-180, 180, n)
    np.savez_compressed(tmp_path / "a.npz", xy=np.c_[f["x"], f["y"]], scales=f["s"].astype(np.float32), responses=f["response"],
                        angles=ang, descs=f["desc"][:, :64].astype(np.uint8))
    g = M.read_regions(str(tmp_path / "a.npz"))
    a = ang * np.pi / 180.0
    assert np.array_equal(g["a11"], np.cos(a)) and np.array_equal(g["a12"], np.sin(a)) and np.array_equal(g["a21"], -np.sin(a))
    assert np.array_equal(g["s"], f["s"].astype(np.float32).astype(np.float64))
    assert np.array_equal(g["desc"][:, :64], f["desc"][:, :64]) and not g["desc"][:, 64:].any()
    np.savez(tmp_path / "u.npz", xy=np.c_[f["x"], f["y"]], scales=f["s"], responses=f["response"], descs=f["desc"].astype(np.uint8))
    g = M.read_regions(str(tmp_path / "u.npz"))
    assert np.all(g["a11"] == 1) and np.all(g["a12"] == 0) and np.all(g["a21"] == 0) and np.all(g["a22"] == 1)
    np.savez(tmp_path / "bad.npz", xy=np.c_[f["x"], f["y"]], scales=f["s"])          # members missing -> error, no crash
    with pytest.raises(M.ModsGpuError):
        M.read_regions(str(tmp_path / "bad.npz"))
    with pytest.raises(M.ModsGpuError):
        M.read_regions(str(tmp_path / "nothere.npz"))
    # text: 2 detectors, the second with an empty descriptor list
    rows = []
    for i in range(5):
        kp = "%g %g %g %g %g %g 1.5 2 %g 1" % (f["x"][i], f["y"][i], f["a11"][i], f["a12"][i], f["a21"][i], f["a22"][i], f["s"][i])
        rows.append("%d 0 3 -1 %s %s 128 %s" % (i, kp, kp, " ".join(str(int(v)) for v in f["desc"][i])))
    (tmp_path / "r.txt").write_text("2\nHessianAffine 1\nZMQ 5\n128\n" + "\n".join(rows) + "\nMSER 1\nZMQ 0\n")
    g = M.read_regions(str(tmp_path / "r.txt"))
    assert len(g) == 5 and np.all(g["view"] == 3) and np.all(g["octave"] == 2) and np.all(g["type"] == 1)
    for k in ("x", "y", "s", "a11", "a12", "a21", "a22"):
        assert np.allclose(g[k], f[k][:5], rtol=1e-5), k
    assert np.array_equal(g["desc"], f["desc"][:5])
    (tmp_path / "t.txt").write_text("1\nHessianAffine 1\nZMQ 2\n128\n0 0 0 0 1 2\n")   # truncated record
    with pytest.raises(M.ModsGpuError):
        M.read_regions(str(tmp_path / "t.txt"))


def test_oracle_matcher_small(oracle):
    rng = np.random.RandomState(0)
    t = rng.randint(0, 256, (70, 128)).astype(np.float32)
    q = t[:10].copy()
    q[:, :4] += 1
    txy = rng.uniform(0, 100, (70, 2))
    m = oracle.match_fginn(q, np.zeros((10, 2)), t, txy)
    assert len(m) == 10 and np.array_equal(m["ti"], np.arange(10)) and np.all(m["d1"] == 4)
    assert len(oracle.match_fginn(q[:0], np.zeros((0, 2)), t, txy)) == 0


def test_oracle_duplicate_filter(oracle):
    xy1 = np.array([[0, 0], [1, 0], [50, 50], [0.5, 0.5]], float)
    xy2 = np.array([[10, 10], [10.5, 10], [70, 70], [30, 30]], float)
    ratio = np.array([0.5, 0.4, 0.6, 0.3])
    keep = oracle.duplicate_filter(xy1, xy2, ratio, 2.0)
    assert list(keep) == [3, 1, 2]


def test_library_exports_every_declared_symbol():
    import mods_light_zmq_b200 as M
    hdr = open(os.path.join(ROOT, "include", "modsgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(modsgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = M.load_library()
    for n in names:
        assert hasattr(lib, n), n


def test_header_is_plain_c_and_example_refuses_without_gpu(tmp_path):
    """include/modsgpu.h compiles as C11 (no C++ in the boundary) and the plain-C example program builds against the
    shipped library; without an sm_100 device it stops at modsgpu_create with a message -- there is no CPU path."""
    import subprocess
    import torch
    exe = str(tmp_path / "mods_pair")
    pkg = os.path.join(ROOT, "mods_light_zmq_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "examples", "mods_pair.c"), "-L" + pkg, "-lmodsgpu", "-Wl,-rpath," + pkg], check=True, timeout=120)
    if torch.cuda.is_available():
        return
    pgm = str(tmp_path / "x.pgm")
    with open(pgm, "wb") as f:
        f.write(b"P5\n8 8\n255\n" + bytes(64))
    r = subprocess.run([exe, pgm, pgm, os.path.join(ROOT, "weights")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "modsgpu_create" in r.stderr and "no CPU path" in r.stderr


def test_no_cpu_fallback_without_gpu():
    """Without an sm_100 device the product refuses to run (MODSGPU_ENODEV) instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mods_light_zmq_b200 as M
    lib = M.load_library()
    ctx = ctypes.c_void_p()
    assert lib.modsgpu_create(0, ctypes.byref(ctx)) == -1
    with pytest.raises(M.ModsGpuError):
        M.ModsGpu(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mods_light_zmq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                for needle in ("pyoracle", "liboracle", "cnn_oracle", "from oracle", "import oracle", "mods_oracle.h"):
                    assert needle not in s, (f, needle)


def test_synthetic_pair_is_deterministic(synth_pair):
    import hashlib
    a, b, H = synth_pair
    assert a.shape == (768, 1024) and b.shape == (768, 1024) and a.dtype == np.uint8
    from mods_light_zmq_b200 import synth
    a2 = synth.blob_image()
    assert hashlib.sha1(a.tobytes()).hexdigest() == hashlib.sha1(a2.tobytes()).hexdigest()


def test_host_mirror_descvec_semantics(tmp_path):
    """csrc/host/mods_host.h: AffineRegion::desc keeps the std::vector<float> surface of the reference's descriptor.vec
    while region copies share one block (tests/cpp/descvec_test.cpp, header only, no GPU)."""
    import subprocess
    exe = str(tmp_path / "descvec_test")
    src = os.path.join(ROOT, "tests", "cpp", "descvec_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "descvec ok" in out.stdout, out.stdout + out.stderr


def test_oracle_knn_matches_cv2_flann_linear_index():
    """The matcher's third-party boundary ("parity unpinned at the FLANN boundary"): the oracle's restatement
    of cvflann::LinearIndex + KNNSimpleResultSet (oracle/mods_oracle.cpp:orc_knn_linear) against cv2 4.13's
    flann_Index(algorithm = LINEAR).knnSearch -- indices AND float distances, with exact duplicates, zero distances and a
    lattice case full of equal distances (tests/golden/flann_pins.npz, generated by make_golden.py flann).  The GPU matcher
    is compared with the same oracle function in test_match_fginn_bit_exact."""
    from oracle import pyoracle as O
    z = np.load(os.path.join(GOLD, "flann_pins.npz"))
    idx, dist = O.knn_linear(z["q"].astype(np.float32), z["t"].astype(np.float32), 50)
    assert np.array_equal(idx, z["idx"]) and np.array_equal(dist, z["dist"])
    assert list(idx[0, :3]) == [10, 50, 200] and np.all(dist[0, :3] == 0)      # equal distances keep train-index order
    idx2, dist2 = O.knn_linear(z["q2"], z["t2"], 50)
    assert np.array_equal(idx2, z["idx2"]) and np.array_equal(dist2, z["dist2"])
    try:
        import cv2
    except ImportError:
        return
    rng = np.random.RandomState(11)                       # live: a fresh case against the installed wheel
    t = rng.randint(0, 256, (300, 128)).astype(np.float32)
    t[100:110] = t[5]
    q = rng.randint(0, 256, (40, 128)).astype(np.float32)
    q[7] = t[5]
    ci, cd = cv2.flann_Index(t, dict(algorithm=0)).knnSearch(q, 50, params={})
    oi, od = O.knn_linear(q, t, 50)
    assert np.array_equal(ci, oi) and np.array_equal(cd, od)


def test_oracle_u8_quantisation_matches_cv2_imencode():
    """DescribeWithZmq hands the daemons a PNG of the float patch column (imagerepresentation.cpp:45): cv::imencode's
    fallback conversion to 8 bit (round half to even, saturate) against the oracle's quantize_u8, which the sampler
    kernel's u8 output is compared with bit for bit (tests/golden/imencode_pins.npz; live when the cv2 wheel is there)."""
    from oracle import pyoracle as O
    z = np.load(os.path.join(GOLD, "imencode_pins.npz"))
    assert np.array_equal(O.quantize_u8(z["patches"]), z["u8"])
    assert list(z["u8"][0, :8]) == [0, 2, 2, 4, 254, 255, 0, 128]
    try:
        import cv2
    except ImportError:
        return
    a = np.random.RandomState(5).uniform(-5, 260, (64, 32)).astype(np.float32)
    a[::7, ::5] = np.floor(a[::7, ::5]) + 0.5
    ok, buf = cv2.imencode(".png", a)
    assert ok and np.array_equal(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED), O.quantize_u8(a))
