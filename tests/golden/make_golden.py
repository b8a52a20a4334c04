#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ (run in the build container, where
/root/reference and the cv2 wheel exist; the GPU box only reads the .npz / .json results).

  cv2_pins.npz      cv2.GaussianBlur / cv2.resize outputs (the OpenCV calls whose arithmetic lives outside
                    the reference tree: helpers.cpp:717-731, pyramid.cpp:476) on small seeded images
  cnn_golden.npz    48 u8 patches + the outputs of the ORIGINAL AffNet/OriNet/HardNet++ checkpoints run
                    through the daemons' own nn.Sequential definitions (unfolded BatchNorm, torch CPU fp32)
  graf_counts.json  keypoint / region / descriptor counts of the oracle pipeline on the reference's
                    graf1/graf6 images, next to the README transcript values (README.md:47-61)
  synth_pins.npz    cv2 outputs of GenerateSynthImageCorr's warpAffine -> GaussianBlur(kx,ky) -> warpAffine chain
  oxaff_golden.npz  seeded affine regions + their OxAff ellipse entries (saveKP_KM_format with cv2.SVDecomp)
  ransac_F_ref.npz  a seeded two-view scene + the result of the reference's own exp_ransacFcustom
  ransac_ref.npz    a seeded correspondence set + the result of the reference's own exp_ransacHcustom
                    (oracle/_ref/libdegensac_ref.so, time() pinned)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def cv2_pins():
    import cv2
    rng = np.random.RandomState(7)
    out = {}
    sizes = [(64, 48), (50, 37), (25, 19), (18, 18), (33, 21)]
    sigmas = [0.75, 0.84375, 1.2262735, 1.5199, 1.9465878, 2.4525, 3.1]
    for i, (w, h) in enumerate(sizes):
        img = (rng.rand(h, w) * 255).astype(np.float32)
        out["img%d" % i] = img
        for j, s in enumerate(sigmas):
            s = float(np.float32(s))   # the reference passes float sigmas (helpers.cpp:717)
            ks = int(2.0 * 3.0 * s + 1.0)
            ks += 1 - ks % 2
            out["blur%d_%d" % (i, j)] = cv2.GaussianBlur(img, (ks, ks), s, borderType=cv2.BORDER_REPLICATE)
        out["half%d" % i] = cv2.resize(img, (0, 0), fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
    out["sigmas"] = np.array(sigmas, np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_pins.npz"), **out)
    print("cv2_pins.npz", len(out))


def _ref_models():
    import torch
    import torch.nn as nn

    def trunk(c):
        return [nn.Conv2d(1, c, 3, padding=1, bias=False), nn.BatchNorm2d(c, affine=False), nn.ReLU(),
                nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c, affine=False), nn.ReLU(),
                nn.Conv2d(c, 2 * c, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(2 * c, affine=False), nn.ReLU(),
                nn.Conv2d(2 * c, 2 * c, 3, padding=1, bias=False), nn.BatchNorm2d(2 * c, affine=False), nn.ReLU(),
                nn.Conv2d(2 * c, 4 * c, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(4 * c, affine=False), nn.ReLU(),
                nn.Conv2d(4 * c, 4 * c, 3, padding=1, bias=False), nn.BatchNorm2d(4 * c, affine=False), nn.ReLU()]

    class Net(nn.Module):
        def __init__(self, layers):
            super().__init__()
            self.features = nn.Sequential(*layers)

    hard = Net(trunk(32) + [nn.Dropout(0.3), nn.Conv2d(128, 128, 8, bias=False), nn.BatchNorm2d(128, affine=False)])
    aff = Net(trunk(16) + [nn.Dropout(0.25), nn.Conv2d(64, 3, 8, bias=True), nn.Tanh(), nn.AdaptiveAvgPool2d(1)])
    ori = Net(trunk(16) + [nn.Dropout(0.25), nn.Conv2d(64, 2, 8, padding=1, bias=True), nn.Tanh(), nn.AdaptiveAvgPool2d(1)])
    for m, f in ((hard, "HardNet++.pth"), (aff, "AffNet.pth"), (ori, "OriNet.pth")):
        ck = torch.load(os.path.join(REF, "build", f), map_location="cpu", weights_only=False)
        m.load_state_dict(ck["state_dict"])
        m.eval()
    return hard, aff, ori


def cnn_golden():
    import torch
    from mods_light_zmq_b200 import synth
    from oracle import pyoracle as O
    a = synth.blob_image(seed=99, w=320, h=240, n_blobs=300)
    g = O.gray_from_bgr(synth.gray_to_bgr(a))
    k = O.detect_hessian(g)
    regs = O.regions_from_keypoints(k[:46])
    patches = O.quantize_u8(O.extract_patches(g, regs))
    extra = np.stack([np.full((32, 32), 77, np.uint8),                         # constant patch: std = 0
                      (np.arange(1024).reshape(32, 32) % 256).astype(np.uint8)])
    patches = np.concatenate([patches, extra])
    hard, aff, ori = _ref_models()
    x = torch.from_numpy(patches.astype(np.float32)).unsqueeze(1)

    def norm(x):
        flat = x.view(x.size(0), -1)
        return (x - flat.mean(dim=1).view(-1, 1, 1, 1)) / (flat.std(dim=1) + 1e-7).view(-1, 1, 1, 1)

    with torch.no_grad():
        f = hard.features(norm(x)).view(len(x), -1)
        d = f / torch.sqrt(torch.sum(f * f, dim=1) + 1e-10).unsqueeze(-1)
        desc_raw = d.numpy()
        desc = np.clip(210 * (desc_raw.astype(np.float64) + 0.45), 0, 255).astype(np.uint8).astype(np.float32)
        af = aff.features(norm(x)).view(-1, 3).clone()
        af[:, 0] += 1
        af[:, 2] += 1
        orr = ori.features(norm(x)).view(-1, 2)
    np.savez_compressed(os.path.join(HERE, "cnn_golden.npz"), patches=patches, hardnet_raw=desc_raw, hardnet=desc,
                        affnet=af.numpy(), orinet=orr.numpy())
    print("cnn_golden.npz", patches.shape)


def graf_counts():
    import cv2
    from oracle import pyoracle as O
    from oracle import cnn_oracle as CN
    res = {"readme": {"regions": [3731, 4527], "descriptors": [3358, 4118]}, "oracle": {}}
    for name in ("graf1", "graf6"):
        bgr = cv2.imread(os.path.join(REF, "build", "imgs", name + ".png"), cv2.IMREAD_COLOR)
        g = O.gray_from_bgr(bgr)
        h, w = g.shape
        k = O.detect_hessian(g)
        regs = O.regions_from_keypoints(k)
        aff = CN.affnet(O.quantize_u8(O.extract_patches(g, regs)))
        r2, _ = O.affnet_postprocess(regs, aff, w, h)
        ori = CN.orinet(O.quantize_u8(O.extract_patches(g, r2)))
        r3 = O.orinet_postprocess(r2, ori)
        r4, _ = O.reproject_filter(r3, w, h)
        res["oracle"][name] = {"keypoints": int(len(k)), "regions": int(len(r2)), "descriptors": int(len(r4))}
        print(name, res["oracle"][name])
    # classic configuration (config_affori_classic.ini + iters_HessianSIFT.ini; README.md:77-105)
    res["readme_classic"] = {"regions": [2665, 3287], "descriptors": [2331, 2912], "tentatives": 76, "unique": 74, "inliers": 21}
    res["oracle_classic"] = {}
    cl = {}
    for name in ("graf1", "graf6"):
        g = O.gray_from_bgr(cv2.imread(os.path.join(REF, "build", "imgs", name + ".png"), cv2.IMREAD_COLOR))
        nk, r, d = O.classic_regions(g)
        cl[name] = (r, d)
        res["oracle_classic"][name] = {"keypoints": int(nk), "descriptors": int(len(r))}
    xy1, xy6 = np.c_[cl["graf1"][0]["x"], cl["graf1"][0]["y"]], np.c_[cl["graf6"][0]["x"], cl["graf6"][0]["y"]]
    m = O.match_fginn(cl["graf1"][1], xy1, cl["graf6"][1], xy6)
    keep = O.duplicate_filter(xy1[m["qi"]], xy6[m["ti"]], m["ratio"], 2.0)
    u = np.c_[xy1[m["qi"]][keep], np.ones(len(keep)), xy6[m["ti"]][keep], np.ones(len(keep))]
    rr = O.ref_ransac_H(np.ascontiguousarray(u), th=16.0)
    res["oracle_classic"]["pair"] = {"tentatives": int(len(m)), "unique": int(len(keep)), "inliers_ref_degensac": int(rr["I"])}
    print("classic", res["oracle_classic"])
    json.dump(res, open(os.path.join(HERE, "graf_counts.json"), "w"), indent=1)


def ransac_ref():
    from oracle import pyoracle as O
    rng = np.random.RandomState(5)
    T, n_in = 260, 150
    Ht = np.array([[0.9, -0.3, 40.0], [0.25, 1.05, -20.0], [2e-4, 1e-5, 1.0]])
    x1 = np.c_[rng.uniform(20, 1000, T), rng.uniform(20, 740, T), np.ones(T)]
    p = x1 @ Ht.T
    x2 = p / p[:, 2:3]
    x2[:, :2] += rng.normal(0, 0.7, (T, 2))
    x2[n_in:, :2] = np.c_[rng.uniform(0, 1024, T - n_in), rng.uniform(0, 768, T - n_in)]
    u = np.ascontiguousarray(np.c_[x1, x2])
    r = O.ref_ransac_H(u, th=16.0, seed_time=12345)
    np.savez_compressed(os.path.join(HERE, "ransac_ref.npz"), u=u, H=r["H"], inl=r["inl"], I=r["I"], J=r["J"],
                        samples=r["samples"], lo_count=r["lo_count"], Htrue=Ht)
    print("ransac_ref.npz I=%d J=%.3f samples=%d lo=%d" % (r["I"], r["J"], r["samples"], r["lo_count"]))


def ransac_F_ref():
    """Seeded two-view scene + the result of the reference's own exp_ransacFcustom (oracle/_ref, time() pinned)."""
    from oracle import pyoracle as O
    from mods_light_zmq_b200 import synth
    u, F_true, mask = synth.two_view_correspondences(21, 300, 170)
    r = O.ref_ransac_F(u, th=16.0, seed_time=12345)
    np.savez_compressed(os.path.join(HERE, "ransac_F_ref.npz"), u=u, F=r["F"], inl=r["inl"], I=r["I"], mask=mask,
                        F_true=F_true, samples=r["samples"], lo=r["lo_count"])
    print("ransac_F_ref.npz I=%d samples=%d lo=%d true inliers found %d" %
          (r["I"], r["samples"], r["lo_count"], int((r["inl"].astype(bool) & mask).sum())))


SYNTH_CASES = [(2, 0.0, 1, 0.2), (2, 0.7, 1, 0.2), (4, 2.1, 1, 0.2), (8, 3.0, 1, 0.2), (6, 1.2, 1, 0.8), (1, 0.0, 0.25, 0.8),
               (3, 0.5, 0.25, 0.8), (-2, 0.9, 1, 0.2), (1, 0.0, 0.125, 0.8), (1, 0, 1, 0.2), (2, 1.5707963267948966, 1, 0.5)]


def cv_synth(gray, tilt, phi, zoom, InitSigma, doBlur=1):
    """Literal transcription of GenerateSynthImageCorr (synth-detection.cpp:324-518) on top of cv2."""
    import cv2
    vertical = False
    if tilt < 0:
        tilt, vertical = -tilt, True
    zoomed = 1 if abs(np.float32(zoom - 1.0)) >= 0.05 else 0
    h, w = gray.shape
    wS1, hS1 = int(w * zoom), int(h * zoom)
    if abs(tilt - 1.) <= 0.1 and abs(int(phi)) <= 0.2 and abs(zoom - 1.) <= 0.1:
        return gray.copy()
    kV = kH = 1.
    if zoomed:
        kV, kH = w / wS1, h / hS1
    c, s = np.cos(phi), np.sin(phi)
    tx, ty = (kH, tilt * kV) if vertical else (tilt * kH, kV)
    if 0 <= phi < np.pi / 2:
        w_new, h_new = np.floor((0.5 + c * w + s * h) / tx), np.floor((0.5 + s * w + c * h) / ty)
        wr, hr = int(np.floor(0.5 + c * w + s * h)), int(np.floor(0.5 + s * w + c * h))
        R = [c, s, 0, -s, c, np.floor(0.5 + s * w)]
    else:
        w_new, h_new = np.floor((0.5 - c * w + s * h) / tx), np.floor((0.5 + s * w - c * h) / ty)
        wr, hr = int(np.floor(0.5 - c * w + s * h)), int(np.floor(0.5 + s * w - c * h))
        R = [c, s, -np.floor(c * w), -s, c, np.floor(0.5 + (s * w - c * h))]
    sa2 = InitSigma / (4.0 * zoom) if zoomed else InitSigma / 2.0
    sa = InitSigma * tilt / (2.0 * zoom)
    sx, sy = (sa2, sa) if vertical else (sa, sa2)
    t = cv2.warpAffine(gray, np.array(R).reshape(2, 3), (wr, hr), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT,
                       borderValue=(128, 128, 128))
    if doBlur:
        kx = int(np.floor(6 * sx + 1))
        kx += (kx % 2 == 0)
        kx = max(kx, 3)
        ky = int(np.floor(6 * sy + 1))
        ky += (ky % 2 == 0)
        ky = max(ky, 3)
        t = cv2.GaussianBlur(t, (kx, ky), sigmaX=sx, sigmaY=sy)
    Wm = np.array([1 / kH, 0, 0, 0, 1 / (tilt * kV), 0]) if vertical else np.array([1 / (tilt * kH), 0, 0, 0, 1 / kV, 0])
    return cv2.warpAffine(t, Wm.reshape(2, 3), (int(w_new), int(h_new)), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=(128, 128, 128))


def synth_pins():
    """cv2 outputs of the view-synthesis chain on a small seeded image (the OpenCV calls whose arithmetic lives
    outside the reference tree: warpAffine synth-detection.cpp:473,:514 and the anisotropic GaussianBlur :499)."""
    import cv2
    rng = np.random.RandomState(5)
    img = cv2.GaussianBlur((rng.rand(45, 62) * 255).astype(np.float32), (5, 5), 1.0)
    out = {"img": img, "cases": np.array(SYNTH_CASES, np.float64)}
    for i, (tilt, phi, zoom, isg) in enumerate(SYNTH_CASES):
        out["view%d" % i] = cv_synth(img, tilt, phi, zoom, isg)
    # the two primitives on their own (ragged sizes, 3-tap kernels, odd widths)
    out["blur_a"] = cv2.GaussianBlur(img[:, :61], (3, 5), sigmaX=0.4, sigmaY=0.8)
    out["blur_b"] = cv2.GaussianBlur(img[:37, :50], (7, 3), sigmaX=1.2, sigmaY=0.1)
    np.savez_compressed(os.path.join(HERE, "synth_pins.npz"), **out)
    print("synth_pins.npz", len(SYNTH_CASES), "views")


def oxaff_golden():
    """Seeded regions + the OxAff line values computed the way saveKP_KM_format does (imagerepresentation.cpp:113-126)
    with cv2.SVDecomp standing in for cv::SVD (same library, float)."""
    import cv2
    rng = np.random.RandomState(11)
    n = 40
    regs = np.zeros(n, [("x", "f8"), ("y", "f8"), ("s", "f8"), ("a11", "f8"), ("a12", "f8"), ("a21", "f8"), ("a22", "f8")])
    regs["x"], regs["y"] = rng.uniform(5, 1000, n), rng.uniform(5, 700, n)
    regs["s"] = np.exp(rng.uniform(np.log(1.5), np.log(40), n))
    ang = rng.uniform(0, 2 * np.pi, n)
    l = np.exp(rng.uniform(-0.8, 0.8, n))
    for i in range(n):
        R = np.array([[np.cos(ang[i]), -np.sin(ang[i])], [np.sin(ang[i]), np.cos(ang[i])]])
        A = R @ np.diag([l[i], 1 / l[i]]) @ R.T @ np.array([[np.cos(ang[i] / 3), np.sin(ang[i] / 3)], [-np.sin(ang[i] / 3), np.cos(ang[i] / 3)]])
        regs["a11"][i], regs["a12"][i], regs["a21"][i], regs["a22"][i] = A.ravel()
    abc = np.zeros((n, 3), np.float32)
    for i in range(n):
        a11, a12, a21, a22 = regs["a11"][i], regs["a12"][i], regs["a21"][i], regs["a22"][i]
        sc = regs["s"][i] * np.sqrt(abs(a11 * a22 - a12 * a21)) * 3.0 * np.sqrt(3.0)
        det = np.sqrt(abs(a11 * a22 - a12 * a21))
        b2a2 = np.sqrt(a12 * a12 + a11 * a11)
        A = np.array([[b2a2 / det, 0], [(a22 * a12 + a21 * a11) / (b2a2 * det), det / b2a2]], np.float32)
        w, u, vt = cv2.SVDecomp(A, flags=cv2.SVD_FULL_UV)
        d = w.ravel().astype(np.float32)
        d0 = np.float32(np.float32(1.0) / (np.float64(d[0] * d[0]) * sc * sc))
        d1 = np.float32(np.float32(1.0) / (np.float64(d[1] * d[1]) * sc * sc))
        E = (u @ np.diag(np.array([d0, d1], np.float32)) @ u.T).astype(np.float32)
        abc[i] = E[0, 0], E[0, 1], E[1, 1]
    np.savez_compressed(os.path.join(HERE, "oxaff_golden.npz"), regs=regs, abc=abc)
    print("oxaff_golden.npz", n)


def flann_pins():
    """The matcher's third-party call: cv::flann::Index with cvflann::LinearIndexParams + knnSearch (matching.cpp:394-415,
    vector_matcher = linear), taken from the cv2 4.13 wheel's flann_Index(algorithm = FLANN_INDEX_LINEAR).  Byte-valued
    128-D descriptors with exact duplicates among the train rows and query rows equal to train rows, so that the
    tie order (equal distances keep train-index order) and zero distances are part of the fixture."""
    import cv2
    rng = np.random.RandomState(7)
    t = rng.randint(0, 256, (700, 128)).astype(np.float32)
    t[50] = t[10]; t[200] = t[10]; t[699] = t[3]; t[350:360] = t[20]
    q = rng.randint(0, 256, (90, 128)).astype(np.float32)
    q[0] = t[10]; q[1] = t[3] + 1; q[2] = t[20]; q[3] = 0; q[4] = 255
    idx, dist = cv2.flann_Index(t, dict(algorithm=0)).knnSearch(q, 50, params={})
    # a low-dimensional case with many equal distances (integer lattice points)
    t2 = np.stack(np.meshgrid(np.arange(12), np.arange(12)), -1).reshape(-1, 2).astype(np.float32)
    q2 = np.array([[5.5, 5.5], [0, 0], [11, 3], [6, 6]], np.float32)
    idx2, dist2 = cv2.flann_Index(t2, dict(algorithm=0)).knnSearch(q2, 50, params={})
    np.savez_compressed(os.path.join(HERE, "flann_pins.npz"), q=q.astype(np.uint8), t=t.astype(np.uint8), idx=idx, dist=dist,
                        q2=q2, t2=t2, idx2=idx2, dist2=dist2, cv2_version=cv2.__version__)
    print("flann_pins.npz", idx.shape, idx2.shape)


def imencode_pins():
    """What the daemons receive: DescribeWithZmq PNG-encodes the FLOAT patch column (imagerepresentation.cpp:45);
    cv::imencode falls back to convertTo(CV_8U) = saturate_cast<uchar>(cvRound(v)): round half to even, clamp to 0..255."""
    import cv2
    rng = np.random.RandomState(3)
    a = rng.uniform(-20, 280, (96, 32)).astype(np.float32)
    a[0, :8] = [0.5, 1.5, 2.5, 3.5, 254.5, 255.5, -0.5, 127.5]
    a[1, :4] = [np.float32(2.5) - np.float32(1e-6), np.float32(2.5) + np.float32(1e-6), 255.49999, 256.0]
    ok, buf = cv2.imencode(".png", a)
    dec = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
    assert ok and dec.dtype == np.uint8
    np.savez_compressed(os.path.join(HERE, "imencode_pins.npz"), patches=a, u8=dec, cv2_version=cv2.__version__)
    print("imencode_pins.npz", dec.shape)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cv2", "cnn", "graf", "ransac", "ransacF", "oxaff", "synth", "flann", "imencode"]
    if "imencode" in which:
        imencode_pins()
    if "flann" in which:
        flann_pins()
    if "cv2" in which:
        cv2_pins()
    if "cnn" in which:
        cnn_golden()
    if "graf" in which:
        graf_counts()
    if "ransac" in which:
        ransac_ref()
    if "synth" in which:
        synth_pins()
    if "oxaff" in which:
        oxaff_golden()
    if "ransacF" in which:
        ransac_F_ref()
