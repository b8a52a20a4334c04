"""LO-RANSAC parity.

Three layers of evidence:
 1. oracle/ransac_batched.c restates the device's batched schedule sequentially; the device must return its inlier
    mask BYTE for byte and its H bit for bit (GPU test).
 2. that schedule against the REFERENCE's own exp_ransacHcustom (oracle/_ref, time() pinned) over 200 seeded sets:
    the rate of byte-equal inlier masks is measured and asserted (CPU test -- by 1. it is the device's rate too).
 3. the empirical checks LORANSACFiltering applies afterwards (NaiveHCheck, H_LAF_check, F_LAF_check) in the host
    mirror against the reference's own HDsSymMax / FDs (CPU test through the C ABI, no GPU needed).
"""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _corr_set(seed, T, n_in, noise=0.7):
    rng = np.random.RandomState(seed)
    Ht = np.array([[0.9, -0.3, 40.0], [0.25, 1.05, -20.0], [2e-4, 1e-5, 1.0]])
    x1 = np.c_[rng.uniform(20, 1000, T), rng.uniform(20, 740, T), np.ones(T)]
    p = x1 @ Ht.T
    x2 = p / p[:, 2:3]
    x2[:, :2] += rng.normal(0, noise, (T, 2))
    x2[n_in:, :2] = np.c_[rng.uniform(0, 1024, T - n_in), rng.uniform(0, 768, T - n_in)]
    return np.ascontiguousarray(np.c_[x1, x2]), Ht


def _seeded_set(seed, noise):
    rng = np.random.RandomState(seed)
    T = int(rng.choice([60, 150, 300, 600, 1200]))
    n_in = max(12, int(T * rng.uniform(0.3, 0.9)))
    return _corr_set(seed, T, n_in, noise)[0]


# ------------------------------------------------------------------------------------------ 2. schedule vs reference
@pytest.mark.parametrize("noise,min_rate,min_jaccard", [(0.7, 0.99, 0.999), (1.5, 0.98, 0.99), (2.5, 0.70, 0.90)])
def test_batched_schedule_mask_equality_rate_vs_reference(oracle, noise, min_rate, min_jaccard):
    """200 seeded correspondence sets (60..1200 tentatives, 30-90 % inliers): how often is the batched schedule's final
    inlier mask byte-equal to the mask of the reference's exp_ransacHcustom?  Measured in this container: 200/200 at
    0.7 px and 1.5 px noise, 170/200 at 2.5 px (threshold 4 px: dozens of correspondences sit on the threshold and the
    two local optimisations stop at models a fraction of a pixel apart; |dI| <= 2, Jaccard >= 0.95).  The reference binary
    is not run-to-run deterministic at this noise level (169..173 equal masks, min Jaccard 0.95..0.98 over repeated runs of
    this very test), hence the margin on the last row."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    eq, jac, dI = 0, [], []
    for seed in range(200):
        u = _seeded_set(seed, noise)
        g = oracle.batched_ransac_H(u, seed=1000 + seed)
        r = oracle.ref_ransac_H(u, th=16.0, seed_time=12345)
        a, b = g["inl"].astype(bool), r["inl"].astype(bool)
        eq += int(np.array_equal(a, b))
        jac.append((a & b).sum() / max((a | b).sum(), 1))
        dI.append(g["I"] - r["I"])
    print("noise %.1f: byte-equal masks %d/200, min Jaccard %.4f, dI in [%d, %d]" % (noise, eq, min(jac), min(dI), max(dI)))
    assert eq / 200.0 >= min_rate, eq
    assert min(jac) >= min_jaccard and min(dI) >= -3


def test_batched_schedule_error_types_and_edges(oracle):
    u, _ = _corr_set(3, 300, 180)
    base = oracle.batched_ransac_H(u, seed=5)
    assert base["I"] == int(base["inl"].sum()) >= 175 and base["lo_count"] >= 1 and base["samples"] >= 512
    assert np.array_equal(base["inl"], (base["resid"] <= 16.0).astype(np.uint8))
    for et in (1, 2):     # SymmMax, SymmSum: their own error functions, same geometry recovered
        r = oracle.batched_ransac_H(u, seed=5, error_type=et)
        assert r["inl"][:180].mean() > 0.9 and r["inl"][180:].sum() <= 6
        if oracle.ref_available():
            rr = oracle.ref_ransac_H(u, th=16.0, error={1: "symm_max", 2: "symm_sum"}[et])
            a, b = r["inl"].astype(bool), rr["inl"].astype(bool)
            assert (a & b).sum() / max((a | b).sum(), 1) >= 0.97
    assert oracle.batched_ransac_H(u[:3])["I"] == 0


# ------------------------------------------------------------------------------------------ 1. device == schedule
@pytest.mark.gpu
def test_device_ransac_H_byte_equal_to_batched_oracle(mg, oracle):
    cases = [(s, n) for s in range(24) for n in (0.7, 2.5)]
    n_bit = 0
    for seed, noise in cases:
        u = _seeded_set(seed, noise)
        et = seed % 3
        sym = 0 if seed % 5 == 4 else 1
        g = mg.ransac_H(u, seed=1000 + seed, error_type=et, sym_check=sym)
        o = oracle.batched_ransac_H(u, seed=1000 + seed, error_type=et, sym_check=sym)
        assert np.array_equal(g["inl"], o["inl"]), (seed, noise, int(g["inl"].sum()), int(o["inl"].sum()))
        assert np.allclose(g["H"], o["H"], rtol=1e-9, atol=0), (seed, noise)
        assert (g["I"], g["samples"], g["lo_count"], g["oc_rejects"]) == (o["I"], o["samples"], o["lo_count"], o["oc_rejects"])
        assert np.allclose(g["resid"], o["resid"], rtol=1e-9, atol=1e-300)
        n_bit += int(np.array_equal(g["H"], o["H"]) and g["J"] == o["J"])
    print("device H bit-identical to the CPU restatement in %d/%d cases" % (n_bit, len(cases)))
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "ransac_H_bit_identity.txt"), "w") as f:
        f.write("device LO-RANSAC(H): inlier mask byte-equal in %d/%d cases, H and J bit-identical in %d/%d\n" % (len(cases), len(cases), n_bit, len(cases)))
    assert n_bit >= len(cases) - 2
    # the size the bench pair produces (~3000 tentatives, almost all inliers) and a many-batch case (5 % inliers)
    u, _ = _corr_set(99, 3000, 2950)
    g, o = mg.ransac_H(u, seed=7), oracle.batched_ransac_H(u, seed=7)
    assert np.array_equal(g["inl"], o["inl"]) and np.allclose(g["H"], o["H"], rtol=1e-9, atol=0) and g["samples"] == o["samples"]
    u, _ = _corr_set(98, 800, 40)
    g, o = mg.ransac_H(u, seed=8), oracle.batched_ransac_H(u, seed=8)
    assert o["samples"] > 1536 and g["samples"] == o["samples"] and np.array_equal(g["inl"], o["inl"])
    assert np.allclose(g["H"], o["H"], rtol=1e-9, atol=0)


@pytest.mark.gpu
def test_device_ransac_F_mask_equality_rate_vs_reference(mg, oracle):
    """Row a25: the reference's exp_ransacFcustom reads uninitialised heap and is nondeterministic between calls, so
    there is no bit-level target; measured instead over 200 seeded two-view scenes: how often are the masks byte-equal,
    and how far apart are they otherwise."""
    from mods_light_zmq_b200 import synth
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    eq, jac, dI, rec_g, rec_r = 0, [], [], [], []
    for seed in range(200):
        rng = np.random.RandomState(seed)
        T = int(rng.choice([80, 150, 300, 600]))
        n_in = max(20, int(T * rng.uniform(0.4, 0.9)))
        u, _, mask = synth.two_view_correspondences(seed, T, n_in)
        g = mg.ransac_F(u, seed=2000 + seed)
        r = oracle.ref_ransac_F(u, th=16.0, seed_time=12345)
        a, b = g["inl"].astype(bool), r["inl"].astype(bool)
        eq += int(np.array_equal(a, b))
        jac.append((a & b).sum() / max((a | b).sum(), 1))
        dI.append(g["I"] - r["I"])
        rec_g.append((a & mask).sum() / mask.sum())
        rec_r.append((b & mask).sum() / mask.sum())
    jac, dI = np.array(jac), np.array(dI)
    conv = np.array(rec_r) >= 0.95          # scenes on which the reference binary itself found the true model
    msg = ("F over 200 scenes: byte-equal masks %d, Jaccard mean %.4f / 5th pct %.4f / min %.4f, dI mean %.2f in [%d, %d], "
           "recall of the true inliers: ours %.4f, reference %.4f; reference converged on %d scenes, Jaccard mean there %.4f" %
           (eq, jac.mean(), np.percentile(jac, 5), jac.min(), dI.mean(), dI.min(), dI.max(), np.mean(rec_g), np.mean(rec_r),
            int(conv.sum()), jac[conv].mean() if conv.any() else 1.0))
    print(msg)
    with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "ransac_F_rate.txt"), "w") as f:
        f.write(msg + "\n")
    # measured on B200 (profiles/r02_ransac_parity.txt); the two estimators are randomised differently and the
    # reference's LO refits on random 8-subsets, so the masks differ by the borderline correspondences
    # The reference binary is not deterministic (uninitialised heap): its own recall moved between 0.91 and 0.94 over
    # repeated runs of this test and drags the all-scene Jaccard mean with it (0.89 .. 0.92), so the mask agreement is
    # asserted on the scenes where the reference recovered the model, and only loosely over all of them.
    assert conv.sum() >= 100 and jac[conv].mean() >= 0.90
    assert jac.mean() >= 0.80
    assert np.mean(rec_g) >= np.mean(rec_r) - 0.01 and np.mean(rec_g) >= 0.95
    assert dI.mean() >= -0.5        # on average at least the reference's consensus


# ------------------------------------------------------------------------------------------ 3. empirical checks
def _random_lafs(rng, n, xy):
    import mods_light_zmq_b200 as M
    r = np.zeros(n, M.REGION_DTYPE)
    r["x"], r["y"] = xy[:, 0], xy[:, 1]
    r["s"] = rng.uniform(1.5, 12.0, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    st = rng.uniform(0.6, 1.6, n)
    A = np.stack([np.stack([st * np.cos(ang), -np.sin(ang) / st], -1), np.stack([st * np.sin(ang), np.cos(ang) / st], -1)], -2)
    r["a11"], r["a12"], r["a21"], r["a22"] = A[:, 0, 0], A[:, 0, 1], A[:, 1, 0], A[:, 1, 1]
    return r, A


def _checks(kp1, kp2, model, use_F, err_thr, coef):
    import mods_light_zmq_b200 as M
    lib = M.load_library()
    n = len(kp1)
    keep = np.zeros(max(n, 1), np.uint8)
    out = np.zeros(9, np.float64)
    m = C.c_int()
    model = np.ascontiguousarray(model, np.float64)
    rc = lib.modsgpu_empirical_checks(kp1.ctypes.data_as(C.c_void_p), kp2.ctypes.data_as(C.c_void_p), n,
                                      model.ctypes.data_as(C.c_void_p), int(use_F), C.c_double(err_thr), C.c_double(coef),
                                      keep.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(m))
    assert rc == 0
    return keep[:n].astype(bool), out, m.value


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_empirical_checks_H_vs_reference_functions(oracle, seed):
    """NaiveHCheck + H_LAF_check of the host mirror (LORANSACFiltering's tail, matching.cpp:764-805) against the
    oracle that calls the REFERENCE's HDsSymMax: correspondences under a homography whose local affine frames agree
    with it, disagree with it mildly, or are random."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.RandomState(seed)
    n = 400
    Ht = np.array([[0.9, -0.3, 40.0], [0.25, 1.05, -20.0], [2e-4, 1e-5, 1.0]])
    xy1 = np.c_[rng.uniform(50, 950, n), rng.uniform(50, 700, n)]
    p = np.c_[xy1, np.ones(n)] @ Ht.T
    xy2 = p[:, :2] / p[:, 2:3] + rng.normal(0, 0.6, (n, 2))
    kp1, A1 = _random_lafs(rng, n, xy1)
    kp2, _ = _random_lafs(rng, n, xy2)
    # consistent frames for most: A2 = J(H) A1 up to a perturbation that grows along the list
    J = Ht[:2, :2]
    A2 = J @ A1
    pert = np.linspace(0, 1.2, n)[:, None, None] * rng.normal(0, 1.0, (n, 2, 2))
    A2 = A2 * (1 + pert)
    sc = np.sqrt(np.abs(np.linalg.det(A2)))
    kp2["s"] = kp1["s"] * sc
    A2n = A2 / sc[:, None, None]
    kp2["a11"], kp2["a12"], kp2["a21"], kp2["a22"] = A2n[:, 0, 0], A2n[:, 0, 1], A2n[:, 1, 0], A2n[:, 1, 1]
    Hloran = np.linalg.inv(Ht).T.ravel()           # degensac convention: column-major, image 2 -> image 1
    for coef in (12.0, 3.0, 1.0):
        keep, Hout, m = _checks(kp1, kp2, Hloran, 0, 4.0, coef)
        okeep, oH = oracle.empirical_checks(kp1, kp2, Hloran, 0, 4.0, coef)
        assert np.array_equal(keep, okeep), (coef, int(keep.sum()), int(okeep.sum()))
        assert m == int(okeep.sum())
        assert np.allclose(Hout / Hout[8], oH / oH[8], rtol=1e-9)
    assert 8 <= _checks(kp1, kp2, Hloran, 0, 4.0, 1.0)[2] < _checks(kp1, kp2, Hloran, 0, 4.0, 12.0)[2] <= n
    # fewer than 8 survivors / a model that does not explain the points: empty list
    keep, _, m = _checks(kp1[:7], kp2[:7], Hloran, 0, 4.0, 12.0)
    assert m == 0 and not keep.any()
    bad = np.linalg.inv(np.array([[1.0, 0, 300.0], [0, 1.0, 300.0], [0, 0, 1.0]])).T.ravel()
    keep, _, m = _checks(kp1, kp2, bad, 0, 4.0, 12.0)
    okeep, _ = oracle.empirical_checks(kp1, kp2, bad, 0, 4.0, 12.0)
    assert m == 0 and not okeep.any()


@pytest.mark.parametrize("seed", [0, 1])
def test_empirical_checks_F_vs_reference_functions(oracle, seed):
    """F_LAF_check (matching.cpp:192-249, :806-820) against the oracle calling the REFERENCE's FDs."""
    from mods_light_zmq_b200 import synth
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.RandomState(10 + seed)
    u, F_true, mask = synth.two_view_correspondences(seed, 300, 300, noise=0.4)
    n = len(u)
    kp1, _ = _random_lafs(rng, n, u[:, 0:2])
    kp2, _ = _random_lafs(rng, n, u[:, 3:5])
    kp1["s"] *= np.linspace(0.05, 1.0, n)      # small frames stay near the epipolar lines, large ones leave them
    kp2["s"] *= np.linspace(0.05, 1.0, n)
    F = np.ascontiguousarray(F_true, np.float64).ravel()
    for coef in (2.0, 6.0):
        keep, Fout, m = _checks(kp1, kp2, F, 1, 4.0, coef)
        okeep, _ = oracle.empirical_checks(kp1, kp2, F, 1, 4.0, coef)
        assert np.array_equal(keep, okeep) and m == int(okeep.sum()), (coef, int(keep.sum()), int(okeep.sum()))
        assert np.array_equal(Fout, F)
    a, b = _checks(kp1, kp2, F, 1, 4.0, 2.0)[2], _checks(kp1, kp2, F, 1, 4.0, 6.0)[2]
    assert 8 <= a < b <= n
