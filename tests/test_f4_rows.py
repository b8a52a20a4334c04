"""SURVEY 8(f) rank-4 rows built this round: MatchFLANNDistance (Hamming 2-NN over binary descriptors, matching.cpp:574-633)
and HMatrixFiltering (verification against a known homography, :917-1013)."""
import ctypes as C
import os

import numpy as np
import pytest


def _binary_desc(rng, n, dim=32, twin=None, flips=12):
    d = rng.randint(0, 256, (n, dim)).astype(np.uint8)
    if twin is not None:
        k = min(n, len(twin)) // 2
        bits = np.unpackbits(twin[:k], axis=1)
        for r in range(k):
            bits[r, rng.choice(dim * 8, flips, replace=False)] ^= 1
        d[:k] = np.packbits(bits, axis=1)
    return d


def test_oracle_hamming_2nn_matches_cv2_bruteforce(oracle):
    """The oracle's exact Hamming 2-NN against cv2.BFMatcher(NORM_HAMMING).knnMatch: both distances always, the nearest
    index wherever the minimum is unique (BFMatcher's tie order is its own)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(1)
    t = _binary_desc(rng, 400)
    q = _binary_desc(rng, 300, twin=t)
    t[350:360] = t[:10]                                        # exact duplicates: distance ties
    m = oracle.match_hamming(q.astype(np.float32) + 0.4, t.astype(np.float32) + 0.7, 1000)     # floor() recovers the bytes
    assert len(m) == len(q)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    d1 = np.array([k[0].distance for k in knn]); d2 = np.array([k[1].distance for k in knn])
    i1 = np.array([k[0].trainIdx for k in knn])
    assert np.array_equal(m["d1"], d1) and np.array_equal(m["d2"], d2)
    uniq = d1 < d2
    assert uniq.sum() > 200 and np.array_equal(m["ti"][uniq], i1[uniq])
    # duplicates: the lower index wins
    dup = np.isin(m["ti"], np.arange(10)) & (m["d2"] == m["d1"])
    assert (m["tj_bad"][dup] >= 350).all()
    # the distance threshold
    m40 = oracle.match_hamming(q, t, 40.9)
    assert 0 < len(m40) < len(q) and (m40["d1"] <= 40).all() and np.allclose(m40["ratio"], m40["d1"] / m40["d2"])


@pytest.mark.gpu
@pytest.mark.parametrize("nq,nt,dim", [(300, 400, 32), (5, 1, 32), (129, 1000, 64), (1, 3, 4)])
def test_match_hamming_bit_exact(mg, oracle, nq, nt, dim):
    import mods_light_zmq_b200 as M
    rng = np.random.RandomState(nq + nt)
    t = _binary_desc(rng, nt, dim)
    q = _binary_desc(rng, nq, dim, twin=t, flips=max(1, dim // 3))
    if nt > 20:
        t[nt - 5:] = t[:5]
    qf, tf = q.astype(np.float32) + 0.25, t.astype(np.float32)
    for thr in (1000.0, dim * 8 * 0.2):
        ref = oracle.match_hamming(qf, tf, thr)
        got = mg.match_hamming(qf, tf, thr)
        assert got.tobytes() == ref.tobytes(), (thr, len(got), len(ref))
    assert len(mg.match_hamming(qf[:0], tf, 10)) == 0 and len(mg.match_hamming(qf, tf[:0], 10)) == 0
    with pytest.raises(M.ModsGpuError):
        mg.match_hamming(np.zeros((2, 6), np.float32), np.zeros((2, 6), np.float32), 10)


@pytest.mark.parametrize("error,etype", [("sampson", 0), ("symm_max", 1), ("symm_sum", 2)])
def test_hmatrix_filter_vs_reference_functions(oracle, error, etype):
    """HMatrixFiltering of the host mirror against the reference's own HDs / HDsSymMax / HDsSym (oracle/_ref) on the
    (image 2, image 1) packing the reference uses: keep masks equal for every error type, transposed H returned."""
    import mods_light_zmq_b200 as M
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    lib = M.load_library()
    rng = np.random.RandomState(3 + etype)
    n = 500
    Ht = np.array([[0.9, -0.3, 40.0], [0.25, 1.05, -20.0], [2e-4, 1e-5, 1.0]])
    xy1 = np.c_[rng.uniform(20, 1000, n), rng.uniform(20, 740, n)]
    p = np.c_[xy1, np.ones(n)] @ Ht.T
    xy2 = p[:, :2] / p[:, 2:3] + rng.normal(0, 1.0, (n, 2)) * np.linspace(0.1, 6.0, n)[:, None]   # growing noise
    # H in the degensac layout for u = (image 2, image 1): column-major matrix mapping image 1 -> image 2
    Hd = np.ascontiguousarray(Ht.T.ravel())
    for thr in (1.0, 4.0):
        keep = np.zeros(n, np.uint8)
        Hout = np.zeros(9)
        m = C.c_int()
        rc = lib.modsgpu_hmatrix_filter(xy1.ctypes.data_as(C.c_void_p), xy2.ctypes.data_as(C.c_void_p), n, Hd.ctypes.data_as(C.c_void_p),
                                        etype, C.c_double(thr), keep.ctypes.data_as(C.c_void_p), Hout.ctypes.data_as(C.c_void_p), C.byref(m))
        assert rc == 0
        okeep, d = oracle.ref_hmatrix_filter(xy1, xy2, Hd, error, thr)
        assert np.array_equal(keep.astype(bool), okeep), (thr, int(keep.sum()), int(okeep.sum()))
        assert m.value == int(okeep.sum()) and 20 < m.value < n
        assert np.array_equal(Hout.reshape(3, 3), Hd.reshape(3, 3).T)
    assert lib.modsgpu_hmatrix_filter(None, None, 0, Hd.ctypes.data_as(C.c_void_p), 0, C.c_double(4.0), None, None, C.byref(m)) == 0 and m.value == 0
