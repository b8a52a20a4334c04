"""Seeded synthetic inputs for the BASELINE.json configs: a blob image that yields
~4k Hessian keypoints at 1024x768, and a second view of it under a fixed homography."""
import numpy as np

N_BLOBS = 3200  # tuned once with the oracle: ~4.2k / ~3.8k Hessian keypoints for the pair, then frozen


def blob_image(seed=1234, w=1024, h=768, n_blobs=N_BLOBS):
    """Sum of anisotropic Gaussian blobs (sigma log-uniform in [1.5,12] px, amplitude +-[20,80]) on a 128
    background + N(0,2^2) noise, clipped and quantised to u8 (so it survives a PNG round trip)."""
    rng = np.random.RandomState(seed)
    img = np.full((h, w), 128.0, np.float64)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n_blobs):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        s1 = np.exp(rng.uniform(np.log(1.5), np.log(12.0)))
        s2 = s1 * rng.uniform(0.5, 1.0)
        th = rng.uniform(0, np.pi)
        amp = rng.uniform(20, 80) * (1 if rng.rand() < 0.5 else -1)
        R = int(4 * s1) + 1
        x0, x1 = max(0, int(cx) - R), min(w, int(cx) + R + 1)
        y0, y1 = max(0, int(cy) - R), min(h, int(cy) + R + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        dx = xx[y0:y1, x0:x1] - cx
        dy = yy[y0:y1, x0:x1] - cy
        u = np.cos(th) * dx + np.sin(th) * dy
        v = -np.sin(th) * dx + np.cos(th) * dy
        img[y0:y1, x0:x1] += amp * np.exp(-0.5 * ((u / s1) ** 2 + (v / s2) ** 2))
    img += rng.normal(0, 2.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def pair_homography(w=1024, h=768):
    """image A -> image B: rotation 20 deg, scale 0.9 about the centre, perspective 1e-4."""
    a = np.deg2rad(20.0)
    c, s = 0.9 * np.cos(a), 0.9 * np.sin(a)
    T1 = np.array([[1, 0, -w / 2], [0, 1, -h / 2], [0, 0, 1.0]])
    Rm = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    P = np.array([[1, 0, 0], [0, 1, 0], [1e-4, 0, 1.0]])
    T2 = np.array([[1, 0, w / 2], [0, 1, h / 2], [0, 0, 1.0]])
    return T2 @ P @ Rm @ T1


def warp_image(img, H, noise_seed=4321, border=128.0):
    """B(x) = A(H^-1 x), bilinear, constant border, + independent N(0,2^2) noise, u8."""
    h, w = img.shape
    Hi = np.linalg.inv(H)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    d = Hi[2, 0] * xx + Hi[2, 1] * yy + Hi[2, 2]
    sx = (Hi[0, 0] * xx + Hi[0, 1] * yy + Hi[0, 2]) / d
    sy = (Hi[1, 0] * xx + Hi[1, 1] * yy + Hi[1, 2]) / d
    x0 = np.floor(sx).astype(np.int64)
    y0 = np.floor(sy).astype(np.int64)
    fx, fy = sx - x0, sy - y0
    inside = (x0 >= 0) & (y0 >= 0) & (x0 < w - 1) & (y0 < h - 1)
    x0c, y0c = np.clip(x0, 0, w - 2), np.clip(y0, 0, h - 2)
    a = img.astype(np.float64)
    v = (a[y0c, x0c] * (1 - fx) * (1 - fy) + a[y0c, x0c + 1] * fx * (1 - fy) +
         a[y0c + 1, x0c] * (1 - fx) * fy + a[y0c + 1, x0c + 1] * fx * fy)
    v = np.where(inside, v, border)
    rng = np.random.RandomState(noise_seed)
    v = v + rng.normal(0, 2.0, v.shape)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def gray_to_bgr(gray_u8):
    """cv::imread of a gray PNG gives three equal channels."""
    return np.repeat(gray_u8[:, :, None], 3, axis=2)


def image_pair(seed=1234, w=1024, h=768):
    a = blob_image(seed, w, h)
    H = pair_homography(w, h)
    b = warp_image(a, H, noise_seed=seed + 3087)
    return a, b, H


def two_view_correspondences(seed, T, n_in, noise=0.5, w=1024, h=768, planar_fraction=0.0):
    """Seeded general two-view scene for the fundamental-matrix tests: n_in 3-D points (depth 4..12, optionally a
    fraction on one plane) seen by two cameras (focal 900 px; rotation ~12 deg about y, ~4 deg about x, baseline
    (1, 0.1, 0.2)) + Gaussian pixel noise, T - n_in uniformly random mismatches.  Returns (u [T x 6], F_true [3 x 3]
    with x2^T F x1 = 0, inlier mask)."""
    rng = np.random.RandomState(seed)
    K = np.array([[900.0, 0, w / 2], [0, 900.0, h / 2], [0, 0, 1]])
    ay, ax = np.deg2rad(12.0), np.deg2rad(4.0)
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    R = Ry @ Rx
    t = np.array([1.0, 0.1, 0.2])
    X = np.c_[rng.uniform(-4, 4, 4 * T), rng.uniform(-3, 3, 4 * T), rng.uniform(4, 12, 4 * T)]
    on_plane = (np.arange(len(X)) % 100) < int(round(100 * planar_fraction))     # interleaved, so the kept subset mixes both
    X[on_plane, 2] = 8.0 + 0.15 * X[on_plane, 0]
    x1 = (K @ X.T).T
    x1 = x1[:, :2] / x1[:, 2:3]
    Xc = (R @ X.T).T + t
    x2 = (K @ Xc.T).T
    x2 = x2[:, :2] / x2[:, 2:3]
    ok = (x1[:, 0] > 5) & (x1[:, 0] < w - 5) & (x1[:, 1] > 5) & (x1[:, 1] < h - 5) & \
         (x2[:, 0] > 5) & (x2[:, 0] < w - 5) & (x2[:, 1] > 5) & (x2[:, 1] < h - 5) & (Xc[:, 2] > 0.5)
    x1, x2 = x1[ok][:n_in], x2[ok][:n_in]
    assert len(x1) == n_in, "not enough visible points"
    x1 = x1 + rng.normal(0, noise, x1.shape)
    x2 = x2 + rng.normal(0, noise, x2.shape)
    no = T - n_in
    o1 = np.c_[rng.uniform(0, w, no), rng.uniform(0, h, no)]
    o2 = np.c_[rng.uniform(0, w, no), rng.uniform(0, h, no)]
    u = np.ones((T, 6))
    u[:n_in, 0:2], u[:n_in, 3:5] = x1, x2
    u[n_in:, 0:2], u[n_in:, 3:5] = o1, o2
    perm = rng.permutation(T)
    mask = np.zeros(T, bool)
    mask[:n_in] = True
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ki = np.linalg.inv(K)
    F = Ki.T @ tx @ R @ Ki
    return u[perm], F / np.linalg.norm(F), mask[perm]
