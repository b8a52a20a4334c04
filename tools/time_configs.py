"""Supplementary timings of the other BASELINE configurations on one GPU (the bench line is config 3):
  config 2  detect-only (HessianAffine, 1024x768)              images/s
  config 4  extract_features (detect + 3 nets + filters)       images/s   (what batch.py runs per image)
  config 5  MODS loop on a tilted 1024x768 pair, LORANSACF     ms per pair
Threads = contexts on one GPU, as in bench.py.  Wall-clock over the whole batch (host API calls, H2D/D2H included)."""
import os, sys, time, threading, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import synth

NW = int(sys.argv[1]) if len(sys.argv) > 1 else 8
imgs = [synth.gray_to_bgr(synth.blob_image(seed=s)) for s in range(8)]
mgs = [M.ModsGpu(0, load_nets=True) for _ in range(NW)]


def run(fn, n_items):
    def worker(w):
        for i in range(w, n_items, NW):
            fn(mgs[w], i)
    ths = [threading.Thread(target=worker, args=(w,)) for w in range(NW)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    return time.perf_counter() - t0


def detect_only(mg, i):
    img = mg.image_from_bgr8(imgs[i % len(imgs)])
    n = len(mg.detect(img))
    img.free()
    return n


def extract(mg, i):
    img = mg.image_from_bgr8(imgs[i % len(imgs)])
    n = len(mg.extract_features(img))
    img.free()
    return n


out = {"workers": NW}
for name, fn, n in (("config2_detect_only_images_per_s", detect_only, 256), ("config4_extract_features_images_per_s", extract, 128)):
    run(fn, 2 * NW)
    dt = run(fn, n)
    out[name] = n / dt
a = synth.blob_image(seed=91)
Ht = np.array([[0.34, 0.06, 100.0], [-0.02, 0.97, 10.0], [0.0, 0.0, 1.0]])
b = synth.warp_image(a, Ht, noise_seed=5)
steps = [dict(tilts=[1.0], phi=360.0), dict(tilts=[1.0, 2.0, 4.0], phi=360.0), dict(tilts=[1.0, 2.0, 4.0, 6.0, 8.0], phi=360.0)]


def mods(mg, i):
    i1, i2 = mg.image_from_bgr8(synth.gray_to_bgr(a)), mg.image_from_bgr8(synth.gray_to_bgr(b))
    r = mg.mods_pair(i1, i2, steps, min_matches=100000, use_F=True, seed=3 + i)
    i1.free(); i2.free()
    return r


r = mods(mgs[0], 0)
t0 = time.perf_counter(); r = mods(mgs[0], 1); t1 = time.perf_counter() - t0
out["config5_mods_pair_ms_single_context"] = 1e3 * t1
out["config5_views"] = r["views"]; out["config5_regions"] = r["regions"]; out["config5_inliers"] = r["inliers"]
dt = run(mods, 2 * NW)
out["config5_mods_pairs_per_s"] = 2 * NW / dt
print(json.dumps(out))
