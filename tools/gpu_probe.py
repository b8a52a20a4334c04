"""Scratch timing of every stage of the pair pipeline on the GPU box (device ms via CUDA events)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import synth
from oracle import pyoracle as O

mg = M.ModsGpu(0, load_nets=True)
a, b, H = synth.image_pair()
res = {}
def timed(name, fn, reps=5):
    fn(); ms = []; wall = []
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); wall.append((time.perf_counter() - t) * 1e3); ms.append(mg.last_device_ms)
    res[name] = dict(dev_ms=min(ms), wall_ms=min(wall)); print(name, res[name], flush=True)
    return r
bgr = synth.gray_to_bgr(a)
img = timed("upload+gray", lambda: mg.image_from_bgr8(bgr))
kp = timed("detect", lambda: mg.detect(img))
print("keypoints", len(kp))
regs = M.regions_from_keypoints(kp)
timed("sampler", lambda: mg.extract_patches(img, regs))
aff = timed("describe_affnet", lambda: mg.describe(M.AFFNET, img, regs))
h, w = a.shape
r2, _ = O.affnet_postprocess(regs, aff, w, h)
ori = timed("describe_orinet", lambda: mg.describe(M.ORINET, img, r2))
r3 = O.orinet_postprocess(r2, ori)
r4, _ = O.reproject_filter(r3, w, h)
d = timed("describe_hardnet", lambda: mg.describe(M.HARDNET, img, r4))
p = mg.extract_patches(img, r4)
timed("hardnet_only", lambda: mg.net_forward_u8(M.HARDNET, p))
timed("affnet_only", lambda: mg.net_forward_u8(M.AFFNET, p))
xy = np.c_[r4["x"], r4["y"]]
m = timed("match", lambda: mg.match_fginn(d, d[::-1].copy(), xy))
print("n desc", len(d), "launches", mg.launch_count)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
