"""Short target for ncu: one warm-up pair + one profiled pair through the C ABI."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import synth
mg = M.ModsGpu(0, load_nets=True)
a, b, H = synth.image_pair()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for k in range(n):
    r = mg.pair_pipeline(synth.gray_to_bgr(a), synth.gray_to_bgr(b), seed=5)
print({k: r[k] for k in ("keypoints", "descriptors", "tentatives", "inliers")}, "launches", mg.launch_count)
