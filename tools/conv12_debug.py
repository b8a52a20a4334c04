"""Stall table of the fused conv1+conv2 kernel (MODSGPU_CONV12_DEBUG=1): prints per-role wait cycles per tile."""
import os, sys
os.environ["MODSGPU_CONV12_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mods_light_zmq_b200 as M
g = M.ModsGpu(0, load_nets=True)
rng = np.random.RandomState(0)
p = rng.randint(0, 256, (4608, 32, 32)).astype(np.uint8)
for net in (M.AFFNET, M.HARDNET):
    for _ in range(2):
        g.net_forward_u8(net, p)
