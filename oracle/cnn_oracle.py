"""CPU oracle for the three networks (TEST INFRASTRUCTURE ONLY -- see oracle/mods_oracle.h).

Restates the forward passes of the reference daemons in fp32 torch on the CPU:
  HardNet     build/desc_server.py:58-92   (+ the uint8 post-scale of :42)
  AffNetFast  build/affnet_server.py:45-84
  OriNetFast  build/orinet_server.py:45-82
using the BatchNorm-folded weights of weights/*.npz (tools/export_weights.py).  Folding changes the
result by ~1e-6 relative; tests/golden/cnn_golden.npz (made from the original .pth files with the
unfolded nn.Sequential, tests/golden/make_golden.py) pins this restatement.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS = os.path.join(os.path.dirname(HERE), "weights")
STRIDES = [1, 1, 2, 1, 2, 1]
_cache = {}


def _load(name):
    if name not in _cache:
        z = np.load(os.path.join(WEIGHTS, name + ".npz"))
        _cache[name] = {k: torch.from_numpy(z[k]) for k in z.files}
    return _cache[name]


def input_norm(x):
    """(x - mean) / (std_unbiased + 1e-7) per patch (desc_server.py:83-87)."""
    flat = x.reshape(x.shape[0], -1)
    mp = flat.mean(dim=1)
    sp = flat.std(dim=1) + 1e-7
    return (x - mp.view(-1, 1, 1, 1)) / sp.view(-1, 1, 1, 1)


def trunk(name, patches_u8):
    w = _load(name)
    x = torch.from_numpy(np.ascontiguousarray(patches_u8, np.uint8).reshape(-1, 1, 32, 32).astype(np.float32))
    x = input_norm(x)
    for i, st in enumerate(STRIDES, start=1):
        k = w["c%d_w" % i].permute(0, 3, 1, 2).contiguous()      # [Cout,3,3,Cin] -> OIHW
        x = F.relu(F.conv2d(x, k, w["c%d_b" % i], stride=st, padding=1))
    return x, w


def hardnet_raw(patches_u8):
    """L2-normalised 128-d descriptor before quantisation."""
    with torch.no_grad():
        x, w = trunk("hardnet", patches_u8)
        k = w["h_w"].permute(0, 3, 1, 2).contiguous()
        x = F.conv2d(x, k, w["h_b"]).reshape(x.shape[0], -1)
        norm = torch.sqrt(torch.sum(x * x, dim=1) + 1e-10)
        return (x / norm.unsqueeze(-1)).numpy()


def hardnet(patches_u8):
    """What the desc daemon replies: float32 holding uint8(clip(210*(d+0.45),0,255)) (desc_server.py:42)."""
    d = hardnet_raw(patches_u8).astype(np.float64)
    return np.clip(210 * (d + 0.45), 0, 255).astype(np.uint8).astype(np.float32)


def affnet(patches_u8):
    with torch.no_grad():
        x, w = trunk("affnet", patches_u8)
        k = w["h_w"].permute(0, 3, 1, 2).contiguous()
        xy = torch.tanh(F.conv2d(x, k, w["h_b"])).reshape(-1, 3).clone()
        xy[:, 0] += 1
        xy[:, 2] += 1
        return xy.numpy()


def orinet(patches_u8):
    with torch.no_grad():
        x, w = trunk("orinet", patches_u8)
        k = w["h_w"].permute(0, 3, 1, 2).contiguous()
        y = torch.tanh(F.conv2d(x, k, w["h_b"], padding=1))       # 3x3 map
        return y.mean(dim=(2, 3)).reshape(-1, 2).numpy()


FORWARD = {0: affnet, 1: orinet, 2: hardnet}
