#!/usr/bin/env python
"""Offline export of the reference's network weights to .npz for libmodsgpu.so.

Reads build/{AffNet,OriNet,HardNet++}.pth from the reference tree (PyTorch-0.4 pickles with keys
features.N.{weight,running_mean,running_var}[,bias]; model definitions in build/affnet_server.py:45-84,
orinet_server.py:45-82, desc_server.py:58-92) and writes weights/{affnet,orinet,hardnet}.npz with
BatchNorm (affine=False, eps 1e-5, running statistics) folded into the preceding convolution:

    c{i}_w  [Cout, 3, 3, Cin]  float32   i = 1..6   (w / sqrt(var + eps))
    c{i}_b  [Cout]             float32              (-mean / sqrt(var + eps))
    h_w     [Cout, 8, 8, Cin]  float32   the 8x8 head (HardNet: BN folded; AffNet/OriNet: plain)
    h_b     [Cout]             float32

Only weights are exported -- no code is copied.  Run in the build container (needs torch and
/root/reference); the .npz files are committed because the GPU box has no reference tree.
"""
import argparse
import os

import numpy as np
import torch

CONV_IDX = [0, 3, 6, 9, 12, 15]   # nn.Sequential indices of the six 3x3 convolutions
HEAD_IDX = {"affnet": 19, "orinet": 19, "hardnet": 19}
FILES = {"affnet": "AffNet.pth", "orinet": "OriNet.pth", "hardnet": "HardNet++.pth"}


def fold(sd, conv_i, eps=1e-5):
    w = sd["features.%d.weight" % conv_i].double()
    mean = sd["features.%d.running_mean" % (conv_i + 1)].double()
    var = sd["features.%d.running_var" % (conv_i + 1)].double()
    inv = 1.0 / torch.sqrt(var + eps)
    wf = (w * inv.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous()
    return wf.float().numpy(), (-mean * inv).float().numpy()


def export(ref_build, name, out_dir):
    ck = torch.load(os.path.join(ref_build, FILES[name]), map_location="cpu", weights_only=False)
    sd = ck["state_dict"]
    out = {}
    for li, ci in enumerate(CONV_IDX, start=1):
        out["c%d_w" % li], out["c%d_b" % li] = fold(sd, ci)
    hi = HEAD_IDX[name]
    if name == "hardnet":
        out["h_w"], out["h_b"] = fold(sd, hi)
    else:
        out["h_w"] = sd["features.%d.weight" % hi].permute(0, 2, 3, 1).contiguous().float().numpy()
        out["h_b"] = sd["features.%d.bias" % hi].float().numpy()
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".npz")
    np.savez(path, **{k: np.ascontiguousarray(v, np.float32) for k, v in out.items()})
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-build", default="/root/reference/build")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights"))
    a = ap.parse_args()
    for n in FILES:
        export(a.ref_build, n, a.out)
