"""Sweep of k_umma_pace (modsgpu_debug_umma_pace): cycles per tcgen05.mma (M128, K16, fp16) against operand layout.
Run on the GPU box:  python tools/umma_pace.py > gpurun_out/umma_pace.txt
Each configuration runs in its own process under `timeout` so a faulting descriptor cannot poison the sweep."""
import ctypes as C
import subprocess
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(cfg):
    import mods_light_zmq_b200 as M
    g = M.ModsGpu(0)
    arr = (C.c_int * 11)(*cfg)
    out = C.c_double()
    best = 1e30
    for _ in range(3):
        rc = g.lib.modsgpu_debug_umma_pace(g.ctx, arr, C.byref(out))
        if rc != 0:
            print("ERR %d %s" % (rc, g.lib.modsgpu_last_error(g.ctx).decode()))
            return
        best = min(best, out.value)
    print("%.1f" % best)


def configs():
    REPS = 4096
    out = []
    for n in (16, 64, 256):
        bl, bs = n * 16, 128
        for grid in ((1, 148) if n == 64 else (1,)):
            # SWIZZLE_NONE, conv-like planes (K halves one plane apart), A start shifted by `off` bytes (rows are 16 B)
            for off in (0, 16, 32, 64):
                out.append(("none plane-lbo off=%d" % off, [n, off, 2608 * 1, 128, bl, bs, 0, REPS, 1, 0, grid]))
            out.append(("none plane-lbo=2560 off=0", [n, 0, 2560, 128, bl, bs, 0, REPS, 1, 0, grid]))
            out.append(("none plane-lbo=2560 off=16", [n, 16, 2560, 128, bl, bs, 0, REPS, 1, 0, grid]))
            # canonical dense packing: K halves adjacent (LBO 128), 8-row groups 256 B apart
            out.append(("none dense off=0", [n, 0, 128, 256, 128, 256, 0, REPS, 1, 0, grid]))
            # two accumulators / walking A
            out.append(("none plane-lbo off=16 2acc", [n, 16, 2608, 128, bl, bs, 0, REPS, 2, 0, grid]))
            out.append(("none plane-lbo off=16 walkA", [n, 16, 2608, 128, bl, bs, 0, REPS, 1, 5216, grid]))
            # swizzled K-major: rows 32/64/128 B apart, start shifted by whole rows
            for off in (0, 32, 96):
                out.append(("sw32 off=%d" % off, [n, off, 16, 256, 16, 256, 6, REPS, 1, 0, grid]))
            for off in (0, 64):
                out.append(("sw64 off=%d" % off, [n, off, 16, 512, 16, 512, 4, REPS, 1, 0, grid]))
            for off in (0, 128, 384):
                out.append(("sw128 off=%d" % off, [n, off, 16, 1024, 16, 1024, 2, REPS, 1, 0, grid]))
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one([int(v) for v in sys.argv[2:]])
        sys.exit(0)
    print("%-34s %4s %5s  cycles/mma" % ("layout", "N", "grid"))
    for name, cfg in configs():
        try:
            r = subprocess.run(["timeout", "60", sys.executable, __file__, "--one"] + [str(v) for v in cfg],
                               capture_output=True, text=True)
            res = (r.stdout.strip().splitlines() or ["(rc %d) %s" % (r.returncode, r.stderr.strip()[-120:])])[-1]
        except Exception as e:  # noqa
            res = "failed: %s" % e
        print("%-34s %4d %5d  %s" % (name, cfg[0], cfg[10], res), flush=True)
