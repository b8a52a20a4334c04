"""bench.py contract tests.  The GPU arm is exercised with the SMALL step counts the driver uses (the round-1 bench
died for any --steps below 64); the reference arm's line shape is checked on the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                       timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.gpu
@pytest.mark.parametrize("steps,warmup", [(1, 1), (3, 1)])
def test_bench_small_step_counts(steps, warmup):
    r = _run(["--gpus", "1", "--steps", str(steps), "--warmup", str(warmup), "--no-cpu-baseline"], 900)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "stages", "kernels"):
        assert k in r, k
    assert r["steps"] == steps and r["value"] > 0 and r["gpu_launches"] > 0
    assert r["e2e"]["value"] > 0 and r["e2e"]["h2d_bytes_per_step"] > 0 and r["e2e"]["d2h_bytes_per_step"] > 0
    rf = r["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    # pairs/s of the line = pairs of the timed arm / its device time
    pairs = steps * r["config"]["pairs_per_step"]
    assert abs(r["value"] - pairs / (r["ms_per_step"] * steps * 1e-3)) < 1e-6 * r["value"]


def test_reference_arm_line_shape():
    """One whole pair on this machine's cores (no GPU): the line carries impl / cpu_baseline / e2e and its
    ms_per_step is the measured time of the step, not an extrapolation."""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert r["impl"] == "reference" and r["unit"] == "pairs/s"
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] == os.cpu_count()
    assert r["e2e"] == {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(r["value"] - 1e3 / r["ms_per_step"]) < 1e-9
    c = r["config"]["last_step"]
    assert c["tentatives"] >= c["unique_tentatives"] >= c["inliers"] > 100      # RANSAC ran on the pair's own tentatives
    sec = r["cpu_baseline"]["stage_seconds_last_pair"]
    assert sum(sec.values()) <= r["ms_per_step"] * 1e-3 * 1.05
