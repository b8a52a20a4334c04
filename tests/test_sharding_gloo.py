"""Host-side logic of the multi-GPU path on CPU: world-size-2 `gloo` runs of the batch extractor (config 4) with a
deterministic stand-in extractor (no GPU here), compared file-by-file with the world-size-1 output; plus the
pair-sharding arithmetic bench.py uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import os, sys, zlib
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import batch

def fake_extractor(bgr):
    # deterministic function of the image content only (stands in for the GPU pipeline)
    seed = zlib.crc32(bgr.tobytes()) & 0x7fffffff
    rng = np.random.RandomState(seed)
    n = 3 + seed %% 5
    f = np.zeros(n, M.FEATURE_DTYPE)
    f["x"], f["y"] = rng.uniform(1, 60, n), rng.uniform(1, 40, n)
    f["s"] = rng.uniform(2, 9, n)
    f["a11"], f["a22"] = 1.0, 1.0
    f["desc"] = rng.randint(0, 256, (n, 128))
    return f

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
d = None
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
    d = dist
imgs = open(sys.argv[1]).read().split()
outs = open(sys.argv[2]).read().split()
counts = batch.extract_features_batch(imgs, outs, fake_extractor, rank, world, d)
if rank == 0:
    print("COUNTS", " ".join(str(c) for c in counts))
if d is not None:
    dist.destroy_process_group()
'''


def _run(tmp, tag, world, imgs, port):
    outdir = tmp / tag
    outdir.mkdir()
    outs = [str(outdir / ("img%02d.oxaff" % i)) for i in range(len(imgs))]
    (tmp / (tag + "_imgs.txt")).write_text("\n".join(imgs) + "\n")
    (tmp / (tag + "_outs.txt")).write_text("\n".join(outs) + "\n")
    script = tmp / (tag + "_worker.py")
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(tmp / (tag + "_imgs.txt")), str(tmp / (tag + "_outs.txt"))],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, res):
        assert p.returncode == 0, e[-2000:]
    counts = [int(v) for v in res[0][0].split("COUNTS")[1].split()]
    return outs, counts


def test_shard_indices_cover_every_image_once():
    from mods_light_zmq_b200 import batch
    for n in (0, 1, 7, 512):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in batch.shard_indices(n, r, world))
            assert seen == list(range(n))
            sizes = [len(batch.shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_batch_extract_world2_equals_world1(tmp_path):
    rng = np.random.RandomState(0)
    imgs = []
    for i in range(7):
        a = rng.randint(0, 256, (40, 64, 3)).astype(np.uint8)
        if i % 2:
            p = tmp_path / ("in%02d.npy" % i)
            np.save(p, a)
        else:   # binary PPM (RGB on disk, BGR in memory like cv::imread)
            p = tmp_path / ("in%02d.ppm" % i)
            with open(p, "wb") as f:
                f.write(b"P6\n# synthetic\n64 40\n255\n" + a[:, :, ::-1].tobytes())
        imgs.append(str(p))
    imgs.append(str(tmp_path / "missing.ppm"))          # unreadable image: reported, not fatal
    outs1, c1 = _run(tmp_path, "w1", 1, imgs, 29611)
    outs2, c2 = _run(tmp_path, "w2", 2, imgs, 29612)
    assert c1 == c2 and c1[-1] == -1 and all(c >= 3 for c in c1[:-1])
    for a, b, c in zip(outs1[:-1], outs2[:-1], c1[:-1]):
        ta, tb = open(a).read(), open(b).read()
        assert ta == tb                                  # file-by-file identical to the 1-rank output
        assert ta.split("\n")[0] == "128" and int(ta.split("\n")[1]) == c
    assert not os.path.exists(outs2[-1])
    # a second run skips everything that exists (extract_features_batch.cpp:108-117)
    (tmp_path / "w2").rename(tmp_path / "w2_first")
    (tmp_path / "w2_first").rename(tmp_path / "w2")
    script = tmp_path / "w2_worker.py"
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, str(script), str(tmp_path / "w2_imgs.txt"), str(tmp_path / "w2_outs.txt")],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert [int(v) for v in out.stdout.split("COUNTS")[1].split()] == [-1] * len(imgs)


def test_pnm_reader_matches_memory(tmp_path):
    from mods_light_zmq_b200 import batch
    rng = np.random.RandomState(1)
    g = rng.randint(0, 256, (9, 13)).astype(np.uint8)
    p = tmp_path / "g.pgm"
    with open(p, "wb") as f:
        f.write(b"P5 13 9 255\n" + g.tobytes())
    a = batch.read_image_bgr(str(p))
    assert a.shape == (9, 13, 3) and np.array_equal(a[:, :, 0], g) and np.array_equal(a[:, :, 2], g)


# ------------------------------------------------------------------------------------------ config 5: views across ranks
MODS_WORKER = r'''
import os, sys, zlib
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import mods_dist as D

def fake_extract_view(k, view):
    # deterministic function of (image, view) only; some views give no region at all
    seed = zlib.crc32(np.array([k, view["tilt"], view["phi"], view["zoom"]], np.float64).tobytes()) & 0x7fffffff
    rng = np.random.RandomState(seed)
    n = seed %% 6
    f = np.zeros(n, M.FEATURE_DTYPE)
    f["x"], f["y"], f["s"] = rng.uniform(1, 60, n), rng.uniform(1, 40, n), rng.uniform(2, 9, n)
    f["a11"], f["a22"] = 1.0, 1.0
    f["desc"] = rng.randint(0, 256, (n, 128))
    return f

calls = []
def fake_match(f1, f2, fginn):
    calls.append((len(f1), len(f2)))
    return dict(inliers=(len(f1) + len(f2)) // int(%(div)d), tentatives=len(f1))

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
d = None
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
    d = dist
steps = [dict(tilts=[1.0], phi=360.0), dict(tilts=[1.0, 2.0, 4.0], phi=360.0), dict(tilts=[1.0, 2.0, 4.0, 6.0], phi=120.0)]
r = D.mods_pair_sharded(fake_extract_view, fake_match, steps, rank, world, d, None, min_matches=%(minm)d)
crc = [zlib.crc32(f.tobytes()) for f in r["features"]]
print("RESULT", rank, r["steps_done"], r["views"], r["regions"][0], r["regions"][1], crc[0], crc[1], r["inliers"], len(calls),
      list(r["features"][0]["view"]) == sorted(r["features"][0]["view"]))
if d is not None:
    dist.destroy_process_group()
'''


def _run_mods(tmp, tag, world, port, minm, div=1000):
    script = tmp / (tag + "_mods_worker.py")
    script.write_text(MODS_WORKER % {"root": ROOT, "port": port, "minm": minm, "div": div})
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = [p.communicate(timeout=240) for p in procs]
    out = []
    for p, (o, e) in zip(procs, res):
        assert p.returncode == 0, e[-2000:]
        out.append(o.split("RESULT")[1].split())
    return out


def test_mods_views_world2_equals_world1(tmp_path):
    """The view-sharded MODS loop (mods_dist.py): with two ranks every rank ends with the region lists, in the order, of
    the one-rank run; only rank 0 matches; all ranks leave the loop at the same step."""
    one = _run_mods(tmp_path, "m1", 1, 29621, minm=10 ** 9)[0]
    two = _run_mods(tmp_path, "m2", 2, 29622, minm=10 ** 9)
    # steps_done, views, regions, checksums of both lists
    assert one[1] == "3" and int(one[2]) == 1 + 3 + 15    # 1; tilt 2: 1 rotation, tilt 4: 2; Phi 120: 3 + 6 + 9 - the 3 already seen
    for r in two:
        assert r[1:8] == one[1:8], (r, one)
        assert r[-1] == "True"                            # rows of one image stay in view order
    assert two[0][8] == "3" and two[1][8] == "0"          # matching ran on rank 0 only
    # early stop is decided on rank 0 and followed by every rank
    one_s = _run_mods(tmp_path, "m3", 1, 29623, minm=1, div=1)[0]
    two_s = _run_mods(tmp_path, "m4", 2, 29624, minm=1, div=1)
    assert one_s[1] == two_s[0][1] == two_s[1][1] and int(one_s[1]) < 3
    assert two_s[0][7] == two_s[1][7] == one_s[7]        # the broadcast verified count


SHARD_MATCH_WORKER = r'''
import os, sys, zlib
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import mods_light_zmq_b200 as M
from mods_light_zmq_b200 import mods_dist as D
from oracle import pyoracle as O          # the CPU matcher stands in for modsgpu_match_fginn (tests only)

def fake_extract_view(k, view):
    seed = zlib.crc32(np.array([k, view["tilt"], view["phi"], view["zoom"]], np.float64).tobytes()) & 0x7fffffff
    rng = np.random.RandomState(seed)
    n = 40 + seed %% 30
    f = np.zeros(n, M.FEATURE_DTYPE)
    base = np.random.RandomState(7).randint(0, 256, (64, 128))          # shared by both images: true matches exist
    pick = rng.randint(0, 64, n)
    f["x"], f["y"], f["s"] = pick * 10.0 + rng.uniform(0, 2, n), pick * 5.0 + rng.uniform(0, 2, n), 3.0
    f["a11"], f["a22"] = 1.0, 1.0
    f["desc"] = np.clip(base[pick] + rng.randint(-5, 6, (n, 128)), 0, 255)
    return f

def match_slice(f1, f2, lo, hi, fginn):
    m = O.match_fginn(f1["desc"][lo:hi], np.zeros((hi - lo, 2)), f2["desc"], np.c_[f2["x"], f2["y"]], ratio=fginn)
    m = m.astype(M.MATCH_DTYPE)
    m["qi"] += lo
    return m

seen = {}
def verify(f1, f2, rows):
    seen["rows"] = rows
    return dict(inliers=len(rows), tentatives=len(rows))

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
d = None
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
    d = dist
steps = [dict(tilts=[1.0, 2.0], phi=360.0)]
r = D.mods_pair_sharded(fake_extract_view, None, steps, rank, world, d, None, min_matches=10 ** 9, match_slice=match_slice, verify=verify)
rows = seen.get("rows")
print("RESULT", rank, r["regions"][0], r["regions"][1], r["inliers"], -1 if rows is None else zlib.crc32(rows.tobytes()),
      -1 if rows is None else int((np.diff(rows["qi"]) > 0).all()))
if d is not None:
    dist.destroy_process_group()
'''


def test_sharded_matcher_world2_equals_world1(tmp_path):
    """mods_dist.py with the matcher sharded by query rows: the tentative rows rank 0 verifies with two ranks are the
    one-rank list byte for byte (a query's FGINN result depends on its own row only; slices are concatenated in rank
    order), and every rank learns the verified count."""
    def run(tag, world, port):
        script = tmp_path / (tag + "_worker.py")
        script.write_text(SHARD_MATCH_WORKER % {"root": ROOT, "port": port})
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        out = []
        for p in procs:
            o, e = p.communicate(timeout=240)
            assert p.returncode == 0, e[-2000:]
            out.append(o.split("RESULT")[1].split())
        return out
    one = run("s1", 1, 29631)[0]
    two = run("s2", 2, 29632)
    assert int(one[3]) > 20 and one[5] == "1"                        # tentatives exist, in query order
    assert two[0][1:5] == one[1:5] and two[0][5] == "1"              # same lists, same tentative rows on rank 0
    assert two[1][1:4] == one[1:4] and two[1][4] == "-1"             # rank 1 verified nothing but knows the count
    from mods_light_zmq_b200 import mods_dist as D
    assert [D.query_slice(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert D.query_slice(0, 1, 2) == (0, 0)


def test_step_views_follow_SetVSPars_history():
    from mods_light_zmq_b200 import mods_dist as D
    hist = []
    n = [len(D.step_views(st, hist)) for st in D.MODS_ZMQ_HESSIAN_STEPS]
    # [HessianAffine2] Phi 360: 1 + (1 + 2 + 3 + 4) rotations; [HessianAffine3] Phi 120: 3 + 6 + 9 + 12 minus the phi = 0 views
    assert n == [11, 20], n
    assert len(hist) == sum(n)
    units, mine = D.deal_units(5, 1, 3)
    assert len(units) == 10 and mine == [units[1], units[4], units[7]]
