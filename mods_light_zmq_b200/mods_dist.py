"""Config 5 across GPUs: one hard pair, the MODS iteration loop of mods.cpp:202-356 with the synthesised
VIEW as the unit of work.

Per iteration step (one section of an iters_*.ini schedule):
  1. SetVSPars (synth-detection.cpp:191-322): the step's view list minus the views of earlier steps;
  2. the (image, view) units of the step are dealt round-robin to the ranks; a rank synthesises, detects and describes
     its units on its own GPU (modsgpu_extract_features_views, one view per call);
  3. ONE exchange: an all-gather of the ranks' region rows (x, y, s, A, response, octave, type, view and the descriptor
     as 128 bytes = 216 bytes per region), sizes first, then the padded byte blocks;
  4. every rank appends the rows to the pair's accumulated region lists in (image, view) order -- the order the
     single-GPU loop (mods_host.cpp:MODSPair, AddViews) produces, so matching sees the same lists;
  5. the linear matcher is sharded by QUERY ROWS: rank r runs the FGINN matcher (modsgpu_match_fginn) for the contiguous
     slice [lo_r, hi_r) of image 1's accumulated regions against ALL regions of image 2 -- a query's result depends on
     nothing but its own row, so the slices concatenated in rank order are the single-GPU tentative list, bit for bit;
     ONE second exchange gathers the tentative rows (32 bytes each);
  6. rank 0 verifies them (duplicate filter, LO-RANSAC H or F, empirical checks: modsgpu_verify_matches) and broadcasts
     the verified count; the loop stops at the first step with >= min_matches.

Host logic only: the GPU work is behind the two callables, so the dealing / exchange / ordering is tested on CPU with a
world-size-2 gloo group (tests/test_sharding_gloo.py) and on the GPU against modsgpu_mods_pair (tests/test_gpu_parity.py).
"""
import numpy as np

from . import FEATURE_DTYPE, VIEW_DTYPE, view_schedule

EPS_VIEW = 0.01   # SetVSPars eps1


def step_views(step, history):
    """The views a step ADDS: its SetVSPars list without the views already in `history` (appended to in place)."""
    vs = view_schedule(step.get("scales", [1.0]), step.get("tilts", [1.0]), step.get("phi", 360.0),
                       step.get("init_sigma", 0.2), step.get("do_blur", 1))
    new = []
    for v in vs:
        seen = any(abs(v["zoom"] - q["zoom"]) <= EPS_VIEW and abs(v["tilt"] - q["tilt"]) <= EPS_VIEW and
                   abs(v["phi"] - q["phi"]) <= EPS_VIEW for q in history)
        if not seen:
            new.append(v.copy())
    history.extend(new)     # like SetVSPars: a step's own views are not compared with each other
    out = np.zeros(len(new), VIEW_DTYPE)
    for i, v in enumerate(new):
        out[i] = v
    return out


def deal_units(n_views, rank, world):
    """(image, view) units of a step in canonical order and the ones this rank owns (round-robin)."""
    units = [(k, j) for k in (0, 1) for j in range(n_views)]
    return units, [u for i, u in enumerate(units) if i % world == rank]


WIRE_DTYPE = np.dtype([(n, FEATURE_DTYPE.fields[n][0]) for n in FEATURE_DTYPE.names if n != "desc"] + [("desc", "u1", (128,))])


def to_wire(rows):
    """FEATURE rows -> the exchange format: the same fields with the descriptor as 128 bytes (HardNet++ / RootSIFT entries
    are integers 0..255: desc_server.py:42, siftdesc.cpp) -- 216 instead of 600 bytes per region"""
    w = np.zeros(len(rows), WIRE_DTYPE)
    for n in WIRE_DTYPE.names:
        w[n] = rows[n]
    return w


def from_wire(w):
    rows = np.zeros(len(w), FEATURE_DTYPE)
    for n in WIRE_DTYPE.names:
        rows[n] = w[n]
    return rows


def all_gather_rows(mine, n_units, dist=None, device=None):
    """mine: {unit index: FEATURE rows} of this rank.  Returns the rows of every unit, on every rank.  One collective for
    the sizes, one for the payload (padded to the largest rank's byte count, descriptors as bytes)."""
    if dist is None:
        return [mine[i] for i in range(n_units)]
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(n_units, dtype=torch.int64)
    for i, rows in mine.items():
        counts[i] = len(rows)
    counts = counts.to(device) if device is not None else counts
    dist.all_reduce(counts)                              # every unit has exactly one owner
    counts = counts.cpu().numpy()
    owner = [i % world for i in range(n_units)]
    per_rank = [int(sum(counts[i] for i in range(n_units) if owner[i] == r)) for r in range(world)]
    item = WIRE_DTYPE.itemsize
    cap = max(max(per_rank), 1) * item
    own = [mine[i] for i in range(n_units) if owner[i] == rank]
    blob = to_wire(np.concatenate(own)).view(np.uint8).reshape(-1) if own and per_rank[rank] else np.zeros(0, np.uint8)
    send = torch.zeros(cap, dtype=torch.uint8)
    send[:blob.size] = torch.from_numpy(np.ascontiguousarray(blob))
    if device is not None:
        send = send.to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    out, cursor = [None] * n_units, [0] * world
    host = [t.cpu().numpy() for t in recv]
    for i in range(n_units):
        r, nb = owner[i], int(counts[i]) * item
        out[i] = from_wire(host[r][cursor[r]:cursor[r] + nb].view(WIRE_DTYPE))
        cursor[r] += nb
    return out


def query_slice(n_queries, rank, world):
    """contiguous, balanced slice of the query rows for one rank"""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_matches(mine, dist=None, device=None):
    """mine: MATCH rows of this rank's query slice (qi already global).  Returns the rows of all ranks in rank order."""
    from . import MATCH_DTYPE
    mine = np.ascontiguousarray(mine, MATCH_DTYPE)
    if dist is None:
        return mine
    import torch
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64)
    counts[dist.get_rank()] = len(mine)
    counts = counts.to(device) if device is not None else counts
    dist.all_reduce(counts)
    counts = counts.cpu().numpy()
    cap = max(int(counts.max()), 1) * MATCH_DTYPE.itemsize
    send = torch.zeros(cap, dtype=torch.uint8)
    blob = mine.view(np.uint8).reshape(-1)
    send[:blob.size] = torch.from_numpy(np.ascontiguousarray(blob))
    if device is not None:
        send = send.to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    parts = [t.cpu().numpy()[:int(counts[r]) * MATCH_DTYPE.itemsize].view(MATCH_DTYPE) for r, t in enumerate(recv)]
    return np.concatenate(parts) if parts else np.zeros(0, MATCH_DTYPE)


def mods_pair_sharded(extract_view, match, steps, rank=0, world=1, dist=None, device=None, min_matches=10,
                      match_slice=None, verify=None):
    """extract_view(image_index, view_row) -> FEATURE rows of that view (reprojected to the original image);
    match(f1, f2, fginn) -> dict with at least `inliers` (rank 0 only): matcher + verification on one GPU; or, with
    match_slice(f1, f2, lo, hi, fginn) -> MATCH rows of queries [lo, hi) (qi global) and verify(f1, f2, matches) -> dict,
    the matcher runs on every rank's slice of the query rows and rank 0 only verifies.
    Returns dict(steps_done, views, regions, features=[f1, f2], result=<match dict on rank 0, None elsewhere>, inliers)."""
    import time
    history = []
    feats = [np.zeros(0, FEATURE_DTYPE), np.zeros(0, FEATURE_DTYPE)]
    n_views_total, inliers, result, steps_done = 0, 0, None, 0
    phase = dict(extract=0.0, gather=0.0, match=0.0, verify=0.0)      # wall seconds of this rank per phase
    for step in steps:
        if inliers >= min_matches:
            break
        t0 = time.perf_counter()
        views = step_views(step, history)
        units, my_units = deal_units(len(views), rank, world)
        mine = {}
        for (k, j) in my_units:
            rows = np.ascontiguousarray(extract_view(k, views[j]), FEATURE_DTYPE).copy()
            rows["view"] = n_views_total + j
            mine[units.index((k, j))] = rows
        t1 = time.perf_counter()
        rows_of = all_gather_rows(mine, len(units), dist if world > 1 else None, device)
        for k in (0, 1):                                 # (image, view) order = AddViews order
            feats[k] = np.concatenate([feats[k]] + [rows_of[i] for i, (kk, j) in enumerate(units) if kk == k])
        t2 = time.perf_counter()
        n_views_total += len(views)
        steps_done += 1
        t3 = t2
        if match_slice is not None and verify is not None:
            lo, hi = query_slice(len(feats[0]), rank, world)
            rows = match_slice(feats[0], feats[1], lo, hi, step.get("fginn", 0.8))
            rows = all_gather_matches(rows, dist if world > 1 else None, device)
            t3 = time.perf_counter()
            if rank == 0:
                result = verify(feats[0], feats[1], rows)
                inliers = int(result["inliers"])
        elif rank == 0:
            result = match(feats[0], feats[1], step.get("fginn", 0.8))
            inliers = int(result["inliers"])
        t4 = time.perf_counter()
        phase["extract"] += t1 - t0; phase["gather"] += t2 - t1; phase["match"] += t3 - t2; phase["verify"] += t4 - t3
        if dist is not None and world > 1:
            import torch
            t = torch.tensor([inliers], dtype=torch.int64)
            t = t.to(device) if device is not None else t
            dist.broadcast(t, src=0)
            inliers = int(t.item())
    return dict(steps_done=steps_done, views=n_views_total, regions=[len(feats[0]), len(feats[1])], features=feats,
                result=result, inliers=inliers, phase_ms={k: round(v * 1e3, 1) for k, v in phase.items()})


def gpu_callables(mg, img1, img2, use_F=False, seed=12345, capacity=8192):
    """The two callables of mods_pair_sharded on one modsgpu context."""
    imgs = (img1, img2)

    def extract_view(k, view):
        v = np.zeros(1, VIEW_DTYPE)
        v[0] = view
        return mg.extract_features_views(imgs[k], v)

    def match(f1, f2, fginn):
        return mg.match_features(f1, f2, fginn=fginn, use_F=use_F, seed=seed, capacity=capacity)

    return extract_view, match


def gpu_sharded_match_callables(mg, use_F=False, seed=12345, capacity=8192):
    """match_slice / verify of mods_pair_sharded on one modsgpu context (matcher sharded by query rows)"""
    def match_slice(f1, f2, lo, hi, fginn):
        m = mg.match_fginn(f1["desc"][lo:hi], f2["desc"], np.c_[f2["x"], f2["y"]], ratio=fginn)
        m["qi"] += lo
        return m

    def verify(f1, f2, rows):
        return mg.verify_matches(f1, f2, rows, use_F=use_F, seed=seed, capacity=capacity)

    return match_slice, verify


def main(argv=None):
    """torchrun entry: python -m mods_light_zmq_b200.mods_dist img1 img2 [--F] -- the iters_MODS_ZMQ.ini HessianAffine
    steps on the ranks' GPUs; rank 0 prints the result."""
    import argparse, json, os
    import mods_light_zmq_b200 as M
    from .batch import read_image_bgr
    ap = argparse.ArgumentParser()
    ap.add_argument("img1")
    ap.add_argument("img2")
    ap.add_argument("--F", action="store_true", help="verify with LO-RANSAC(F) (LORANSACF, mods.cpp:325)")
    ap.add_argument("--min-matches", type=int, default=15)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--time", action="store_true", help="run twice and report the wall time of the second run (max over ranks)")
    ap.add_argument("--rank0-matcher", action="store_true", help="match on rank 0 only (the round-1 layout) instead of sharding the query rows")
    a = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)
    mg = M.ModsGpu(local, load_nets=True)
    i1, i2 = mg.image_from_bgr8(read_image_bgr(a.img1)), mg.image_from_bgr8(read_image_bgr(a.img2))
    ev, mt = gpu_callables(mg, i1, i2, use_F=a.F, seed=a.seed)
    ms_, vf_ = (None, None) if a.rank0_matcher else gpu_sharded_match_callables(mg, use_F=a.F, seed=a.seed)
    ms = None
    if a.time:   # one untimed run, then one between barriers; the slowest rank's wall clock
        import time
        mods_pair_sharded(ev, mt, MODS_ZMQ_HESSIAN_STEPS, rank, world, dist, device, a.min_matches, ms_, vf_)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
    r = mods_pair_sharded(ev, mt, MODS_ZMQ_HESSIAN_STEPS, rank, world, dist, device, a.min_matches, ms_, vf_)
    if a.time:
        ms = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            import torch
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    if rank == 0:
        res = r["result"] or {}
        print(json.dumps(dict(world=world, matcher="rank 0" if a.rank0_matcher else "sharded by query rows", ms=ms,
                              rank0_phase_ms=r["phase_ms"], steps_done=r["steps_done"], views=r["views"], regions=r["regions"],
                              tentatives=res.get("tentatives"), inliers=r["inliers"],
                              model=[float(v) for v in res.get("model", [])])))
    if dist is not None:
        dist.destroy_process_group()
    mg.close()


# the HessianAffine steps of build/iters_MODS_ZMQ.ini:30-50 ([HessianAffine2]: TiltSet 1,2,4,6,8, ScaleSet 1, Phi 360;
# [HessianAffine3]: the same sets with Phi 120); steps 0 and 1 of that file are MSER-only and stay with the reference's
# CPU path
MODS_ZMQ_HESSIAN_STEPS = [
    dict(scales=[1.0], tilts=[1.0, 2.0, 4.0, 6.0, 8.0], phi=360.0, init_sigma=0.2, fginn=0.8),
    dict(scales=[1.0], tilts=[1.0, 2.0, 4.0, 6.0, 8.0], phi=120.0, init_sigma=0.2, fginn=0.8),
]

if __name__ == "__main__":
    main()
