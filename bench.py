#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the MODS hot path (detect -> AffNet -> OriNet -> HardNet++ -> FGINN match ->
duplicate filter -> LO-RANSAC(H)) on synthetic 1024x768 pairs (~4k keypoints per image).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one BATCH of PAIRS_PER_STEP (16) independent image pairs through modsgpu_pair_pipeline*; `value`
counts pairs (steps x 16 x ranks / time).  Pairs are independent, so ranks shard pairs (weak scaling: every
rank runs K steps on its own pairs); the only collective is one NCCL gather of the final correspondences /
homographies of all pairs, inside the timed region.

value   whole-job pairs/s with the images already resident in HBM (modsgpu_pair_pipeline_images)
e2e     the same through the reference-facing call with HOST (pinned) BGR images: H2D of both images and D2H
        of the verified correspondences inside the timed region (modsgpu_pair_pipeline)
roofline  dominant kernel by accumulated CUDA-event time over the same K steps (per-launch events recorded
        by the library on its own stream), algorithmic work per launch as defined in DESIGN.md
cpu_baseline / --impl reference  the CPU path (oracle port of the reference C++ stages, the daemons' torch
        models on the CPU, the reference's own degensac on the pair's own tentatives) on the host cores; every
        reference step times ONE WHOLE pair of the step's batch (the bounded sample), nothing is extrapolated
The L2 is flushed (256 MB memset) before every step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# 16 contexts (32 streams) per GPU: one hardware work queue per stream (the default of 8 costs 24 % more host CPU per pair);
# must be in the environment before torch creates the CUDA context
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "image-pairs/sec (1024x768, ~4k kp)"
WORKLOAD = "config3: 1024x768 synthetic pair, Hessian-AffNet-OriNet-HardNet++ + linear FGINN + LO-RANSAC(H)"
W_IMG, H_IMG = 1024, 768
N_DISTINCT_PAIRS = 4
CAPACITY = 2048
PAIRS_PER_STEP = 16


def make_pairs(n, rank):
    from mods_light_zmq_b200 import synth
    pairs = []
    for i in range(n):
        a, b, H = synth.image_pair(seed=1234 + 17 * i + 1000 * rank, w=W_IMG, h=H_IMG)
        pairs.append((synth.gray_to_bgr(a), synth.gray_to_bgr(b), H))
    return pairs


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU path
def cpu_pair(pair, threads):
    """The reference's CPU path on ONE WHOLE pair (every keypoint, no sub-sampling): detection + sampler + AffNet +
    OriNet + HardNet++ per image (the two images concurrently, as mods.cpp:234-251 does with OpenMP tasks), linear
    FGINN, duplicate filter, and LO-RANSAC(H) by the reference's own degensac (oracle/_ref, exp_ransacHcustom as
    matching.cpp:731 calls it) on the pair's OWN tentatives.  Returns per-stage seconds, the pair's wall seconds
    and the counts."""
    import torch
    from oracle import pyoracle as O
    from oracle import cnn_oracle as CN
    torch.set_num_threads(threads)
    t = {"detect": 0.0, "perkp": 0.0, "match": 0.0, "dup": 0.0, "ransac": 0.0}
    out = [None, None]
    t_begin = time.perf_counter()

    def one(i):
        bgr = pair[i]
        t0 = time.perf_counter()
        g = O.gray_from_bgr(bgr)
        h, w = g.shape
        kp = O.detect_hessian(g)
        t1 = time.perf_counter()
        regs = O.regions_from_keypoints(kp)
        aff = CN.affnet(O.quantize_u8(O.extract_patches(g, regs)))
        r2, _ = O.affnet_postprocess(regs, aff, w, h)
        ori = CN.orinet(O.quantize_u8(O.extract_patches(g, r2)))
        r3 = O.orinet_postprocess(r2, ori)
        r4, _ = O.reproject_filter(r3, w, h)
        d = CN.hardnet(O.quantize_u8(O.extract_patches(g, r4)))
        t2 = time.perf_counter()
        out[i] = (d, np.c_[r4["x"], r4["y"]], t1 - t0, t2 - t1, len(kp))

    ths = [threading.Thread(target=one, args=(i,)) for i in (0, 1)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    tot = sum(o[2] + o[3] for o in out)
    t["detect"] = wall * sum(o[2] for o in out) / tot
    t["perkp"] = wall * sum(o[3] for o in out) / tot
    t0 = time.perf_counter()
    m = O.match_fginn(out[0][0], out[0][1], out[1][0], out[1][1])
    t["match"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    xy1, xy2 = out[0][1][m["qi"]], out[1][1][m["ti"]]
    keep = O.duplicate_filter(xy1, xy2, m["ratio"], 2.0) if len(m) else np.zeros(0, np.int32)
    t["dup"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_inl = 0
    if len(keep) >= 8:
        u = np.ascontiguousarray(np.c_[xy1[keep], np.ones(len(keep)), xy2[keep], np.ones(len(keep))])
        if not O.ref_available():
            raise RuntimeError("oracle/_ref/libdegensac_ref.so missing: the CPU arm runs the reference's own degensac")
        n_inl = int(O.ref_ransac_H(u, th=16.0)["inl"].sum())      # err_threshold 4 px squared, matching.cpp:731
    t["ransac"] = time.perf_counter() - t0
    counts = {"keypoints": [out[0][4], out[1][4]], "descriptors": [len(out[0][0]), len(out[1][0])],
              "tentatives": int(len(m)), "unique_tentatives": int(len(keep)), "inliers": n_inl}
    return t, time.perf_counter() - t_begin, counts


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path with all host threads; under torchrun rank 0 alone works.
    Each step times one WHOLE pair (pair k mod 4 of the bench workload) -- the bounded sample of the GPU arm's 16-pair
    step; value = pairs / measured seconds, no extrapolation."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pairs = make_pairs(N_DISTINCT_PAIRS, 0)
    for k in range(max(min(args.warmup, 2), 1)):
        cpu_pair(pairs[k % len(pairs)], threads)
    secs, parts, counts = [], None, None
    for k in range(args.steps):
        parts, sec, counts = cpu_pair(pairs[k % len(pairs)], threads)
        secs.append(sec)
    per_pair = float(np.mean(secs))
    value = 1.0 / per_pair
    sample = ("each step = ONE whole pair of the 16-pair batch the GPU arm calls a step (full Hessian detection of both "
              "1024x768 images, every keypoint through sampler + AffNet + OriNet + HardNet++, linear FGINN, duplicate "
              "filter, the reference's own exp_ransacHcustom on the pair's own tentatives); %d pairs, %.1f s of CPU work, "
              "pairs/s = pairs / measured seconds" % (len(secs), float(np.sum(secs))))
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_pair * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "pairs_per_step": 1, "distinct_pairs": N_DISTINCT_PAIRS, "l2": "n/a (CPU)",
                       "last_step": counts},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": sample, "stage_seconds_last_pair": parts,
                             "note": "kind=port: detector / sampler / matcher are the oracle's C++ restatement, the nets "
                                     "are the daemons' torch models on the CPU; the LO-RANSAC stage is the reference's own "
                                     "degensac (oracle/_ref)"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU path
ALG = {  # kernels whose algorithmic work the library reports in bytes (kind 0) or flops (kind 1)
    0: ("hbm", "GB/s", 1e9), 1: ("tensor", "TFLOP/s", 1e12)}


def pin_to_gpu_numa_node(local_rank):
    """Keep this rank's threads (and the pinned buffers they first touch) on the NUMA node its GPU hangs off: on the
    2-socket 8-GPU box unpinned ranks pay cross-socket hops on every driver call.  Silently does nothing when the node is
    unknown (single-socket boxes report -1)."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out          # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        # never shrink the rank to a sliver of the machine (a cgroup may expose only a few CPUs of that node)
        if len(cpus) >= 4 and 4 * len(cpus) >= len(allowed):
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return None


def run_gpu(args, rank, world, local_rank):
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 and not os.environ.get("MODSGPU_NO_NUMA_PIN") else None
    import torch
    import mods_light_zmq_b200 as M
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nwk = args.workers
    if nwk <= 0:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        # host waits sleep (blocking-sync events in libmodsgpu), so the number of pairs in flight need not follow the
        # core count: 16 contexts per GPU also on the 8-GPU box with its 4 cores per rank
        nwk = 16
    # one modsgpu_ctx (= one CUDA stream + workspaces) per worker thread, as the C ABI prescribes; pairs are
    # independent, so workers overlap one pair's host-side seams with another pair's kernels
    mgs = [M.ModsGpu(local_rank, load_nets=True) for _ in range(nwk)]
    pairs = make_pairs(N_DISTINCT_PAIRS, rank)
    host = []
    for a, b, _ in pairs:
        pa = torch.empty(a.shape, dtype=torch.uint8).pin_memory()
        pb = torch.empty(b.shape, dtype=torch.uint8).pin_memory()
        pa.numpy()[:] = a
        pb.numpy()[:] = b
        host.append((pa, pb))
    dev = [[(mg.image_from_bgr8(pa.numpy()), mg.image_from_bgr8(pb.numpy())) for pa, pb in host] for mg in mgs]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_pairs = max(args.steps, 1) * PAIRS_PER_STEP      # pairs of the timed arms (steps x 16 per rank)
    results = np.zeros((n_pairs, 16 + 4 * CAPACITY), np.float64)
    last = {}
    host_cpu = {"ms": 0.0}
    rank_ms = {"last": []}

    # k = pair index inside the arm (step = k // PAIRS_PER_STEP); store=False for warm-up / rehearsal passes, whose
    # pair count is independent of --steps
    def step_value(wk, k, store=True):
        mg = mgs[wk]
        mg.lib.modsgpu_flush_l2(mg.ctx)
        i1, i2 = dev[wk][k % len(host)]
        r = mg.pair_pipeline_images(i1, i2, seed=1000 + k, capacity=CAPACITY)
        if store:
            results[k, :9] = r["H"].ravel()
            results[k, 9] = r["inliers"]
            results[k, 16:16 + 4 * len(r["inlier_xy"])] = r["inlier_xy"].ravel()
            last["r"] = r

    def step_e2e(wk, k, store=True):
        mg = mgs[wk]
        mg.lib.modsgpu_flush_l2(mg.ctx)
        pa, pb = host[k % len(host)]
        r = mg.pair_pipeline(pa.numpy(), pb.numpy(), seed=1000 + k, capacity=CAPACITY)
        if store:
            results[k, :9] = r["H"].ravel()
            results[k, 9] = r["inliers"]
            results[k, 16:16 + 4 * len(r["inlier_xy"])] = r["inlier_xy"].ravel()

    def gather_results():
        if dist is None:
            return
        t = torch.from_numpy(results).cuda()
        lst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, lst, dst=0)
        torch.cuda.synchronize()

    def timed(fn, steps, store=True):
        """`steps` PAIRS dealt round-robin to the worker threads; device time = CUDA events on worker 0's stream
        bracketing the whole region (start recorded after all workers are ready, stop after all have joined
        and the gather is done)."""
        errs = []
        gate = threading.Barrier(nwk + 1)

        def work(wk):
            try:
                gate.wait()
                for k in range(wk, steps, nwk):
                    fn(wk, k, store)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ths = [threading.Thread(target=work, args=(wk,)) for wk in range(nwk)]
        for t in ths:
            t.start()
        barrier()
        ms = C.c_float()
        mgs[0].lib.modsgpu_timer_start(mgs[0].ctx)
        t0 = time.perf_counter()
        c0 = time.process_time()
        gate.wait()
        for t in ths:
            t.join()
        host_cpu["ms"] = (time.process_time() - c0) * 1e3
        torch.cuda.synchronize()
        gather_results()
        mgs[0].lib.modsgpu_timer_stop(mgs[0].ctx, C.byref(ms))
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if errs:
            raise errs[0]
        dev_ms = float(ms.value)
        if os.environ.get("MODSGPU_BENCH_DEBUG"):
            print("[rank %d] %s: dev %.1f ms wall %.1f ms for %d steps" % (rank, fn.__name__, dev_ms, wall_ms, steps), file=sys.stderr, flush=True)
        rank_ms["last"] = [dev_ms]
        if dist is not None:
            t = torch.tensor([dev_ms, wall_ms], device="cuda")
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            rank_ms["last"] = [float(x[0]) for x in allt]            # per-rank device time of this arm (max = the reported one)
            dev_ms, wall_ms = max(float(x[0]) for x in allt), max(float(x[1]) for x in allt)
        return dev_ms, wall_ms

    # ---- warm-up (every worker), then the timed arms
    W = max(args.warmup, 3)
    # every context sees every distinct pair three times: the detector's launch sequence becomes a CUDA graph on the
    # second / third request per image buffer, and that one-time capture must not land in the timed region
    for wk in range(nwk):
        for k in range(max(W, 3 * len(host))):
            step_value(wk, k, False)
            step_e2e(wk, k, False)
    gather_results()      # NCCL communicators are created lazily: build them outside the timed region
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # rehearsal of both arms through the same threaded / collective machinery (first use of the NCCL barrier and
    # all-reduce costs hundreds of ms once; it must not land in the first timed arm)
    timed(step_e2e, 2 * nwk, store=False)
    timed(step_value, 4 * nwk, store=False)
    launches0 = sum(mg.launch_count for mg in mgs)
    dev_ms, wall_ms = timed(step_value, n_pairs)
    value_rank_ms = list(rank_ms["last"])
    host_cpu_ms_per_pair = host_cpu["ms"] / n_pairs
    launches = sum(mg.launch_count for mg in mgs) - launches0
    e2e_ms, e2e_wall = timed(step_e2e, n_pairs)
    clk = clocks.stop() if rank == 0 else None
    # ---- per-kernel events over K steps on one worker (separate pass: the timed arms carry no event overhead)
    mg = mgs[0]
    lib, ctx = mg.lib, mg.ctx
    lib.modsgpu_profile_enable(ctx, 1)
    n_prof = min(n_pairs, 32)
    for k in range(n_prof):
        step_value(0, k, False)
    buf = C.create_string_buffer(1 << 16)
    lib.modsgpu_profile_report(ctx, buf, len(buf))
    lib.modsgpu_profile_enable(ctx, 0)
    prof = json.loads(buf.value.decode() or "{}")
    if rank != 0:
        for m in mgs:
            m.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- one pair alone on the GPU (what a single `mods` process sees): wall clock per pair through the host-facing call on
    # ONE context, the two images one after the other and side by side (modsgpu_set_pair_overlap, mods.cpp:234-251)
    single = None
    if rank == 0:
        for m in mgs[1:]:            # only one context stays alive: its host waits spin, as in a single-pair process
            m.close()
        mgs = mgs[:1]
        single = {}
        pa, pb = host[0]
        for name, on in (("sequential_ms", False), ("overlapped_ms", True)):
            mgs[0].set_pair_overlap(on)
            for _ in range(3):
                mgs[0].pair_pipeline(pa.numpy(), pb.numpy(), seed=1, capacity=CAPACITY)
            t0 = time.perf_counter()
            for k in range(8):
                mgs[0].pair_pipeline(pa.numpy(), pb.numpy(), seed=1000 + k, capacity=CAPACITY)
            single[name] = (time.perf_counter() - t0) * 1e3 / 8
        mgs[0].set_pair_overlap(False)
        single["note"] = "host BGR in, verified correspondences out, one context, 8 pairs back to back, wall clock"
    lastr = last.get("r")
    value = world * n_pairs / (dev_ms * 1e-3)
    e2e_value = world * n_pairs / (e2e_ms * 1e-3)
    total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
    # "dominant kernel" = the __global__ function (all template instantiations together) with the largest accumulated time
    # among those with an algorithmic work figure; latency-bound helpers (kind 2) are listed in `kernels` only.  The BOUND
    # of a kernel is the one DESIGN.md section 4 / SURVEY 8(d) assign to it (reported by the library as `kind`: the conv
    # and distance kernels are tensor-bound and counted in flops against the sustained bf16 peak, the image / patch
    # kernels are HBM-bound and counted in algorithmic bytes) -- not whichever fraction happens to look larger.
    groups = {}
    for k, v in prof.items():
        if v["kind"] not in ALG:
            continue
        g = groups.setdefault(k.split("<")[0], {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0, "members": [], "kind": v["kind"]})
        g["ms"] += v["ms"]
        g["launches"] += v["launches"]
        g["flops"] += v["work"] if v["kind"] == 1 else 0.0
        g["bytes"] += v["work"] if v["kind"] == 0 else v.get("bytes", 0.0)
        g["members"].append(k)
    top = max(groups, key=lambda k: groups[k]["ms"], default=None)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tc_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if "bf16_tflops_sustained" in peaks else "fallback 1400 TFLOP/s"
    roofline = None
    if top:
        g = groups[top]
        gbs = g["bytes"] / (g["ms"] * 1e-3) / 1e9
        tfs = g["flops"] / (g["ms"] * 1e-3) / 1e12
        if g["kind"] == 1:
            bound, achieved, peak, unit, src = "tensor", tfs, tc_peak, "TFLOP/s", tc_src
        else:
            bound, achieved, peak, unit, src = "hbm", gbs, hbm_peak, "GB/s", hbm_src
        # DRAM traffic of that kernel from the committed `ncu --set full` capture (one launch, the grid named there)
        traffic, traffic_of = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            for name, rec in tj.items():
                if top.startswith(name) or name.startswith(top):
                    traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
                    traffic_of = "one launch, %s grid %s (%s); algorithmic bytes of that launch %s" % (
                        rec.get("instance", name), rec["grid"], rec["source"], rec.get("algorithmic_bytes"))
                    break
        except (OSError, ValueError, KeyError):
            pass
        roofline = {"kernel": top, "instantiations": sorted(g["members"]), "bound": bound, "achieved": achieved, "peak": peak,
                    "unit": unit, "frac": achieved / peak, "traffic": traffic, "traffic_of": traffic_of, "peak_source": src,
                    "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "tensor_tflops": tfs, "tensor_frac": tfs / tc_peak,
                    "launches": g["launches"], "avg_us": 1e3 * g["ms"] / max(g["launches"], 1),
                    "share_of_kernel_time": g["ms"] / total_kernel_ms,
                    "pairs_profiled": n_prof}
    # whole-stage figures: the CNN stage in flops over ALL kernels of the nets (conv1, tcgen05 convs, heads),
    # the detector and the sampler in their algorithmic bytes, each against the measured peak of its bound
    def stage(prefixes, kind):
        ms = sum(v["ms"] for k, v in prof.items() if k.startswith(prefixes))
        if kind == 1:
            work = sum(v["work"] for k, v in prof.items() if k.startswith(prefixes) and v["kind"] == 1)
            return {"ms_per_pair": ms / n_prof, "gflop_per_pair": work / n_prof / 1e9,
                    "tflops": work / (ms * 1e-3) / 1e12 if ms > 0 else None,
                    "frac_of_tensor_peak": work / (ms * 1e-3) / 1e12 / tc_peak if ms > 0 else None}
        work = sum((v["work"] if v["kind"] == 0 else v.get("bytes", 0.0)) for k, v in prof.items() if k.startswith(prefixes))
        return {"ms_per_pair": ms / n_prof, "mb_per_pair": work / n_prof / 1e6,
                "gbs": work / (ms * 1e-3) / 1e9 if ms > 0 else None,
                "frac_of_hbm_peak": work / (ms * 1e-3) / 1e9 / hbm_peak if ms > 0 else None}
    stages = {"cnn": stage(("k_conv", "k_head", "k_trunk", "k_patch_prep"), 1),
              "detect": stage(("k_gray", "k_blur", "k_response", "k_half", "k_nms", "k_resolve", "k_rank", "k_det", "k_pyr"), 0),
              "sampler": stage(("k_sample", "k_large", "k_smp"), 0),
              "match": stage(("k_pack_desc", "k_dist", "k_select", "k_dup"), 1),
              "ransac": {"ms_per_pair": sum(v["ms"] for k, v in prof.items() if k.startswith(("k_rs", "k_rf"))) / n_prof},
              "all_kernels_ms_per_pair": total_kernel_ms / n_prof}
    kernels = {k: {"ms_per_pair": v["ms"] / n_prof, "launches_per_pair": v["launches"] / n_prof,
                   "share": v["ms"] / total_kernel_ms,
                   "achieved": (v["work"] / (v["ms"] * 1e-3) / ALG[v["kind"]][2]) if v["kind"] in ALG and v["ms"] > 0 else None,
                   "unit": ALG[v["kind"]][1] if v["kind"] in ALG else "latency-bound",
                   "hbm_gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v.get("bytes") and v["ms"] > 0 else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload: WHOLE pairs (every keypoint, the pair's own tentatives, no extrapolation)
        # until >= 12 s of CPU work have been timed, at most 6 pairs (a pair takes ~4 s on the 16-core box)
        threads = os.cpu_count() or 1
        cpu_pair(pairs[0], threads)          # warm-up (library loads, torch thread pool)
        t0 = time.perf_counter()
        fulls, parts, counts = [], None, None
        while len(fulls) < 6 and (time.perf_counter() - t0 < 12.0 or not fulls):
            parts, full, counts = cpu_pair(pairs[len(fulls) % len(pairs)], threads)
            fulls.append(full)
        cpu = {"value": 1.0 / float(np.mean(fulls)), "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "%d whole pairs of the bench workload (full detection, every keypoint through sampler + AffNet + "
                         "OriNet + HardNet++, linear FGINN, duplicate filter, the reference's own degensac on the pair's own "
                         "tentatives); %.1f s of CPU work" % (len(fulls), time.perf_counter() - t0),
               "stage_seconds_last_pair": parts, "last_pair": counts}
    h2d = 2 * W_IMG * H_IMG * 3
    d2h = 4 * 8 * int(lastr["inliers"]) + 9 * 8 + 9 * 4 if lastr else 0
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (nets: fp16 operands, fp32 accumulate); f32 detector/sampler; f64 RANSAC",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": PAIRS_PER_STEP, "distinct_pairs": N_DISTINCT_PAIRS,
                       "workers_per_gpu": nwk,
                       "l2": "flushed before every pair (256 MB memset on the pipeline stream)",
                       "parallelism": ("pairs sharded across %d ranks; one NCCL gather of the verified correspondences" % world) if world > 1 else "single GPU",
                       "last_pair": {k: lastr[k] for k in ("keypoints", "regions", "descriptors", "tentatives", "unique_tentatives", "inliers")} if lastr else None},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d * PAIRS_PER_STEP,
                    "d2h_bytes_per_step": d2h * PAIRS_PER_STEP, "ms_per_step": e2e_ms / args.steps,
                    "note": "h2d/d2h count the API-level buffers of the 16 pairs of a step (two BGR images in, verified "
                            "correspondences + H out per pair)"},
            "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
            "host_cpu_ms_per_pair": host_cpu_ms_per_pair, "host_cores": os.cpu_count(),
            "rank_ms": value_rank_ms, "numa_pin": numa, "single_pair": single,
            "roofline": roofline, "stages": stages, "kernels": kernels, "clocks": clk}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    for m in mgs:
        m.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workers", type=int, default=int(os.environ.get("MODSGPU_BENCH_WORKERS", "0")),
                    help="worker threads (one modsgpu_ctx / CUDA stream each) per GPU; 0 = 16 (measured on B200 with 16 "
                         "host cores: 8 -> 216, 12 -> 221, 16 -> 233, 24 -> 233, 32 -> 230 pairs/s)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
