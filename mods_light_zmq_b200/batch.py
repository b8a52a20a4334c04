"""Batch feature extraction sharded one-image-per-GPU (BASELINE config 4; mirrors extract_features_batch.cpp:56-160).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
      -m mods_light_zmq_b200.batch imfnames.txt out_keys.txt

imfnames.txt / out_keys.txt: one input image / one output file per line (extract_features_batch.cpp:84-101).
Image i is processed by rank i mod N: images are independent, so there is NO data-path collective;
the only exchange is one gather of the per-image region counts to rank 0 at the end (NCCL on GPUs, gloo in the
CPU tests).  Output = OxAff text (`SaveRegionsMichal` text mode: "128\\nN\\n" + `x y a b c d0..d127` per region),
written by the rank that extracted the image; files that already exist are skipped like the reference does
(extract_features_batch.cpp:108-117).
"""
import os
import sys

import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin deal: image i -> rank i mod world."""
    return list(range(rank, n_items, world))


def read_image_bgr(path):
    """8-bit BGR image as cv::imread(path, IMREAD_COLOR) gives it.  .npy (h x w x 3 or h x w uint8) and binary
    PGM/PPM are read natively; other formats need the cv2 wheel."""
    if path.endswith(".npy"):
        a = np.load(path)
    elif path.lower().endswith((".pgm", ".ppm")):
        a = _read_pnm(path)
    else:
        import cv2
        a = cv2.imread(path, cv2.IMREAD_COLOR)
        if a is None:
            raise IOError("could not open or find the image " + path)
        return np.ascontiguousarray(a)
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, axis=2)
    return np.ascontiguousarray(a, np.uint8)


def _read_pnm(path):
    with open(path, "rb") as f:
        data = f.read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    magic, w, h, maxv = toks[0], int(toks[1]), int(toks[2]), int(toks[3])
    if maxv != 255 or magic not in (b"P5", b"P6"):
        raise IOError("only 8-bit binary PGM/PPM are supported: " + path)
    ch = 1 if magic == b"P5" else 3
    a = np.frombuffer(data, np.uint8, count=w * h * ch, offset=pos + 1).reshape(h, w, ch)
    return a[:, :, 0] if ch == 1 else a[:, :, ::-1]   # PPM is RGB; cv::imread gives BGR


def gpu_extractor(device):
    """The product path: modsgpu_image_from_bgr8 + modsgpu_extract_features on `device`."""
    import mods_light_zmq_b200 as M
    mg = M.ModsGpu(device, load_nets=True)

    def run(bgr):
        img = mg.image_from_bgr8(bgr)
        try:
            return mg.extract_features(img)
        finally:
            img.free()
    run.close = mg.close
    return run


def extract_features_batch(img_fnames, out_fnames, extractor, rank=0, world=1, dist=None, log=None, fmt=None):
    """Process this rank's share; returns (on rank 0) the list of region counts per image, -1 for skipped/failed.
    extractor: one callable, or a list of callables (one modsgpu context each): the rank's images are then dealt to as
    many threads, so that the host side of one image (read, H2D, file writing) overlaps the kernels of another."""
    import mods_light_zmq_b200 as M
    if len(img_fnames) != len(out_fnames):
        raise ValueError("Length of input and output file lists are not equal %d %d" % (len(img_fnames), len(out_fnames)))
    n = len(img_fnames)
    counts = np.full(n, -2, np.int64)          # -2: not mine
    mine = shard_indices(n, rank, world)
    extractors = list(extractor) if isinstance(extractor, (list, tuple)) else [extractor]

    def work(t):
        _extract_some(mine[t::len(extractors)], img_fnames, out_fnames, extractors[t], counts, log, fmt)
    if len(extractors) == 1:
        work(0)
    else:
        import threading
        errs = []

        def guarded(t):
            try:
                work(t)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=guarded, args=(t,)) for t in range(len(extractors))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise errs[0]
    if dist is not None and world > 1:
        import torch
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(counts).to(dev)
        lst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, lst, dst=0)
        if rank != 0:
            return None
        allc = torch.stack(lst).cpu().numpy()
        counts = allc.max(axis=0)               # every image has exactly one owner (>= -1), the rest say -2
    return counts.tolist()


def _extract_some(indices, img_fnames, out_fnames, extractor, counts, log, fmt):
    import mods_light_zmq_b200 as M
    for i in indices:
        out = out_fnames[i]
        if os.path.exists(out) or os.path.exists(out + "ZMQ"):
            counts[i] = -1                      # "exists, skip"
            continue
        try:
            bgr = read_image_bgr(img_fnames[i])
        except (IOError, OSError, ValueError) as e:
            if log:
                log("Could not open or find the image %s (%s)" % (img_fnames[i], e))
            counts[i] = -1
            continue
        feats = extractor(bgr)
        M.write_regions(out, feats, fmt)    # .npz -> npz, else OxAff unless fmt says "text"
        counts[i] = len(feats)
        if log:
            log("%d %s %s %d" % (i, img_fnames[i], out, len(feats)))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    contexts, quiet = 4, False
    if "--contexts" in argv:
        k = argv.index("--contexts")
        contexts = max(1, int(argv[k + 1]))
        del argv[k:k + 2]
    if "--quiet" in argv:
        argv.remove("--quiet")
        quiet = True
    if len(argv) < 2:
        print("Usage: python -m mods_light_zmq_b200.batch imfnames.txt out_keys.txt [--contexts K] [--quiet]", file=sys.stderr)
        return 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    imgs = [l.rstrip("\n") for l in open(argv[0]) if l.strip()]
    outs = [l.rstrip("\n") for l in open(argv[1]) if l.strip()]
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import time
    exts = [gpu_extractor(local_rank) for _ in range(contexts)]        # K contexts (streams) per GPU
    # every context runs this rank's first image three times before the clock starts: workspace growth, the detector's
    # graph capture and the first-use costs of the kernels are not part of the batch's throughput
    mine0 = shard_indices(len(imgs), rank, world)
    if mine0:
        try:
            warm = read_image_bgr(imgs[mine0[0]])
            for e in exts:
                for _ in range(3):
                    e(warm)
        except (IOError, OSError, ValueError):
            pass
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    counts = extract_features_batch(imgs, outs, exts, rank, world, dist, log=None if quiet else (lambda s: print(s, flush=True)))
    sec = time.perf_counter() - t0
    for e in exts:
        e.close()
    if rank == 0:
        done = [c for c in counts if c >= 0]
        print("images %d, extracted %d, skipped %d, regions %d, %.2f s, %.1f images/s on %d rank(s) x %d contexts" %
              (len(counts), len(done), len(counts) - len(done), sum(done), sec, len(done) / max(sec, 1e-9), world, contexts))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
