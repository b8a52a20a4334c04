#!/bin/bash
# Configs 4 and 5 on N GPUs of one box: run through gpurun --gpus N.
#   config 4: 512 synthetic 1024x768 images (generator of config 2, seeds 0..511) extracted on N ranks and on 1 rank,
#             OxAff files compared byte for byte, images/s of both runs
#   config 5: the view-sharded MODS loop on a tilted synthetic pair with the matcher sharded by query rows, next to
#             the one-GPU run and to the rank-0 matcher layout of round 1
N=${1:-8}
NIMG=${2:-512}
mkdir -p gpurun_out/r2 /tmp/c4/img /tmp/c4/outN /tmp/c4/out1
python - <<PY
import sys, os, numpy as np
from multiprocessing import Pool
sys.path.insert(0, ".")
from mods_light_zmq_b200 import synth
def gen(s):
    np.save("/tmp/c4/img/%04d.npy" % s, synth.gray_to_bgr(synth.blob_image(seed=s)))
    return s
with Pool(min(32, os.cpu_count() or 4)) as p:
    p.map(gen, range($NIMG))
open("/tmp/c4/imgs.txt", "w").write("".join("/tmp/c4/img/%04d.npy\n" % s for s in range($NIMG)))
open("/tmp/c4/outsN.txt", "w").write("".join("/tmp/c4/outN/%04d.txt\n" % s for s in range($NIMG)))
open("/tmp/c4/outs1.txt", "w").write("".join("/tmp/c4/out1/%04d.txt\n" % s for s in range($NIMG)))
a = synth.blob_image(seed=91, w=1024, h=768, n_blobs=4000)
Ht = np.array([[0.30, 0.05, 90.0], [-0.02, 0.97, 10.0], [0.0, 0.0, 1.0]])
b = synth.warp_image(a, Ht, noise_seed=5)
np.save("/tmp/c4/mods_a.npy", synth.gray_to_bgr(a)); np.save("/tmp/c4/mods_b.npy", synth.gray_to_bgr(b))
PY
echo "== config 4 on $N ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 -m mods_light_zmq_b200.batch /tmp/c4/imgs.txt /tmp/c4/outsN.txt --quiet 2>/tmp/c4/errN.txt | tail -1 | tee gpurun_out/r2/config4_N$N.txt
echo "== config 4 on 1 rank"
timeout 900 python -m mods_light_zmq_b200.batch /tmp/c4/imgs.txt /tmp/c4/outs1.txt --quiet 2>/tmp/c4/err1.txt | tail -1 | tee gpurun_out/r2/config4_N1.txt
ndiff=0; for f in /tmp/c4/out1/*.txt; do cmp -s $f /tmp/c4/outN/$(basename $f) || ndiff=$((ndiff+1)); done
echo "files $(ls /tmp/c4/out1 | wc -l) / $(ls /tmp/c4/outN | wc -l), differing: $ndiff" | tee -a gpurun_out/r2/config4_N$N.txt
tail -3 /tmp/c4/errN.txt | cut -c1-300
echo "== config 5"
for cfg in "1 " "$N " "$N --rank0-matcher"; do
  set -- $cfg; n=$1; flag=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) -m mods_light_zmq_b200.mods_dist /tmp/c4/mods_a.npy /tmp/c4/mods_b.npy --min-matches 1000000 --time $flag 2>/tmp/c4/err5.txt | grep steps_done | cut -c1-400 | tee -a gpurun_out/r2/config5.jsonl
  [ ${PIPESTATUS[0]} -eq 0 ] || tail -c 600 /tmp/c4/err5.txt
done
