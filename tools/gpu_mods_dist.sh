#!/bin/bash
# Config 5 on N GPUs: the view-sharded MODS loop on a tilted synthetic pair (NCCL exchange), next to the one-GPU run.
N=${1:-2}
mkdir -p gpurun_out
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from mods_light_zmq_b200 import synth
a = synth.blob_image(seed=91, w=1024, h=768, n_blobs=4000)
Ht = np.array([[0.30, 0.05, 90.0], [-0.02, 0.97, 10.0], [0.0, 0.0, 1.0]])
b = synth.warp_image(a, Ht, noise_seed=5)
np.save("gpurun_out/mods_a.npy", synth.gray_to_bgr(a)); np.save("gpurun_out/mods_b.npy", synth.gray_to_bgr(b))
PY
for n in 1 $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) -m mods_light_zmq_b200.mods_dist gpurun_out/mods_a.npy gpurun_out/mods_b.npy --min-matches 1000000 --time > gpurun_out/mods_dist_N$n.txt 2> gpurun_out/mods_dist_N$n.err
  echo "N=$n rc=$?"; grep steps_done gpurun_out/mods_dist_N$n.txt | cut -c1-300 || tail -c 800 gpurun_out/mods_dist_N$n.err
  [ -s gpurun_out/mods_dist_N$n.txt ] || tail -c 1200 gpurun_out/mods_dist_N$n.err
done
