#!/bin/bash
# one GPU session of round 2: tests, bench, launch list of one warm pair, full captures of the kernels named in $1
# usage (through gpurun): bash tools/gpu_session_r2.sh "k_trunk k_rsb_lo" tag
KERNELS=${1:-k_trunk}
TAG=${2:-r2}
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_${TAG}.json 2> gpurun_out/r2/bench_${TAG}.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2/bench_${TAG}.err
MODSGPU_HOST_PROFILE=1 timeout 120 python tools/ncu_target.py 6 2> gpurun_out/r2/hostprof_${TAG}.txt | tail -1
tail -12 gpurun_out/r2/hostprof_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/launches_${TAG}.csv python tools/ncu_target.py 3 > gpurun_out/r2/ncu_list_${TAG}.log 2>&1
for k in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 6 -f -o gpurun_out/r2/prof_${k}_${TAG} python tools/ncu_target.py 2 > gpurun_out/r2/ncu_${k}_${TAG}.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/r2 | tail -12
