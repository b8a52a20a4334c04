#!/bin/bash
# usage: tools/gpu_round.sh TAG [pytest -k expression]   -- GPU tests, then the driver's bench command; outputs under gpurun_out/r3
TAG=${1:-x}; KEXPR=${2:-}
mkdir -p gpurun_out/r3
if [ -n "$KEXPR" ]; then python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -15; else python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/bench_$TAG.json 2> gpurun_out/r3/bench_$TAG.err
python tools/show_bench.py gpurun_out/r3/bench_$TAG.json 2>/dev/null || tail -3 gpurun_out/r3/bench_$TAG.err
