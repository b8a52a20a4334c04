#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the kernels touched this round (small cases; ~10-50x slower)
mkdir -p gpurun_out/r3
K=${1:-"blur_matches or ragged or patches_edge or duplicate_filter or detect_idempotent or hamming"}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r3/sanitizer.log python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -5
echo "rc=$?"; grep -c "Invalid\|out of bounds\|Uninit" gpurun_out/r3/sanitizer.log; tail -5 gpurun_out/r3/sanitizer.log
