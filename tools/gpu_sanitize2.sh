#!/bin/bash
mkdir -p gpurun_out/r3
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r3/sanitizer_mem.log python -m pytest tests -m gpu -x -q -k "detect_full_size or extract_patches_bit_exact or pair_pipeline_config3 or overlap or device_descriptors or detector_modes" 2>&1 | tail -3
echo "memcheck rc=$?"; tail -2 gpurun_out/r3/sanitizer_mem.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file gpurun_out/r3/sanitizer_race.log python -m pytest tests -m gpu -x -q -k "patches_edge or duplicate_filter or blur_matches or ragged" 2>&1 | tail -3
echo "racecheck rc=$?"; tail -3 gpurun_out/r3/sanitizer_race.log
