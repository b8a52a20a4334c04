#!/bin/bash
# usage: tools/gpu_ncu_multi.sh "regex:skip:count:tag" ...   -- several `ncu --set full` captures of tools/ncu_target.py
mkdir -p gpurun_out/r3
for spec in "$@"; do
  IFS=: read -r rx skip count tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $count -f -o gpurun_out/r3/prof_$tag python tools/ncu_target.py 2 > gpurun_out/r3/ncu_$tag.log 2>&1
  tail -1 gpurun_out/r3/ncu_$tag.log
done
