#!/bin/bash
# ncu launch list of one warm pair + optional full captures: NCU="regex1 regex2", NCU_KSKIP / NCU_KCOUNT
mkdir -p gpurun_out
SKIP=${NCU_SKIP:-190}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c ${NCU_COUNT:-260} --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 2 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
for k in $NCU; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-0} -c ${NCU_KCOUNT:-2} -f -o gpurun_out/prof_$k python tools/ncu_target.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
