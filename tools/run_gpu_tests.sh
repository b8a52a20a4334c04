#!/bin/bash
# Runs every GPU parity test in its own process (a hung or faulted kernel then costs one test, not the run).
mkdir -p gpurun_out
LOG=gpurun_out/tests.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
TESTS=$(python -m pytest tests/test_gpu_parity.py -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
for t in $TESTS; do
  echo "=== $t" >> $LOG
  timeout -k 5 ${PER_TEST_TIMEOUT:-240} python -m pytest "$t" -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -${TAIL:-25} >> $LOG
  echo "exit=$?" >> $LOG
done
grep -E "^=== |passed|failed|error|exit=" $LOG | paste - - - | sed 's/tests\/test_gpu_parity.py:://' | tail -40
