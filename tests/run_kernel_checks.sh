#!/bin/bash
# Runs every kernel check in its own process (bounded by `timeout`), logging to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/kernel_checks.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
for c in $(python tests/kernel_checks.py list); do
  echo "=== $c" >> $LOG
  timeout 180 python tests/kernel_checks.py $c >> $LOG 2>&1
  echo "--- exit $? ($c)" >> $LOG
done
grep -E "^--- exit|^=== " $LOG | paste - - | tee gpurun_out/kernel_checks.summary
