#!/bin/bash
# Runs every GPU test function of a file in its own process (a trapped kernel poisons the CUDA
# context, so isolation keeps the remaining results meaningful). Usage: tools/gpu_isolated_tests.sh FILE LOG
FILE=${1:-tests/test_ops_gpu.py}
LOG=${2:-gpurun_out/ops_isolated.log}
mkdir -p "$(dirname "$LOG")"
: > "$LOG"
names=$(python -m pytest "$FILE" --collect-only -q -m gpu 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
rc=0
for n in $names; do
  echo "=== $n" >> "$LOG"
  timeout 300 python -m pytest "$n" -q -m gpu -x 2>&1 | tail -25 >> "$LOG"
  s=${PIPESTATUS[0]}
  echo "--- exit $s" >> "$LOG"
  [ "$s" != "0" ] && rc=1
done
grep -E "^=== |--- exit|passed|failed|error" "$LOG" | tail -80
exit $rc
