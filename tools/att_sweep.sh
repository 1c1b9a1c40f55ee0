#!/bin/bash
# Correctness matrix + timing of the attention kernels through the C ABI (tools/att_bench.cu). Run under gpurun.
cd "$(dirname "$0")/.."
B=tools/bin/att_bench; L=vfmseg_b200/lib/libvfmseg_b200.so
run() { echo "== $*"; timeout 25 $B "$@"; rc=$?; [ $rc -ne 0 ] && echo "   exit code $rc"; [ $rc -eq 124 ] && { echo "TIMEOUT: abort"; exit 1; }; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
run $L 1 128 1 1,4 3 || exit 1
run $L 2 256 2 1,4 3 || exit 1
run $L 1 1024 2 1,4 3 || exit 1
run $L 3 1025 4 1,2,4,5 3
run $L 2 197 2 1,4,5 3
run $L 1 2049 2 1,4,5 3
run $L 2 17 2 4,5 3
run $L 2 2 1 5 3
run $L 1 385 3 4,5 3
run $L 2 641 2 4,5 3
run $L 20 257 12 4,5 3
run $L 40 65 8 4,5 3
run $L 600 17 1 4,5 3
run $L 300 129 2 4,5 3
run $L 7 1025 16 4,5 3
run $L 2 1025 3 4,5 3 3.0      # peaked scores: rescale path
echo "---- timing, 36 windows x 16 heads x 1025 tokens"
run $L 36 1025 16 1,2,4,5 20
for v in nohand poly4; do
  [ -f vfmseg_b200/lib/libvfm_$v.so ] && run vfmseg_b200/lib/libvfm_$v.so 36 1025 16 4,5 20
done
echo "---- 18 windows"
run $L 18 1025 16 1,4,5 20
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
