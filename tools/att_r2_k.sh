B=tools/bin/att_bench
timeout 200 bash tools/att_sweep.sh > gpurun_out/att_r2_l.log 2>&1
for v in h1 pk2 p4; do echo "== $v"; timeout 30 $B vfmseg_b200/lib/libvfm_$v.so 36 1025 16 5,4 20; done >> gpurun_out/att_r2_l.log 2>&1
timeout 30 $B vfmseg_b200/lib/libvfm_tr.so 36 1025 16 5 3 > gpurun_out/att_trace_l5.log 2>&1
cat gpurun_out/att_r2_l.log; head -30 gpurun_out/att_trace_l5.log
