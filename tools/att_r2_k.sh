B=tools/bin/att_bench
for v in vfmseg_b200 vfm_so0 vfm_bo1 vfm_so0h2; do echo "== $v"; timeout 30 $B vfmseg_b200/lib/lib$v.so 40 65 8 5 3;  timeout 30 $B vfmseg_b200/lib/lib$v.so 36 1025 16 5,5,4 20; timeout 30 $B vfmseg_b200/lib/lib$v.so 18 1025 16 5 20; done > gpurun_out/att_r2_p.log 2>&1
cat gpurun_out/att_r2_p.log
