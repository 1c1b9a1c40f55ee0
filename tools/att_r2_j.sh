B=tools/bin/att_bench
tools/bin/xu_share_bench > gpurun_out/xu_share.log 2>&1
for v in vfmseg_b200 vfm_pk2 vfm_pk2nh vfm_pk2p4 vfm_pk2p2 vfm_pk2p1; do
  L=vfmseg_b200/lib/lib$v.so; [ $v = vfmseg_b200 ] && L=vfmseg_b200/lib/libvfmseg_b200.so
  echo "== $v"; timeout 100 $B $L 36 1025 16 5,4 20
done > gpurun_out/att_r2_j.log 2>&1
for a in "3 1025 4" "2 197 2" "1 2049 2" "2 641 2" "20 257 12"; do echo "== pk2 $a"; timeout 60 $B vfmseg_b200/lib/libvfm_pk2.so $a 4,5 2; done >> gpurun_out/att_r2_j.log 2>&1
echo "== peaked pk2"; timeout 60 $B vfmseg_b200/lib/libvfm_pk2.so 2 1025 3 4,5 2 3.0 >> gpurun_out/att_r2_j.log 2>&1
timeout 300 python tools/bench_attn_libs.py > gpurun_out/attn_libs.log 2>&1
cat gpurun_out/xu_share.log gpurun_out/att_r2_j.log gpurun_out/attn_libs.log
