# ncu evidence for round 2 (one GPU): full capture of the ping-pong attention kernel at the bench shape + launch list of the bench step
L=vfmseg_b200/lib/libvfmseg_b200.so
timeout 60 tools/bin/att_bench $L 36 1025 16 5 5 > gpurun_out/r2_prof_plain.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pp -c 1 -o gpurun_out/r2_prof_attn_pp -f \
  tools/bin/att_bench $L 36 1025 16 5 1 > gpurun_out/r2_prof_ncu.log 2>&1
timeout 120 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2_quick.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2_launches_run.log 2>&1
ls -la gpurun_out/r2_prof_attn_pp.ncu-rep gpurun_out/r2_launches.csv; tail -2 gpurun_out/r2_prof_plain.log
