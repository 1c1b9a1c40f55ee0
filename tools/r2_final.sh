#!/bin/bash
# round-2 final evidence (one GPU): full test suite, smoke, default bench, configs 3/4/5, ncu launch list + full capture of the attention kernel
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 400 python bench.py > gpurun_out/r2f_bench_config2.json 2> gpurun_out/r2f_bench_config2.err; echo "bench rc=$?"
for c in 3 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_config$c.json 2> gpurun_out/r2f_bench_config$c.err; echo "config $c rc=$?"; done
python - <<'PY'
import json
for c in (2, 3, 4, 5):
    try:
        d = json.loads(open(f"gpurun_out/r2f_bench_config{c}.json").read().strip().splitlines()[-1])
        print(c, d["metric"], round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), "steps", d["steps"], d.get("clocks"), d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("whole_step"))
    except Exception as e:
        print(c, "unreadable", e)
PY
timeout 120 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2f_quick.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2f_launches_run.log 2>&1
L=vfmseg_b200/lib/libvfmseg_b200.so
timeout 60 tools/bin/att_bench $L 36 1025 16 5 5 > gpurun_out/r2f_att_plain.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pp -c 1 -o gpurun_out/r2f_prof_attn_pp -f \
  tools/bin/att_bench $L 36 1025 16 5 1 > gpurun_out/r2f_prof_att_ncu.log 2>&1
tail -1 gpurun_out/r2f_att_plain.log
ls -la gpurun_out/r2f_prof_attn_pp.ncu-rep gpurun_out/r2f_launches.csv
