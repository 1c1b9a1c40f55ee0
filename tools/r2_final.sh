#!/bin/bash
# round-2 final evidence (one GPU): full test suite, smoke, default bench, configs 3/4/5, ncu launch list + full captures of the new GEMM epilogues
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 400 python bench.py > gpurun_out/r2f_bench_config2.json 2> gpurun_out/r2f_bench_config2.err; echo "bench rc=$?"
for c in 3 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2f_bench_config$c.json 2> gpurun_out/r2f_bench_config$c.err; echo "config $c rc=$?"; done
python - <<'PY'
import json
for c in (2, 3, 4, 5):
    try:
        d = json.loads(open(f"gpurun_out/r2f_bench_config{c}.json").read().strip().splitlines()[-1])
        print(c, d["metric"], round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), "steps", d["steps"], d.get("clocks"), d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("whole_step"))
    except Exception as e:
        print(c, "unreadable", e)
PY
# ncu: launch list of the step (serialised, cold: shares only), then full captures of the folded-LayerNorm GEMMs at the bench shape
timeout 120 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2f_quick.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2f_launches_run.log 2>&1
timeout 120 python tools/bench_kernels.py --crops 36 --only "stats" --iters 1 --warmup 1 > gpurun_out/r2f_kernels_plain.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -c 8 -o gpurun_out/r2f_prof_lnfold -f \
  python tools/bench_kernels.py --crops 36 --only "stats" --iters 1 --warmup 1 > gpurun_out/r2f_prof_ncu.log 2>&1
timeout 120 python tools/bench_kernels.py --crops 36 --only "gemm" --iters 10 2>&1 | cut -c1-140 | tee gpurun_out/r2f_kernels.log
ls -la gpurun_out/r2f_prof_lnfold.ncu-rep gpurun_out/r2f_launches.csv
