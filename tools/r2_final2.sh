#!/bin/bash
# round-2 closing evidence (one GPU, ~3.5 min): full GPU test suite, smoke, default bench (config 2), config 3 (new stage-1 merge kernel),
# ncu full captures of the two class-major merge kernels
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -3 gpurun_out/r2h_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -1 gpurun_out/r2h_smoke.log
timeout 120 python bench.py > gpurun_out/r2h_bench_config2.json 2> gpurun_out/r2h_bench_config2.err; echo "bench rc=$?"
timeout 60 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_config3.json 2> gpurun_out/r2h_bench_config3.err; echo "config 3 rc=$?"
python - <<'PY'
import json
for c in (2, 3):
    try:
        d = json.loads(open(f"gpurun_out/r2h_bench_config{c}.json").read().strip().splitlines()[-1])
        fam = d["roofline"].get("families", {})
        print(c, d["metric"], round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d.get("clocks"),
              d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"],
              {k: round(v["ms_total"] / v["launches"], 4) for k, v in fam.items() if "merge" in k})
    except Exception as e:
        print(c, "unreadable", e)
PY
timeout 90 ncu --set full --clock-control none --import-source on -k regex:merge_class -s 3 -c 1 -o gpurun_out/r2h_prof_slide_merge_class -f \
  python tools/bench_merge.py 0 none > gpurun_out/r2h_prof_merge_ncu.log 2>&1; echo "ncu slide rc=$?"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:ms_merge_class -s 3 -c 1 -o gpurun_out/r2h_prof_ms_merge_class -f \
  python tools/bench_merge.py none 0 > gpurun_out/r2h_prof_msmerge_ncu.log 2>&1; echo "ncu ms rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
