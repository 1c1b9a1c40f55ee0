# full GPU suite + bench of the current tree (one gpurun call)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench_err.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'clocks',d['clocks'])
for k,v in d['roofline']['families'].items(): print(k,v)
print(d['roofline']['whole_step'])
PY
