# full GPU suite + bench of the current tree (one gpurun call)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log; grep -h "RAW label\|resize_argmax" gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --gpu-comparator > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench_err.log; echo "bench rc=$?"
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 10 > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c${c}_err.log; echo "bench c$c rc=$?"; done
python - <<'PY'
import json
for f in ('r2_bench','r2_bench_c3','r2_bench_c4','r2_bench_c5'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, d['metric'], 'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'steps',d['steps'],'clocks',d['clocks'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],4))
    if f=='r2_bench':
        for k,v in d['roofline']['families'].items(): print('   ',k,v)
        print(d['roofline']['whole_step']); print(d.get('gpu_comparator'))
PY
