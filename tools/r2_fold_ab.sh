#!/bin/bash
# A/B of the folded-LayerNorm schedule: VFM_LN_FOLD bit 0 = norm1 (fc2 -> qkv), bit 1 = norm2 (proj -> fc1); 0 = LayerNorm kernels.
# usage: tools/r2_fold_ab.sh [steps] [modes...]
steps=${1:-20}; shift
modes=${@:-3 1 0}
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "residual_stats or lnfold" 2>&1 | tail -5
for m in 1 2; do echo "== VFM_RS_MODE=$m (1 deep, 2 serial)"; VFM_RS_MODE=$m timeout 120 python tools/bench_kernels.py --crops 36 --only stats --iters 10 2>&1 | grep -E "stats|lnfold" | cut -c1-130; done
for f in $modes; do
  VFM_LN_FOLD=$f timeout 200 python bench.py --steps $steps --warmup 3 --no-cpu-baseline 2>gpurun_out/fold$f.err | tail -1 > gpurun_out/fold$f.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/fold$f.json"))
    print("fold=$f", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d.get("clocks", {}).get("sm_mhz"))
    for k, v in sorted(d["roofline"]["families"].items(), key=lambda kv: -kv[1]["ms_total"])[:10]:
        print("   ", k, v["launches"], v["ms_total"], v["share"], v.get("tflops"))
except Exception as e:
    print("fold=$f failed:", e); print(open("gpurun_out/fold$f.err").read()[-1500:])
PY
done
