#!/bin/bash
# folded-LayerNorm schedule: parity tests under every fold mode the engines support, then bench A/B per config
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    fam = sorted(d["roofline"]["families"].items(), key=lambda kv: -kv[1]["ms_total"])[:7]
    print(sys.argv[2], round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d.get("clocks", {}).get("sm_mhz"),
          " | ".join(f"{k} {v['launches']} {v['ms_total']:.1f} {v.get('tflops')}" for k, v in fam))
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
}
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "residual_stats or lnfold" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_e2e_gpu.py -x -q 2>&1 | tail -3
for f in 1 3; do echo "== parity EVA02 / SAM with VFM_LN_FOLD=$f"; VFM_LN_FOLD=$f timeout 600 python -m pytest tests/test_eva_gpu.py tests/test_sam_gpu.py -x -q 2>&1 | tail -3; done
for cfg in 2 4 5; do for f in 1 0 3; do
  VFM_LN_FOLD=$f timeout 200 python bench.py --config $cfg --steps ${1:-20} --warmup 3 --no-cpu-baseline 2>gpurun_out/fold_c${cfg}_$f.err | tail -1 > gpurun_out/fold_c${cfg}_$f.json
  line gpurun_out/fold_c${cfg}_$f.json "config $cfg fold=$f"
done; done
