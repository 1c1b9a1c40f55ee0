#!/bin/bash
# bench.py config 2 under experiment knobs given as "NAME=VAL[,NAME=VAL]" arguments ("-" = defaults)
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    fam = sorted(d["roofline"]["families"].items(), key=lambda kv: -kv[1]["ms_total"])[:7]
    print(sys.argv[2], round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 3), d.get("clocks", {}).get("sm_mhz"),
          " | ".join(f"{k} {v['launches']} {v['ms_total']:.1f}" for k, v in fam))
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
}
i=0
for k in "$@"; do
  i=$((i+1))
  envs=""; [ "$k" != "-" ] && envs=$(echo "$k" | tr ',' ' ')
  env $envs VFM_VERBOSE=1 timeout 200 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2>gpurun_out/knob_$i.err | tail -1 > gpurun_out/knob_$i.json
  line gpurun_out/knob_$i.json "[$k]"; grep -h "vfm:" gpurun_out/knob_$i.err | head -2
done
