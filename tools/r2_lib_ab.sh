#!/bin/bash
# bench.py config 2 with alternative builds of the library: tools/r2_lib_ab.sh <suffix|-> ... ("-" = the shipped library)
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
  lib=""; [ "$k" != "-" ] && lib="VFMSEG_B200_LIB=$PWD/vfmseg_b200/lib/libvfm_$k.so"
  env $lib timeout 200 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2>gpurun_out/lib_$i.err | tail -1 > gpurun_out/lib_$i.json
  line gpurun_out/lib_$i.json "[$k]"
done
