# usage: tools/r2_scale.sh N "configs" [extra flags]   (run under gpurun --gpus N)
N=$1; CFGS=$2; shift 2
for c in $CFGS; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + c)) bench.py --gpus $N --config $c --steps 20 --warmup 3 "$@" \
    > gpurun_out/r2_scale${N}_c$c$(echo "$@" | tr -d ' -').json 2> gpurun_out/r2_scale${N}_c$c.err
  echo "N=$N config $c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_scale*_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['metric'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], d['config']['parallelism'][-60:])
    except Exception as e: print(f, 'unreadable', e)
PY
