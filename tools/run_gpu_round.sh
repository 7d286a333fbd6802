set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err
tail -c 300 gpurun_out/r2f_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('copy_only'), d['roofline']['frac'], d['clocks'])
for k,v in d.get('extra',{}).items():
    if 'value' in v: print(k, v.get('value'), v.get('roofline',{}).get('frac'), v.get('e2e',{}).get('value'), v.get('shard_check'))
    else:
        for kk,vv in v.items(): print(k,kk,vv.get('value'))
print(d.get('shard_check'))
PY
