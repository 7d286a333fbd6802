set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2u_pytest.log
python tools/quick_bench.py flowhist > gpurun_out/r2u_quick.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu --no-flow-frames > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
cat gpurun_out/r2u_pytest.log gpurun_out/r2u_quick.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['clocks'])
ex=d.get('extra',{})
for k,v in ex.items():
    print(k, v.get('value'), v.get('roofline',{}).get('frac'))
PY
