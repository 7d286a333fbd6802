# Developer A/B on ONE box: library variants tools/ab/<name>.so x environment knobs, 1080p bench without extras.
set -x
mkdir -p gpurun_out
cp scannertools_b200/libscannertools_b200.so /tmp/current.so
for rep in 1 2; do
  for v in chain hybrid; do
    cp tools/ab/$v.so scannertools_b200/libscannertools_b200.so
    python bench.py --steps 12 --warmup 4 --no-cpu --no-flow-frames --no-extra > gpurun_out/abq_${v}_win_$rep.json 2> gpurun_out/abq_${v}_win_$rep.err
    STB_NO_WIN=1 python bench.py --steps 12 --warmup 4 --no-cpu --no-flow-frames --no-extra > gpurun_out/abq_${v}_nowin_$rep.json 2> gpurun_out/abq_${v}_nowin_$rep.err
  done
done
cp /tmp/current.so scannertools_b200/libscannertools_b200.so
python - <<'PY'
import json
for v in ("chain","hybrid"):
  for k in ('win','nowin'):
    for rep in (1,2):
        try:
            d=json.loads(open('gpurun_out/abq_%s_%s_%d.json'%(v,k,rep)).read().strip().splitlines()[-1])
            print(v, k, rep, round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
        except Exception as e:
            print(v, k, rep, 'failed', e)
PY
