# Developer A/B on ONE box: tools/ab/<name>.so variants of the library are swapped in and the 1080p
# bench (no extras) is run alternately; prints value / e2e / roofline frac / level-0 launch us per run.
set -x
mkdir -p gpurun_out
cp scannertools_b200/libscannertools_b200.so /tmp/current.so
for rep in 1 2; do
  for v in "$@"; do
    cp tools/ab/$v.so scannertools_b200/libscannertools_b200.so
    python bench.py --steps 12 --warmup 4 --no-cpu --no-flow-frames --no-extra > gpurun_out/ab_${v}_$rep.json 2> gpurun_out/ab_${v}_$rep.err
  done
done
cp /tmp/current.so scannertools_b200/libscannertools_b200.so
python - "$@" <<'PY'
import json, sys
for v in sys.argv[1:]:
    for rep in (1, 2):
        try:
            d = json.loads(open('gpurun_out/ab_%s_%d.json' % (v, rep)).read().strip().splitlines()[-1])
            print(v, rep, round(d['value'], 1), round(d['e2e']['value'], 1), round(d['roofline']['frac'], 4), round(d['roofline']['avg_launch_us'], 1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
        except Exception as e:
            print(v, rep, 'failed', e)
PY
