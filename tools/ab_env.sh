# Developer A/B on ONE box: environment knobs x the 1080p bench without extras.
#   usage: bash tools/ab_env.sh "NAME=ENV=VAL ..." ...   e.g.  bash tools/ab_env.sh "base=" "nowin=STB_NO_WIN=1"
set -x
mkdir -p gpurun_out
for rep in 1 2; do
  for spec in "$@"; do
    name=${spec%%=*}; envs=${spec#*=}
    env $envs python bench.py --steps 12 --warmup 4 --no-cpu --no-flow-frames --no-extra > gpurun_out/abe_${name}_$rep.json 2> gpurun_out/abe_${name}_$rep.err
  done
done
python - "$@" <<'PY'
import json, sys
for spec in sys.argv[1:]:
    name = spec.split('=', 1)[0]
    for rep in (1, 2):
        try:
            d = json.loads(open('gpurun_out/abe_%s_%d.json' % (name, rep)).read().strip().splitlines()[-1])
            print(name, rep, round(d['value'], 1), round(d['e2e']['value'], 1), round(d['roofline']['frac'], 4), round(d['roofline']['avg_launch_us'], 1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
        except Exception as e:
            print(name, rep, 'failed', e)
PY
