# The round-end measurement set, as run through `gpurun -- 'bash tools/run_gpu_round.sh'` on one B200
# (outputs under gpurun_out/; the summaries copied into profiles/ are listed in profiles/README.md):
#   GPU parity suite, bench (our arm + reference arm), ncu launch list of the bench command,
#   one `ncu --set full` capture of every kernel family (tools/prof_target.py).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --pairs 32 --batch 16 --no-extra --no-cpu --no-flow-frames > gpurun_out/r2f_ncu_launch_stdout.log 2>&1
PROF_PAIRS=16 ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r2f_full -f python tools/prof_target.py flow hist flowhist > gpurun_out/r2f_ncu_full_stdout.log 2>&1
tail -3 gpurun_out/r2f_pytest.log
