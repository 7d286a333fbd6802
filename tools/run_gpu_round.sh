set -x
python -m pytest tests -m gpu -x -q -k "histogram or shot or pipe or c5 or abi" 2>&1 | tail -3
python tools/quick_bench.py hist
python tools/hist_batch_probe.py
