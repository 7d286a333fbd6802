set -x
python -m pytest tests -m gpu -x -q -k "farneback or flow or window or graph or c5" 2>&1 | tail -3
bash tools/ab_run.sh old new
