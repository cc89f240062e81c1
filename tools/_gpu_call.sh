set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/bench_all.sh c2_frozenlake8 c2_frozenlake8_16m
