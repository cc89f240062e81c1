set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/bench_all.sh c5_bridge c5_bridge_rollout32 c5_bridge_rollout100 c5_bridge_split_rollout32 c4_hetero c2_frozenlake8_16m
