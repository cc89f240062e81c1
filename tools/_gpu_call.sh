set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/bench_all.sh c4_hetero c1_cartpole_rollout32 c3_acrobot_rollout32 c1_cartpole
