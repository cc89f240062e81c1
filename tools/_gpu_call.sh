set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
L=$PWD/ns_gym_b200/_lib
bash tools/bench_all.sh c2_frozenlake8 c2_frozenlake8_16m
for v in g8; do for W in c2_frozenlake8_16m c2_frozenlake8; do echo "== $v $W"; NSGYM_B200_LIB=$L/libnsgym_b200_$v.so python bench.py --workload $W --no-cpu-baseline --e2e-steps 2 | grep -o '"value": [0-9.e+]*, "unit": "env-steps/s", "n_gpus"'; done; done
python bench.py --workload c2_frozenlake8_16m --no-cpu-baseline --e2e-steps 2 | python -c "import sys,json; print(json.dumps(json.loads(sys.stdin.read())['roofline'], indent=1))"
