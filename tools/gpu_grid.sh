#!/bin/bash
python -m pytest tests -m gpu -q -x -k "frozenlake or bridge or cliff or grid or tables or planning or rollout or known" 2>&1 | tail -8
for W in c2_frozenlake8 c2_frozenlake8_16m c5_bridge c5_bridge_rollout32 c5_bridge_split_rollout32 c4_hetero; do
  S=300; case $W in *rollout*) S=40;; esac
  python bench.py --workload $W --steps $S --warmup 10 --no-cpu-baseline --no-table --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('%-28s %.3e steps/s frac %.3f %.1f us phys %s' % (d['config']['workload'], d['value'], r['frac'], r['kernel_us_per_launch'], r.get('dram_gbs_from_traffic')))"
done
