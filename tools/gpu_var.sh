#!/bin/bash
# usage: tools/gpu_var.sh "<workloads>" [variant lib]
for lib in "" $2; do
  echo "== lib: ${lib:-default}"
  for W in $1; do
    S=300; case $W in *rollout*) S=40;; esac
    for rep in 1 2; do
    NSGYM_B200_LIB=$lib python bench.py --workload $W --steps $S --warmup 10 --no-cpu-baseline --no-table --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('%-28s %.3e steps/s frac %.3f %.1f us' % (d['config']['workload'], d['value'], r['frac'], r['kernel_us_per_launch']))"
    done
  done
done
