#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp32.py tests/test_gpu_kernel_variants.py tests/test_gpu_native_parity.py tests/test_gpu_golden.py tests/test_gpu_rollout.py -m gpu -q -k "acrobot" 2>&1 | tail -15
for lib in "" /root/repo/ns_gym_b200/_lib/libnsgym_acro3.so; do
  echo "== lib: ${lib:-default}"
  for W in c3_acrobot c3_acrobot_fp64 c3_acrobot_rollout32 c3_acrobot_fp64_rollout32; do
    S=300; case $W in *rollout*) S=40;; esac
    NSGYM_B200_LIB=$lib python bench.py --workload $W --steps $S --warmup 10 --no-cpu-baseline --no-table --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('%-28s %.3e steps/s frac %.3f %.1f us' % (d['config']['workload'], d['value'], r['frac'], r['kernel_us_per_launch']))"
  done
done
