#!/bin/bash
# rollouts specialised + occupancy sweeps of the specialised gridworld step kernels + C1 full capture
tag=$1
mkdir -p gpurun_out
export NSGYM_B200_JIT_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_kernel_variants.py -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
B="python bench.py --no-cpu-baseline --no-table --warmup 10 --e2e-steps 2"
for wl in c1_cartpole_rollout32 c5_bridge_rollout32 c3_acrobot_rollout32 c3_acrobot_fp64_rollout32; do
  $B --workload $wl --steps 20 >> gpurun_out/${tag}_${wl}_spec.jsonl 2>> gpurun_out/${tag}.err
  $B --workload $wl --steps 20 --no-specialize >> gpurun_out/${tag}_${wl}_pre.jsonl 2>> gpurun_out/${tag}.err
done
for mb in 4 5 6 7 8; do
  for wl in c5_bridge c2_frozenlake8_16m; do
    NSGYM_B200_SPEC_MIN_BLOCKS=$mb $B --workload $wl --steps 200 >> gpurun_out/${tag}_${wl}_mb${mb}.jsonl 2>> gpurun_out/${tag}.err
  done
done
for mb in 3 4 5 6; do
  for wl in c1_cartpole_fp64 c3_acrobot c3_acrobot_fp64 c3_pendulum_fp64; do
    NSGYM_B200_SPEC_MIN_BLOCKS=$mb $B --workload $wl --steps 200 >> gpurun_out/${tag}_${wl}_mb${mb}.jsonl 2>> gpurun_out/${tag}.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.jsonl")):
    for l in open(f):
        d=json.loads(l); r=d["roofline"]
        print(f.split("/")[-1], "%.3e steps/s" % d["value"], "%.1f us" % r["kernel_us_per_launch"], "frac %.3f" % r["frac"], d["config"]["kernels"][:22], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
# full capture of the specialised C1 kernel (steady state: skip the first launches)
ncu --set full --import-source on --clock-control none -k regex:nsgym_spec_ -s 30 -c 1 -o gpurun_out/${tag}_full_c1_spec -f \
  python bench.py --workload c1_cartpole --no-cpu-baseline --no-table --steps 40 --warmup 3 --e2e-steps 2 > gpurun_out/${tag}_full_c1.log 2>&1
ncu -i gpurun_out/${tag}_full_c1_spec.ncu-rep --page details > gpurun_out/${tag}_full_c1_spec.txt 2>&1
ncu -i gpurun_out/${tag}_full_c1_spec.ncu-rep --page source --csv > gpurun_out/${tag}_full_c1_spec_source.csv 2>&1
