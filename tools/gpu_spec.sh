#!/bin/bash
# Program-specialised kernels: variant tests, A/B bench lines (specialised vs precompiled), light ncu pass.
# usage: tools/gpu_spec.sh <tag> [workloads...]
tag=$1; shift; wls=${@:-c1_cartpole}
mkdir -p gpurun_out
export NSGYM_B200_JIT_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -x -q > gpurun_out/${tag}_variants.log 2>&1; tail -5 gpurun_out/${tag}_variants.log
for wl in $wls; do
  for i in 1 2; do
    python bench.py --workload $wl --no-cpu-baseline --no-table --steps 300 --warmup 20 --e2e-steps 2 >> gpurun_out/${tag}_${wl}_spec.jsonl 2>> gpurun_out/${tag}_${wl}.err
    python bench.py --workload $wl --no-specialize --no-cpu-baseline --no-table --steps 300 --warmup 20 --e2e-steps 2 >> gpurun_out/${tag}_${wl}_pre.jsonl 2>> gpurun_out/${tag}_${wl}.err
  done
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:'nsgym_spec_|step_kernel' -s 40 -c 3 --csv --log-file gpurun_out/${tag}_${wl}_ncu.csv \
    python bench.py --workload $wl --no-cpu-baseline --no-table --steps 60 --warmup 3 --e2e-steps 2 > gpurun_out/${tag}_${wl}_ncu.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.jsonl")):
    for l in open(f):
        d=json.loads(l); r=d["roofline"]
        print(f.split("/")[-1], "%.3e steps/s" % d["value"], "%.1f us" % r["kernel_us_per_launch"], "frac %.3f" % r["frac"], d["config"]["kernels"][:22], "e2e %.3e" % d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
