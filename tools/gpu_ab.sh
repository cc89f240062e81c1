#!/bin/bash
# A/B helper: bench lines (plain) and a light ncu metrics pass of the step kernel of one workload.
# usage: tools/gpu_ab.sh <tag> <workload> [kernel regex]
tag=$1; wl=$2; pat=${3:-step_kernel}
mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --workload $wl --no-cpu-baseline --no-table --steps 300 --warmup 20 --e2e-steps 2 >> gpurun_out/${tag}_${wl}.jsonl 2>> gpurun_out/${tag}_${wl}.err; done
python bench.py --workload $wl --no-cpu-baseline --no-table --steps 60 --warmup 3 --e2e-steps 2 > gpurun_out/${tag}_${wl}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
  --clock-control none -k regex:$pat -s 40 -c 3 --csv --log-file gpurun_out/${tag}_${wl}_ncu.csv \
  python bench.py --workload $wl --no-cpu-baseline --no-table --steps 60 --warmup 3 --e2e-steps 2 > gpurun_out/${tag}_${wl}_ncu.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/${tag}_${wl}.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print("${wl}", "%.3e steps/s" % d["value"], "%.1f us" % r["kernel_us_per_launch"], "frac %.3f" % r["frac"], "phys", r.get("dram_gbs_from_traffic"), "e2e %.3e" % d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
