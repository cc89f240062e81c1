#!/bin/bash
# one ncu --set full capture of the (specialised) step kernel of a workload, summary + source page brought back
# usage: tools/gpu_full.sh <tag> <workload> [kernel regex]
tag=$1; wl=$2; pat=${3:-nsgym_spec_}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:$pat -s 30 -c 1 -o gpurun_out/${tag}_full_${wl} -f \
  python bench.py --workload $wl --no-cpu-baseline --no-table --steps 40 --warmup 3 --e2e-steps 2 > gpurun_out/${tag}_full_${wl}.log 2>&1
ncu -i gpurun_out/${tag}_full_${wl}.ncu-rep --page details > gpurun_out/${tag}_full_${wl}.txt 2>&1
ncu -i gpurun_out/${tag}_full_${wl}.ncu-rep --page source --csv > gpurun_out/${tag}_full_${wl}_source.csv 2>&1
rm -f gpurun_out/${tag}_full_${wl}.ncu-rep
