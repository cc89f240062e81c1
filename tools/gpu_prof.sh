#!/bin/bash
# ncu --set full capture of one kernel launch of a bench workload, summarised ON the GPU box
# (the reports exceed what gpurun brings back): gpurun_out/<tag>_full_<wl>.txt and <tag>_sass_<wl>.csv
# usage: tools/gpu_prof.sh <tag> <workload> <kernel regex> <launches to skip> [steps]
tag=$1; wl=$2; pat=$3; skip=${4:-60}; steps=${5:-70}
mkdir -p gpurun_out
cmd="python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline --no-table --e2e-steps 2"
$cmd > gpurun_out/${tag}_plain_${wl}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o /tmp/prof_${wl} $cmd > gpurun_out/${tag}_ncu_${wl}.log 2>&1
python profiles/summarize.py full /tmp/prof_${wl}.ncu-rep > gpurun_out/${tag}_full_${wl}.txt 2>&1
ncu -i /tmp/prof_${wl}.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_sass_${wl}.csv 2>/dev/null
head -40 gpurun_out/${tag}_full_${wl}.txt
