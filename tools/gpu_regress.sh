#!/bin/bash
# full regression on the GPU box: all GPU tests, default bench line (+ reference arm), traffic / metrics of
# every workload, every workload's bench line.  usage: tools/gpu_regress.sh <tag>
tag=$1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.log 2>&1
python tools/collect_traffic.py > gpurun_out/${tag}_traffic.log 2>&1; tail -3 gpurun_out/${tag}_traffic.log
bash tools/bench_all.sh > gpurun_out/${tag}_bench_all.txt 2>&1; cat gpurun_out/${tag}_bench_all.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-table > gpurun_out/${tag}_launches.log 2>&1
for wl in c5_bridge c2_frozenlake8_16m c4_hetero; do bash tools/gpu_full.sh $tag $wl; done
