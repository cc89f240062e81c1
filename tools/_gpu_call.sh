set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/bench_all.sh c1_cartpole c3_acrobot c3_mountaincar c3_pendulum c3_mountaincar_fp64 c3_pendulum_fp64 c3_acrobot_fp64 c4_hetero
cap() {  # workload, kernel regex
  W=$1; K=$2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -s 60 -f -o /tmp/prof_$W python bench.py --workload $W --steps 70 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$W.log 2>&1
  python profiles/summarize.py full /tmp/prof_$W.ncu-rep > gpurun_out/full_${W}_steady.txt 2>&1
  ncu -i /tmp/prof_$W.ncu-rep --page source --csv > gpurun_out/sass_${W}_steady.csv 2>/dev/null
}
cap c3_mountaincar step_kernel
cap c3_pendulum step_kernel
cap c1_cartpole step_kernel
