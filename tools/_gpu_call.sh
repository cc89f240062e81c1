set -x
cap() {  # workload, kernel regex
  W=$1; K=$2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -s 5 -f -o /tmp/prof_$W python bench.py --workload $W --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$W.log 2>&1
  python profiles/summarize.py full /tmp/prof_$W.ncu-rep > gpurun_out/full_$W.txt 2>&1
  ncu -i /tmp/prof_$W.ncu-rep --page source --csv > gpurun_out/sass_$W.csv 2>/dev/null
}
for W in c5_bridge c5_bridge_rollout32; do
  python bench.py --workload $W --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r1e_bench_$W.log 2>&1
  NSGYM_B200_LIB=$PWD/ns_gym_b200/_lib/libnsgym_b200_mb6.so python bench.py --workload $W --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r1e_bench_mb6_$W.log 2>&1
done
cap c1_cartpole step_kernel
grep -H -o '"value": [0-9.e+]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/r1e_bench_*.log
