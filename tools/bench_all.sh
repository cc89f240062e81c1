#!/bin/bash
# Every bench workload with the default step counts (no CPU baseline leg); one JSON line each in
# gpurun_out/bench_all/<workload>.log and a table on stdout.  Run on the GPU box:
#   gpurun -- 'bash tools/bench_all.sh'
mkdir -p gpurun_out/bench_all
WL=${@:-"c1_cartpole c1_cartpole_fp64 c2_frozenlake8 c2_frozenlake8_16m c3_acrobot c3_acrobot_fp64 c3_mountaincar c3_mountaincar_fp64 c3_pendulum c3_pendulum_fp64 c4_hetero c5_bridge c5_bridge_rollout32 c5_bridge_rollout100 c5_bridge_split_rollout32 c1_cartpole_rollout32 c3_acrobot_rollout32 c3_acrobot_fp64_rollout32"}
for W in $WL; do
  S=400; case $W in *rollout100) S=20;; *rollout*) S=50;; esac
  python bench.py --workload $W --steps $S --warmup 20 --no-cpu-baseline --no-table --e2e-steps 5 > gpurun_out/bench_all/$W.log 2> gpurun_out/bench_all/$W.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_all/*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); r = d["roofline"]
            print(f'{d["config"]["workload"]:28s} {d["value"]:.3e} steps/s  frac {r["frac"]:.3f}  {r["kernel_us_per_launch"]:8.1f} us  {r["algorithmic_bytes_per_env_step"]:6.1f} B  e2e {d["e2e"]["value"]:.3e}  clk {d["clocks"]["sm_mhz"]} {d["clocks"]["reasons"]}')
PY
