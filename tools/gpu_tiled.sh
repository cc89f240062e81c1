#!/bin/bash
tag=$1
python -m pytest tests/test_gpu_kernel_variants.py -x -q 2>&1 | tail -5
B="python bench.py --no-cpu-baseline --no-table --steps 200 --warmup 10 --e2e-steps 2"
for wl in c1_cartpole c1_cartpole_fp64 c3_acrobot c3_acrobot_fp64 c3_mountaincar c3_mountaincar_fp64 c3_pendulum c3_pendulum_fp64; do
  NSGYM_B200_NO_TILED=1 $B --workload $wl >> gpurun_out/${tag}_${wl}_plain.jsonl
  $B --workload $wl >> gpurun_out/${tag}_${wl}_tiled.jsonl
  for mb in 3 4 5 6 8; do NSGYM_B200_SPEC_MIN_BLOCKS=$mb $B --workload $wl >> gpurun_out/${tag}_${wl}_tiled_mb$mb.jsonl; done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.jsonl")):
  for l in open(f):
    d=json.loads(l); r=d["roofline"]; print(f.split("/")[-1], "%.3e"%d["value"], "%.1f us"%r["kernel_us_per_launch"], "frac %.3f"%r["frac"], r.get("frac_physical"))
PY
