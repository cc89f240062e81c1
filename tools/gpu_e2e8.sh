#!/bin/bash
# e2e (host-buffer step) on all 8 GPUs of one box against the number of chunks per step
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
p=29540
for c in 2 4 8 16; do
  p=$((p+1))
  timeout 200 $TR $p bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 25 --chunks $c --no-table --no-cpu-baseline > gpurun_out/r2_e2e8_c$c.log 2> gpurun_out/r2_e2e8_c$c.err
done
python - <<PY
import json
for c in (2,4,8,16):
    for l in open(f"gpurun_out/r2_e2e8_c{c}.log"):
        if l.startswith("{"):
            d=json.loads(l); e=d["e2e"]; print("chunks", c, "e2e %.3e"%e["value"], "%.1f GB/s"%e["gbs"], "ceiling d2h %.1f"%e["pcie_ceiling_gbs"]["d2h"], "frac %.3f"%e["frac_of_pcie_ceiling"])
PY
