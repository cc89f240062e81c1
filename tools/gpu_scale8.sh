#!/bin/bash
# 8-GPU bench lines with the round-2 kernels (one box): C1 default, C4 16 M envs, Bridge single step + K = 100 rollout
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
timeout 300 $TR 29521 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2_scale8_c1.log 2> gpurun_out/r2_scale8_c1.err
timeout 300 $TR 29522 bench.py --gpus 8 --steps 200 --warmup 10 --workload c4_hetero --log2-envs 21 --e2e-steps 3 --no-table > gpurun_out/r2_scale8_c4.log 2> gpurun_out/r2_scale8_c4.err
timeout 300 $TR 29523 bench.py --gpus 8 --steps 200 --warmup 10 --workload c5_bridge --e2e-steps 3 --no-table > gpurun_out/r2_scale8_c5.log 2> gpurun_out/r2_scale8_c5.err
timeout 300 $TR 29524 bench.py --gpus 8 --steps 20 --warmup 3 --workload c5_bridge_rollout100 --log2-envs 23 --e2e-steps 3 --no-table > gpurun_out/r2_scale8_c5r.log 2> gpurun_out/r2_scale8_c5r.err
grep -H -o '"value": [0-9.e+]*, "unit": "env-steps/s", "n_gpus": [0-9]*' gpurun_out/r2_scale8_*.log
