set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r1_scale8_c1_cartpole.log 2> gpurun_out/r1_scale8_c1_cartpole.err
timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 3 --workload c5_bridge_rollout100 --log2-envs 23 --e2e-steps 3 > gpurun_out/r1_scale8_c5_bridge_rollout100_64m.log 2> gpurun_out/r1_scale8_c5.err
timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 10 --workload c4_hetero --log2-envs 21 --e2e-steps 3 > gpurun_out/r1_scale8_c4_hetero_16m.log 2> gpurun_out/r1_scale8_c4.err
timeout 300 $TR bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/r1_scale8_reference.log 2>&1
tail -c 300 gpurun_out/r1_scale8_*.err
grep -H -o '"value": [0-9.e+]*, "unit": "env-steps/s", "n_gpus": [0-9]*' gpurun_out/r1_scale8_*.log
