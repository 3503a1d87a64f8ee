#!/bin/bash
# round 2, call C (8 GPUs): every multi-GPU test that needs more than 2 ranks (reference's own 8-rank [2,2,2]->[8,1,1] 10^3 case, one-directional
# 6-rank pattern, back-pressure, distributed FNO at 4 and 8 ranks) + the 8-GPU bench with its distributed-vs-serial-oracle check
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs -k "many_ranks or backpressure or fno_nccl or init_all" --durations=10 > gpurun_out/r2c_dist_tests.log 2>&1
echo "dist pytest rc=$?" >> gpurun_out/r2c_dist_tests.log
tail -25 gpurun_out/r2c_dist_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2c_bench_8gpu.json 2> gpurun_out/r2c_bench_8gpu.err
echo "bench rc=$?"; head -c 1200 gpurun_out/r2c_bench_8gpu.json; grep -i "error\|disagree" gpurun_out/r2c_bench_8gpu.err | head -5
