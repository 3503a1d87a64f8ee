#!/bin/bash
# round 2, call B (2 GPUs): multi-GPU tests that need <= 2 ranks (new peer-store protocol: acks, timeout, init_all) + ParTensor GPU tests + 2-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -rs > gpurun_out/r2b_dist_tests.log 2>&1
echo "dist pytest rc=$?" >> gpurun_out/r2b_dist_tests.log
tail -8 gpurun_out/r2b_dist_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rP -k "par_tensor" > gpurun_out/r2b_tensor_tests.log 2>&1
echo "tensor pytest rc=$?" >> gpurun_out/r2b_tensor_tests.log
grep -E "rel-L2|passed|failed|rc=" gpurun_out/r2b_tensor_tests.log | tail
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err
echo "bench rc=$?"; head -c 1500 gpurun_out/r2b_bench_2gpu.json; tail -3 gpurun_out/r2b_bench_2gpu.err
