#!/bin/bash
# round 2, FINAL 8-GPU call: every multi-GPU test + the scaling ladder N = 8, 4, 2 with the distributed-vs-serial-oracle check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs --durations=5 > gpurun_out/r2i_dist_tests.log 2>&1
echo "dist pytest rc=$?" >> gpurun_out/r2i_dist_tests.log
tail -6 gpurun_out/r2i_dist_tests.log
for n in 8 4 2; do
  PO_PEER_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2i_bench_${n}gpu.json 2> gpurun_out/r2i_bench_${n}gpu.err
  echo "bench N=$n rc=$?"; head -c 400 gpurun_out/r2i_bench_${n}gpu.json; echo; grep "peer exchange" gpurun_out/r2i_bench_${n}gpu.err | tail -2
done
