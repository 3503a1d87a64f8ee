#!/bin/bash
# 2 GPUs: distributed tests (peer-store exchange with vectorised copies) + exchange phase trace + 2-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -rs > gpurun_out/r2f_dist_tests.log 2>&1
echo "dist pytest rc=$?" >> gpurun_out/r2f_dist_tests.log; tail -4 gpurun_out/r2f_dist_tests.log
PO_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err
echo "bench rc=$?"; grep "peer exchange" gpurun_out/r2f_bench_2gpu.err | tail -6
python - <<PY
import json
d=json.load(open('gpurun_out/r2f_bench_2gpu.json'))
print('ms_per_step', d['ms_per_step'], d.get('dist_check'))
for k in d['roofline']['kernels']: print('%.4f ms %7.0f GB/s  %s' % (k['avg_ms'], k['GBps'] or 0, k['kernel'][:80]))
PY
