#!/bin/bash
# CTA-pair forward kernel: parity at the C4 local shapes + timing at the 8-GPU local shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4_local or fused_tx or fast_path" > gpurun_out/r2d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_tests.log; tail -3 gpurun_out/r2d_tests.log
PO_TRACE=1 timeout 300 python bench.py --shape 16,128,128,64 --steps 20 --warmup 3 --no-e2e --no-cpu --no-weak-base > gpurun_out/r2d_bench_local8.json 2> gpurun_out/r2d_bench_local8.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2d_bench_local8.json'))
print('ms_per_step', d['ms_per_step'])
for k in d['roofline']['kernels']:
    if k['avg_ms'] > 0.1: print('%.4f ms %7.0f GB/s  %s' % (k['avg_ms'], k['GBps'] or 0, k['kernel'][:70]))
PY
