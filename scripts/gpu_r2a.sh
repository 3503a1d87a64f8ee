#!/bin/bash
# round 2, call A (1 GPU): full GPU suite (incl. the full-size C2/C3/C5 vs oracle tests) + default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -rP --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | head -c 3000
