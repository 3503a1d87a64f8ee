#!/bin/bash
# One GPU call: parity tests, smoke, bench line, ncu launch list (each ncu run only after the
# same command exited 0 without ncu).  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -60 > gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${TAG}_ncu.log
tail -5 gpurun_out/${TAG}_smoke.log; tail -15 gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
