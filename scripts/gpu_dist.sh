#!/bin/bash
N=${1:-2}; TAG=${2:-dist}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
bash scripts/gpu_scale.sh $N $TAG 2>&1 | grep -E "N=|repartition|fused|exit|rror" | head -30
