#!/bin/bash
# 1 GPU: ParTensor / C5 parity + secondary configs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rP -k "c5_full_size or kron_of_matrices or matrix_diagonal or par_tensor" > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?"; grep -E "vs oracle|passed|failed" gpurun_out/r2g_tests.log | tail -8
timeout 900 python scripts/bench_configs.py > gpurun_out/r2g_configs.json 2> gpurun_out/r2g_configs.err
echo "configs rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_configs.json'))
for k,v in d.items():
    print(k, {a:b for a,b in v.items() if not a.endswith('kernels') and not a.startswith('kernels') and a!='workload'})
for k in d['C3ii_ParTensor']['kernels']: print(k)
PY
