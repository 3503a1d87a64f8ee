#!/bin/bash
# round 2, short 1-GPU call: second A/B of the fused C2 kernels' instantiations (compile-time row pitch, 5 CTAs / SM) + their parity tests
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_fused_pair_variants or empty_operands or c2_full_size_vs" > gpurun_out/r2m_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2m_tests.log
timeout 100 python scripts/bench_configs.py C2 > gpurun_out/r2m_c2.json 2> gpurun_out/r2m_c2.err
echo "bench_configs rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_c2.json"))["C2"]
for k, v in d.items():
    if "kernels" in k or k == "workload":
        continue
    print(k, v)
PY
tail -2 gpurun_out/r2m_c2.err
