#!/bin/bash
# round 2, short 1-GPU call: fused C2 factor-pair kernels -- parity at full size, then timing (fused vs PO_C2_FUSE=0)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_full_size or c2_fused or truncated_dft_kron or pruned_real16" > gpurun_out/r2k_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2k_tests.log
timeout 120 python scripts/bench_configs.py C2 > gpurun_out/r2k_c2.json 2> gpurun_out/r2k_c2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_c2.json"))["C2"]
for k, v in d.items():
    if "kernels" in k:
        print(k, [(r["kernel"][:34], round(r["avg_ms"] * 1e3, 1)) for r in v])
    else:
        print(k, v)
PY
tail -3 gpurun_out/r2k_c2.err
