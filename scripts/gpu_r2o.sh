#!/bin/bash
# round 2, LAST 1-GPU call (one GPU-minute left): the secondary configs exactly as bench.py's subprocess runs them
mkdir -p gpurun_out
BENCH_CONFIGS_PARTIAL=gpurun_out/r2o_secondary_partial.json timeout 45 python scripts/bench_configs.py > gpurun_out/r2o_secondary.json 2> gpurun_out/r2o_secondary.err
echo "bench_configs rc=$?"; python - <<'PY'
import json, os
p = "gpurun_out/r2o_secondary.json"
if not os.path.getsize(p):
    p = "gpurun_out/r2o_secondary_partial.json"
d = json.load(open(p))
for name, v in d.items():
    print(name, {k: (round(x, 5) if isinstance(x, float) else x) for k, x in v.items() if "kernels" not in k and k not in ("workload", "note", "fused_kernel_variants") and not isinstance(x, dict)})
PY
tail -3 gpurun_out/r2o_secondary.err
