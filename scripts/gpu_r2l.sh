#!/bin/bash
# round 2, short 1-GPU call: C2 with the packed / Hermitian r2c arithmetic, and the C3 step (packed FFT16 butterflies reach the y-DFT and z-mix-z kernels)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_full_size or c2_fused or pruned_real16 or zmz_fused_step or test_dft_sizes" > gpurun_out/r2l_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2l_tests.log
timeout 100 python scripts/bench_configs.py C2 > gpurun_out/r2l_c2.json 2> gpurun_out/r2l_c2.err
echo "bench_configs rc=$?"
timeout 100 python bench.py --no-cpu --no-secondary --no-weak-base --no-e2e --steps 20 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_c2.json"))["C2"]
for k, v in d.items():
    if "kernels" in k:
        print(k, [(r["kernel"][:20], round(r["avg_ms"] * 1e3, 1)) for r in v])
    elif k != "workload":
        print(k, v)
b = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("C3 ms/step", b["ms_per_step"], "step_frac", b["roofline"]["step_frac"])
print([(r["kernel"][:22], round(r["avg_ms"] * 1e3, 1)) for r in b["roofline"]["kernels"]])
PY
