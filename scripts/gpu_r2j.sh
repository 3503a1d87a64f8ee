#!/bin/bash
# round 2, short 1-GPU call (GPU budget nearly spent): the pipelined host (e2e) schedule + the async host-apply test
mkdir -p gpurun_out
timeout 150 python bench.py --no-cpu --no-secondary --no-weak-base --steps 10 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "step_frac", d["roofline"]["step_frac"])
    print("e2e", json.dumps(d["e2e"])[:900])
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/r2j_bench.err
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_buffer" > gpurun_out/r2j_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2j_tests.log
