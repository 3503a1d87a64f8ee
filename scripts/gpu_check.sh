#!/bin/bash
# One GPU call: parity tests, smoke, bench line, ncu launch list + one --set full capture of the
# top kernels (each ncu run only after the same command exited 0 without ncu).
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag] [kernel-regex]'
TAG=${1:-r1}
KREGEX=${2:-po_fwd_tx_kernel|po_inv_xt_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s ${SKIP:-6} -c ${COUNT:-2} -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/${TAG}_ncu.log
tail -3 gpurun_out/${TAG}_smoke.log; tail -15 gpurun_out/${TAG}_gpu_tests.log; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    r = d["roofline"]
    print("value %.4g pts/s  ms/step %.3f  step_frac %.3f  dominant %s frac %.3f" % (d["value"], d["ms_per_step"], r["step_frac"], r["kernel"], r["frac"]))
    for k in r["kernels"][:12]:
        print("   %-100s %.4f ms %s GB/s" % (k["kernel"][:100], k["avg_ms"], "%.0f" % k["GBps"] if k["GBps"] else "-"))
    print("e2e", d["e2e"], "cpu", d["cpu_baseline"], "launches", d["gpu_launches"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
