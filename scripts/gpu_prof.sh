#!/bin/bash
# ncu --set full capture of selected kernels of the bench step.  usage: gpu_prof.sh TAG KERNEL_REGEX [skip] [count]
TAG=${1:-prof}; KREGEX=${2:-pipe2}; SKIP=${3:-6}; CNT=${4:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s ${SKIP} -c ${CNT} -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_plain.json")); r = d["roofline"]
print("ms/step %.3f step_frac %.3f" % (d["ms_per_step"], r["step_frac"]))
for k in r["kernels"]: print("   %-100s %.4f ms %s GB/s" % (k["kernel"][:100], k["avg_ms"], "%.0f" % k["GBps"] if k["GBps"] else "-"))
PY
