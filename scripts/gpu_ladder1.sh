#!/bin/bash
# single-GPU A/B of the C4 ladder shapes (no NCCL): workload c4 at N=1 plus wide-row kernels via the 2-rank shape on 1 GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/lad_build.log 2>&1
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/lad_c4_n1.json 2> gpurun_out/lad_c4_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/lad_c4_n1.json")); r=d["roofline"]
print("c4@N=1 value %.4g ms/step %.3f step_frac %.3f"%(d["value"],d["ms_per_step"],r["step_frac"]))
seen=set()
for k in r["kernels"]:
    if k["kernel"][:30] not in seen: seen.add(k["kernel"][:30]); print("   %-90s %.4f ms %s"%(k["kernel"][:90],k["avg_ms"],"%.0f"%k["GBps"] if k["GBps"] else "-"))
PY
