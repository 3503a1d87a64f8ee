#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c2_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py > gpurun_out/configs_r1x.json 2> gpurun_out/configs_r1x.err; tail -3 gpurun_out/configs_r1x.err
python - <<PY
import json
d=json.load(open("gpurun_out/configs_r1x.json"))
for k,v in d.items():
    print(k, {kk:vv for kk,vv in v.items() if not isinstance(vv,list) and kk!="workload"})
    for kk in ("fwd_kernels","adj_kernels","kernels"):
        for r in v.get(kk,[]): print("     %-90s %.4f ms %s"%(r["kernel"][:90], r["avg_ms"], "%.0f GB/s"%r["GBps"] if r["GBps"] else ""))
PY
