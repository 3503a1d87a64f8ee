#!/bin/bash
# Usage: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scale.sh N tag'
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
echo "exit $?" >> gpurun_out/${TAG}_n${N}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_n${N}.json"))
    r = d["roofline"]
    print("N=%d value %.4g pts/s  ms/step %.3f  step_frac %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], r["step_frac"]))
    for k in r["kernels"]:
        print("   %-100s %.4f ms" % (k["kernel"][:100], k["avg_ms"]))
    print("e2e", d["e2e"])
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/${TAG}_n${N}.err
