#!/bin/bash
# A/B of the staged-input inverse kernel (PO_INV_STAGE) and its warp split, C3 and the C4 base shape; then parity
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/stage_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
def g(p):
    v=[k['avg_ms'] for k in ks if k['kernel'].startswith(p)]
    return v[0] if v else float('nan')
print('$1 ms/step %.4f step_frac %.4f  fwd %.4f  c16inv %.4f inv %.4f  pb-of-fwd %.4f'%(d['ms_per_step'],r['step_frac'],g('fused-fwd'),g('pruned-c16 c2c inv'),g('fused-inv'),g('pullback of fused-fwd')))"; }
PO_INV_STAGE=0 run "c3 stage=0" "--workload c3"
PO_INV_STAGE=1 run "c3 stage=1" "--workload c3"
PO_INV_STAGE=1 PO_PIPE2_STREAM_INV=352 run "c3 stage=1 ns=352" "--workload c3"
PO_INV_STAGE=1 PO_PIPE2_STREAM_INV=384 run "c3 stage=1 ns=384" "--workload c3"
PO_INV_STAGE=1 PO_PIPE2_STREAM_INV=416 run "c3 stage=1 ns=416" "--workload c3"
PO_INV_STAGE=0 run "c3 stage=0" "--workload c3"
PO_INV_STAGE=1 run "c3 stage=1" "--workload c3"
PO_INV_STAGE=0 run "c4 stage=0" "--workload c4"
PO_INV_STAGE=1 run "c4 stage=1" "--workload c4"
PO_INV_STAGE=1 PO_PIPE2_STREAM_INV=384 run "c4 stage=1 ns=384" "--workload c4"
PO_INV_STAGE=1 PO_INV_CHUNKS=1 run "c4 stage=1 chunks=1" "--workload c4"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5
