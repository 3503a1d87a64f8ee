#!/bin/bash
# A/B of the chunked (L2-resident) y-inverse + fused inverse pair: PO_INV_CHUNKS = 1 (off), 2, 4, 8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/chunks_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
def g(p): 
    v=[k['avg_ms'] for k in ks if k['kernel'].startswith(p)]
    return v[0] if v else float('nan')
print('$1 ms/step %.4f step_frac %.4f  fwd %.4f  c16inv %.4f inv %.4f  pb:c16 %.4f pb-of-fwd %.4f'%(d['ms_per_step'],r['step_frac'],g('fused-fwd'),g('pruned-c16 c2c inv'),g('fused-inv'),g('pullback of pruned-c16 c2c n'),g('pullback of fused-fwd')))"; }
for ch in 1 2 4 8 1 2 4; do
PO_INV_CHUNKS=$ch run "c3 chunks=$ch" "--workload c3"
done
for ch in 1 2 4; do
PO_INV_CHUNKS=$ch run "c4 chunks=$ch" "--workload c4"
done
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5
