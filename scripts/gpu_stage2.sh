#!/bin/bash
# A/B of the staged inverse kernel's CTA size (PO_INV_STAGE_THREADS = 512 / 640 / 768; x-phase keeps 4 warps)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/stage2_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
def g(p):
    v=[k['avg_ms'] for k in ks if k['kernel'].startswith(p)]
    return v[0] if v else float('nan')
print('$1 ms/step %.4f step_frac %.4f  fwd %.4f  c16inv %.4f inv %.4f  pb-of-fwd %.4f'%(d['ms_per_step'],r['step_frac'],g('fused-fwd'),g('pruned-c16 c2c inv'),g('fused-inv'),g('pullback of fused-fwd')))"; }
for nt in 512 640 768; do
PO_INV_STAGE_THREADS=$nt run "c3 threads=$nt" "--workload c3"
done
PO_INV_STAGE_THREADS=640 PO_PIPE2_STREAM_INV=480 run "c3 threads=640 ns=480" "--workload c3"
PO_INV_STAGE_THREADS=768 PO_PIPE2_STREAM_INV=608 run "c3 threads=768 ns=608" "--workload c3"
for nt in 512 640 768; do
PO_INV_STAGE_THREADS=$nt run "c3 threads=$nt" "--workload c3"
done
for nt in 512 640 768; do
PO_INV_STAGE_THREADS=$nt run "c4 threads=$nt" "--workload c4"
done
PO_INV_STAGE_THREADS=640 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -3
PO_INV_STAGE_THREADS=768 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short 2>&1 | tail -3
