#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
run() { timeout 120 python bench.py --workload ${2:-c3} --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
print('$1 ms/step %.3f | '%d['ms_per_step'] + ' '.join('%.4f'%k['avg_ms'] for k in ks[:7]))"; }
PO_L2_HINTS=0 run "hints=0 pf=1"
PO_L2_HINTS=1 run "hints=1 pf=1"
PO_L2_HINTS=1 PO_PIPE2_PREFETCH_INV=0 run "hints=1 pf=0"
PO_L2_HINTS=1 run "hints=1 pf=1 c4" c4
PO_L2_HINTS=0 run "hints=0 pf=1 c4" c4
PO_L2_HINTS=1 PO_PIPE2_PREFETCH_INV=0 run "hints=1 pf=0 c4" c4
