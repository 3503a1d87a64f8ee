#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
run() { timeout 120 python bench.py --workload ${2:-c3} --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
print('$1 ms/step %.3f step_frac %.3f | '%(d['ms_per_step'], r['step_frac']) + ' '.join('%.4f'%k['avg_ms'] for k in ks[:5]))"; }
PO_C16_PAR=0 run "c16par=0"
PO_C16_PAR=1 run "c16par=1"
PO_C16_PAR=0 run "c16par=0 c4" c4
PO_C16_PAR=1 run "c16par=1 c4" c4
