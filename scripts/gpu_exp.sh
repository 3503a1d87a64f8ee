#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -k "zmz or pullback" 2>&1 | tail -3
run() { timeout 120 python bench.py --workload ${2:-c3} --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
print('$1 ms/step %.3f step_frac %.3f'%(d['ms_per_step'], r['step_frac']))
for k in ks: print('    %-90s %.4f'%(k['kernel'][:90],k['avg_ms']))"; }
run "zmz=1"
PO_ZMZ=0 run "zmz=0" | head -1
run "zmz=1 c4" c4 | head -1
