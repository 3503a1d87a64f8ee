#!/bin/bash
# quick A/B of the C3/C4 step; extra env comes from the caller
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
pb=[k for k in r['kernels'] if k['kernel'].startswith('pullback of fused-inv')][0]; pf=[k for k in r['kernels'] if k['kernel'].startswith('pullback of fused-fwd')][0]
print('$1 ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s) pb-of-inv %.4f pb-of-fwd %.4f'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps'],pb['avg_ms'],pf['avg_ms']))"; }
run "c3" "--workload c3"
run "c4" "--workload c4"
