#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g3_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -8 > gpurun_out/g3_tests.log; cat gpurun_out/g3_tests.log
run() { timeout 120 python bench.py $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
pb=[k for k in r['kernels'] if k['kernel'].startswith('pullback of fused-inv')][0]
print('$1 ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s) pb-of-inv %.4f'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps'],pb['avg_ms']))"; }
for gen in 2 3; do
  export PO_FUSED_GEN=$gen
  run "gen$gen c3" "--workload c3"
  run "gen$gen c4" "--workload c4"
done
