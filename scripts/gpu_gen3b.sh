#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g3_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
pb=[k for k in r['kernels'] if k['kernel'].startswith('pullback of fused-inv')][0]
print('$1 ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s) pb-of-inv %.4f'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps'],pb['avg_ms']))"; }
run "gen3 c3" "--workload c3"
run "gen3 c4" "--workload c4"
PO_FUSED_GEN=2 run "gen2 c3" "--workload c3"
PO_DEBUG_INV_SKIP=1 run "inv no-x c3" "--workload c3"
PO_DEBUG_INV_SKIP=2 run "inv no-t c3" "--workload c3"
PO_DEBUG_INV_SKIP=1 PO_PIPE2_STREAM_INV=480 run "inv no-x ns=480 c3" "--workload c3"
PO_DEBUG_INV_SKIP=2 PO_PIPE2_STREAM_INV=192 run "inv no-t ns=192 (10 x-warps) c3" "--workload c3"
