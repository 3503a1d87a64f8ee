#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -k "fused_tx or fast_path or c3_full" 2>&1 | tail -3
run() { timeout 120 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
f=[k for k in ks if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in ks if k['kernel'].startswith('fused-inv')][0]
print('$1 ms/step %.3f step_frac %.3f fwd %.4f (%.0f GB/s) inv %.4f (%.0f GB/s)'%(d['ms_per_step'], r['step_frac'], f['avg_ms'], f['GBps'], i['avg_ms'], i['GBps']))"; }
PO_RING_SHARE_X=0 run "share=0 c3" c3
PO_RING_SHARE_X=1 run "share=1 c3" c3
PO_RING_SHARE_X=0 run "share=0 c4" c4
PO_RING_SHARE_X=1 run "share=1 c4" c4
