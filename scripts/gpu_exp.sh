#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1

run() { timeout 120 python bench.py --shape $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
ks=r['kernels']
f=[k for k in ks if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in ks if k['kernel'].startswith('fused-inv')][0]
print('$1 ms/step %.3f step_frac %.3f fwd %.4f (%.0f GB/s) inv %.4f (%.0f GB/s)'%(d['ms_per_step'], r['step_frac'], f['avg_ms'], f['GBps'], i['avg_ms'], i['GBps']))"; }
PO_TILE2=0 run "tile2=0 local-N2 (32,64,128,64)" 32,64,128,64
PO_TILE2=1 run "tile2=1 local-N2 (32,64,128,64)" 32,64,128,64
PO_TILE2=0 run "tile2=0 local-N8 (16,128,128,64)" 16,128,128,64
PO_TILE2=1 run "tile2=1 local-N8 (16,128,128,64)" 16,128,128,64
PO_TILE2=1 PO_PIPE2_PREFETCH_INV=0 run "tile2=1 pf=0 local-N8" 16,128,128,64
