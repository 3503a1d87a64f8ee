#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for wl in c3 c4; do
for ns in 256 288 320 352 384; do
  PO_PIPE_STREAM_FWD=$ns PO_PIPE_STREAM_INV=$ns python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
print('$wl ns=$ns ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s)'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps']))"
done; done
