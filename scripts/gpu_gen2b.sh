#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g2_build.log 2>&1
run() { python bench.py $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
print('$1 ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s)'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps']))"; }
for pf in 0 1 2; do for ns in 288 320 352; do
  export PO_PIPE2_STREAM_FWD=$ns PO_PIPE2_STREAM_INV=$ns PO_PIPE2_PREFETCH=$pf
  run "gen2 pf=$pf ns=$ns c3" "--workload c3"
  run "gen2 pf=$pf ns=$ns c4" "--workload c4"
done; done
