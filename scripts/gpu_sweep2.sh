#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
run() { python bench.py $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
print('$1 ms/step %.3f fwd %.4f ms (%.0f GB/s) inv %.4f ms (%.0f GB/s)'%(d['ms_per_step'],f['avg_ms'],f['GBps'],i['avg_ms'],i['GBps']))"; }
for rows in 0 1; do
  export PO_FWD_ROWS=$rows
  run "rows=$rows c3" "--workload c3"
  run "rows=$rows c4(64^4)" "--workload c4"
  run "rows=$rows nx128 (128,64,32,64)" "--shape 128,64,32,64"
  run "rows=$rows nx128 nt32 (128,64,64,32)" "--shape 128,64,64,32"
done
