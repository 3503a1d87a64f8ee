#!/bin/bash
# A/B of programmatic dependent launch (PO_PDL) and of the shuffle x-phase combine of the ring kernel (PO_RING_XSHFL)
# on the C3 / C4-base step, then the GPU parity suite with the defaults.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/pdl_build.log 2>&1
run() { timeout 120 python bench.py $2 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
f=[k for k in r['kernels'] if k['kernel'].startswith('fused-fwd')][0]; i=[k for k in r['kernels'] if k['kernel'].startswith('fused-inv')][0]
print('$1 ms/step %.4f step_frac %.4f sum-of-kernels %.4f ms  fwd %.4f inv %.4f'%(d['ms_per_step'],r['step_frac'],sum(k['avg_ms'] for k in r['kernels']),f['avg_ms'],i['avg_ms']))"; }
for rep in 1 2; do
PO_PDL=0 PO_RING_XSHFL=0 run "c3 pdl=0 xshfl=0" "--workload c3"
PO_PDL=1 PO_RING_XSHFL=0 run "c3 pdl=1 xshfl=0" "--workload c3"
PO_PDL=0 PO_RING_XSHFL=1 run "c3 pdl=0 xshfl=1" "--workload c3"
PO_PDL=1 PO_RING_XSHFL=1 run "c3 pdl=1 xshfl=1" "--workload c3"
done
PO_PDL=0 PO_RING_XSHFL=0 run "c4 pdl=0 xshfl=0" "--workload c4"
PO_PDL=1 PO_RING_XSHFL=1 run "c4 pdl=1 xshfl=1" "--workload c4"
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -5
