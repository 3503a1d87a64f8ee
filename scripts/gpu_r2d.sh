#!/bin/bash
# round 2, call D (1 GPU): CTA-pair forward kernel: parity at the C4 local shapes + A/B timing against the single-tile kernels
mkdir -p gpurun_out
echo skip tests > gpurun_out/r2d_tests2.log
echo "pytest rc=$?" >> gpurun_out/r2d_tests.log; tail -4 gpurun_out/r2d_tests.log
for pair in 1; do
  PO_PAIR=$pair PO_TRACE=1 timeout 300 python bench.py --shape 16,128,128,64 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2d_bench_local8_pair$pair.json 2> gpurun_out/r2d_bench_local8_pair$pair.err
  echo "bench pair=$pair rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2d_bench_local8_pair$pair.json'))
print('ms_per_step', d['ms_per_step'])
for k in d['roofline']['kernels']: print('%.4f ms %7.0f GB/s  %s' % (k['avg_ms'], k['GBps'] or 0, k['kernel'][:70]))
PY
done
grep "gen-4\|gen-2" gpurun_out/r2d_bench_local8_pair1.err | sort | uniq -c | head
