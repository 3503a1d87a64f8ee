#!/bin/bash
# ncu --set full of the two wide-row kernels (CTA-pair forward, single-tile inverse) at the 8-GPU local shape 128x128x16x64
mkdir -p gpurun_out
CMD="python bench.py --shape 16,128,128,64 --steps 2 --warmup 3 --no-e2e --no-cpu --no-weak-base"
timeout 300 $CMD > gpurun_out/r2e_plain.json 2> gpurun_out/r2e_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pair_kernel|inv_xt_tile2" -s 6 -c 2 -o gpurun_out/r2e_prof $CMD > gpurun_out/r2e_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2e_ncu.log
