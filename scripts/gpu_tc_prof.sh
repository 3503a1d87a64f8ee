#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/tc_build.log 2>&1
cat > /tmp/tc_run.py <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch, parops
from parops import Float32
g = torch.Generator(device="cuda").manual_seed(0)
c, npts = 20, 64 * 64 * 64 * 32
Wc = parops.channel_mixing(Float32, [64, 64, 64, 32], c)
thw = parops.gpu(parops.init(Wc, rng=3)); Wm = list(thw.values())[0]
xb = torch.randn((1, c * npts), generator=g, device="cuda", dtype=torch.float32).t()
sb = torch.randn((1, c * npts), generator=g, device="cuda", dtype=torch.float32).t()
for _ in range(4): parops.block_epilogue(Wm, xb, sb)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/tc_run.py && timeout 600 ncu --set full --clock-control none --import-source on -k "regex:block_epilogue" -s 2 -c 1 -o gpurun_out/r1tc_prof python /tmp/tc_run.py > gpurun_out/r1tc_ncu.log 2>&1
echo "ncu exit $?"
