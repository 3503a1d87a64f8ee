#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c1_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -3
timeout 300 python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch, parops
from parops import Float32
eng = parops.default_engine()
g = torch.Generator(device="cuda").manual_seed(0)
S = parops.spectral_convolution(Float32, [64, 64], [[(1, 12), (53, 64)], [(1, 12)]], 20)
th = parops.gpu(parops.init(S, rng=2)); St = S(th)
for b in (1, 16):
    x = torch.randn((b, S.Domain), generator=g, device="cuda", dtype=torch.float32).t()
    for _ in range(5): St * x
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): St * x
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    plan, _ = parops.get_plan(St, b, eng)
    plan.set_profiling(True)
    for _ in range(5): St * x
    rows = plan.step_profile(); plan.set_profiling(False)
    print("C1 b=%d apply %.4f ms; kernels: %s" % (b, ms, ", ".join("%.1f us" % (1e3 * r[1] / max(1, r[2])) for r in rows)))
PY
