#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c5_build.log 2>&1
timeout 300 python -m pytest tests -m gpu -q -x --tb=short -k "kron_of_matrices or matrix_diagonal or leaf_pullbacks" 2>&1 | tail -3
for big in 0 1; do PO_DENSE_BIG=$big timeout 300 python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch, parops
from parops import Float64, Float32
g = torch.Generator(device="cuda").manual_seed(0)
Ms = [parops.ParMatrix(Float64, 256, 256) for _ in range(3)]
D = parops.ParDiagonal(Float64, 256 ** 3)
K5 = parops.compose(D, parops.ParKron(*Ms))
th = parops.gpu(parops.init(K5, rng=2)); K5t = K5(th)
x = torch.randn(256 ** 3, generator=g, device="cuda", dtype=torch.float64)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: K5t * x)
fl = 3 * 2 * 256 * 256 ** 3
print("PO_DENSE_BIG=%s  C5 matvec %.3f ms  %.2f TFLOP/s (%.0f%% of the measured 34.0 TFLOP/s DFMA peak)" % (os.environ.get("PO_DENSE_BIG"), ms, fl / ms / 1e9, 100 * fl / ms / 1e9 / 34.0))
PY
done
