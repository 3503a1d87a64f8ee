"""Run a few fused-path parity cases on the CPU kernel emulator in THIS process, under whatever PO_* environment the
caller set (the kernel-selection switches are read once per process, so tests/test_emul_switches.py runs this file in
a subprocess per setting).  Prints the kernels the plan compiler chose (PO_TRACE) and 'ok' at the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "parametricoperators.jl_b200"))
os.environ.setdefault("PO_TRACE", "1")

import tests.conftest  # noqa: E402,F401  (puts the package directory on sys.path the way pytest does)
import parops  # noqa: E402
from tests import parity_cases as pc  # noqa: E402
from tests.emul.build_emul import build  # noqa: E402

eng = parops.Engine(device="cpu", lib=parops._lib.bind(build()))
F32 = pc.F32
for shape, c, batch in [([4, 64, 32], 20, 2), ([4, 32, 64, 16], 2, 1)]:
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    pc.check_case(eng, pc.fno_forward_transform(F32, shape, modes, c), batch=batch)
    pc.check_case(eng, pc.fno_forward_transform(F32, shape, modes, c, adj=True), batch=batch)
    pc.check_case(eng, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    pc.check_pullback(eng, pc.fno_spectral(F32, shape, modes, c), batch=batch)
eng.destroy()
print("ok")
