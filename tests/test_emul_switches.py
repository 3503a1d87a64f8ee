"""The kernel-selection switches ("A/B switches") are read once per process, so the default suite
only ever exercises the newest kernel generation for a shape.  These tests re-run a few fused-path parity + reverse-mode
cases on the CPU kernel emulator in a subprocess per setting, so that the fallback generations and the A/B partners
(generation-2 / generation-1 kernels, shared-memory x-phase combine, register-staged inverse, unfused z-mix-z, plain
launches, z-chunked inverse pair) stay correct."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SETTINGS = [
    ({"PO_FUSED_GEN": "2"}, "fused gen-2", "gen-3"),                 # pipelined generation-2 kernels only
    ({"PO_FUSED_GEN": "1"}, None, "gen-2"),                          # first-generation kernels only
    ({"PO_RING_XSHFL": "0", "PO_INV_STAGE": "0"}, "fused gen-2 inv", "staged inv"),  # tile combine, register-staged inverse
    ({"PO_INV_STAGE_THREADS": "512", "PO_INV_CHUNKS": "2"}, "inverse pair in 2 z-chunks", None),
    ({"PO_ZMZ": "0", "PO_L2_HINTS": "0", "PO_PDL": "0"}, "ring fwd", None),
]


@pytest.mark.parametrize("env,must_have,must_not_have", SETTINGS)
def test_fused_paths_under_switch(env, must_have, must_not_have):
    e = dict(os.environ)
    e.update(env)
    e["PO_TRACE"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "run_switch_case.py")], env=e, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-2000:], r.stderr[-4000:])
    if must_have:
        assert must_have in r.stderr, r.stderr[-4000:]
    if must_not_have:
        assert must_not_have not in r.stderr, r.stderr[-4000:]
