"""Generates the committed golden fixtures (tests/golden/*.npz) from the oracle.

The reference itself cannot run in the build container (Julia/FFTW/MPI absent), so these are
ORACLE outputs frozen at the time the oracle was pinned against its known-answer tests
(tests/test_oracle.py); they guard against silent drift of the oracle and are what the GPU
suite compares against besides the live oracle.  Run: python -m tests.golden.make_golden"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import parops_oracle as O  # noqa: E402

CASES = {
    "fno2d_f32": dict(kind="fno", T="float32", shape=[16, 16], modes=[[(1, 4), (13, 16)], [(1, 5)]], c=4, batch=2),
    "fno2d_f64": dict(kind="fno", T="float64", shape=[16, 16], modes=[[(1, 4), (13, 16)], [(1, 5)]], c=4, batch=2),
    "kron_c2_f32": dict(kind="c2", T="float32", n=16, keep=4, batch=3),
}


def build(case):
    T = np.dtype(case["T"])
    if case["kind"] == "fno":
        S, parts = O.spectral_convolution(T, case["shape"], case["modes"], case["c"])
        return S, [parts["elementwise_weighting"], parts["mix_channels"]]
    CT = O.complex_of(T)
    n = case["n"]
    RF = O.ParCompose(O.ParRestriction(CT, n // 2 + 1, [(1, case["keep"])]), O.ParDFT(T, n))
    return O.ParKron(RF, O.ParDFT(CT, n), O.ParDFT(CT, n)), []


def run_case(case, g):
    A, plist = build(case)
    theta = {p: g["theta%d" % i] for i, p in enumerate(plist)}
    return A(np.asfortranarray(g["x"]), theta)


def main():
    for name, case in CASES.items():
        A, plist = build(case)
        theta = O.init(A, seed=2)
        x = np.random.default_rng(0).standard_normal((A.Domain, case["batch"])).astype(case["T"])
        y = A(np.asfortranarray(x), theta)
        out = {"x": x, "y": y}
        for i, p in enumerate(plist):
            out["theta%d" % i] = theta[p]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, x.shape, y.shape, y.dtype)


if __name__ == "__main__":
    main()
