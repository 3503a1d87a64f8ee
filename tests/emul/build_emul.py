"""Build the CPU kernel emulator of libparops (TEST TOOLING, see emul_cuda.h).

The same csrc/*.cu sources are compiled with g++ -DPO_EMUL so that kernel logic can be
debugged without a GPU.  The product never loads this library."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "parametricoperators.jl_b200", "csrc")
OUT = os.path.join(HERE, "libparops_emul.so")


def build(force=False):
    # PO_EMUL_LIB: use a pre-built emulator library (e.g. an AddressSanitizer build: see scripts/emul_asan.sh)
    if os.environ.get("PO_EMUL_LIB"):
        return os.environ["PO_EMUL_LIB"]
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(HERE, "emul_cuda.h"), os.path.join(ROOT, "include", "parops.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-DPO_EMUL", "-I" + HERE, "-x", "c++",
           "-Wl,-Bsymbolic",  # entry points that call other entry points (po_block_epilogue_pullback -> po_block_epilogue) must stay inside THIS
           # library even when libparops.so (same exported names, RTLD_GLOBAL) was loaded first by another test of the session
           os.path.join(CSRC, "po_api.cu"), "-o", OUT, "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
