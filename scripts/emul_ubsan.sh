#!/bin/bash
# UndefinedBehaviorSanitizer run of the kernel emulator: misaligned 128-bit vector accesses (float4 / double2 are alignas(16) in the shim, as on
# the GPU, where a misaligned LDG.128 faults), signed overflow in index arithmetic, out-of-range shifts.  usage: bash scripts/emul_ubsan.sh [-k expr]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/libparops_emul_ubsan.so
g++ -std=c++20 -O1 -g -fsanitize=undefined -fno-sanitize-recover=undefined -fPIC -shared -DPO_EMUL -I$ROOT/tests/emul -x c++ \
    $ROOT/parametricoperators.jl_b200/csrc/po_api.cu -o $OUT -lpthread
cd $ROOT
PO_EMUL_LIB=$OUT LD_PRELOAD=$(gcc -print-file-name=libubsan.so) UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_emul_parity.py tests/test_emul_pullback.py -x -q "$@"
