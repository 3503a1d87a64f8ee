#!/bin/bash
# Run the CPU kernel-emulator parity tests with the kernels compiled under AddressSanitizer: "device" memory is
# host memory there, so every out-of-bounds global access of a kernel is reported (compute-sanitizer is not
# available on the GPU pool).  usage: bash scripts/emul_asan.sh [-k expr]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/libparops_emul_asan.so
g++ -std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -shared -DPO_EMUL -I$ROOT/tests/emul -x c++ \
    $ROOT/parametricoperators.jl_b200/csrc/po_api.cu -o $OUT -lpthread
cd $ROOT
PO_EMUL_LIB=$OUT LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
    python -m pytest tests/test_emul_parity.py tests/test_emul_pullback.py -x -q "$@"
