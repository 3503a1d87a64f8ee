#!/bin/bash
# Race detection for the kernels: the CPU kernel emulator runs every CUDA thread of a block as a real std::thread (std::barrier =
# __syncthreads, atomics = atomics, mbarriers = acquire / release counters), so ThreadSanitizer sees a missing barrier or an unordered
# shared-memory hand-over as a data race (checked: removing the barrier between the two FFT stages of po_c2_fft128_rows is reported).
# usage: bash scripts/emul_tsan.sh [-k expr]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/libparops_emul_tsan.so
g++ -std=c++20 -O1 -g -fsanitize=thread -fPIC -shared -DPO_EMUL -I$ROOT/tests/emul -x c++ \
    $ROOT/parametricoperators.jl_b200/csrc/po_api.cu -o $OUT -lpthread
cd $ROOT
PO_EMUL_LIB=$OUT LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
    python -m pytest tests/test_emul_parity.py tests/test_emul_pullback.py -x -q "$@" 2>&1 | tee /tmp/emul_tsan.log | tail -5
echo "ThreadSanitizer data-race reports: $(grep -c 'WARNING: ThreadSanitizer: data race' /tmp/emul_tsan.log || true)"
