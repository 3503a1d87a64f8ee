#!/bin/bash
# compute-sanitizer passes over the new kernels' parity tests (small shapes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/san_build.log 2>&1
K='zmz_fused_step or pruned_real16_kron or full_fft_one_pass or fno_block_fused_epilogue or test_fused_tx_fast_path'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line -k "$K" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/san_memcheck.log | head -20
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=line -k "zmz_fused_step or fno_block_fused_epilogue or test_fused_tx_fast_path" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/san_racecheck.log | head -20
