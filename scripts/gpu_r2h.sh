#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/mb_tensor.py > gpurun_out/r2h_plain.log 2>&1 || exit 1
for kb in 18 22 24; do
  PO_TS_CHUNK_KB=$kb timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:po_tensor" --csv --log-file gpurun_out/r2h_tensor_kb$kb.csv python scripts/mb_tensor.py > gpurun_out/r2h_ncu.log 2>&1
  echo "chunk $kb KB:"; grep -v "^==" gpurun_out/r2h_tensor_kb$kb.csv | awk -F'","' '{print substr($5,1,50), $(NF)}' | sort | uniq -c | sort -k1nr | head -8
done
