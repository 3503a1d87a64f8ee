#!/bin/bash
# per-kernel DRAM bytes with caches left alone (inter-kernel L2 reuse visible)
TAG=${1:-dram}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k "regex:po_" -s 21 -c 21 --csv --log-file gpurun_out/${TAG}_dram.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_dram.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_dram.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki][:48]),{})[r[mi]]=float(r[vi].replace(",",""))
for (i,k),m in sorted(d.items()):
    print("%3d %-48s %8.1f us rd %8.2f MB wr %8.2f MB"%(i,k,m.get("gpu__time_duration.sum",0)/1e3,m.get("dram__bytes_read.sum",0)/1e6,m.get("dram__bytes_write.sum",0)/1e6))
PY
