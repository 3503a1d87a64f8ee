"""Opcode mix + stall samples from `ncu -i rep --page source --csv --kernel-name ...` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
tot = 0
byop, samples = collections.Counter(), collections.Counter()
nblocks = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        ins = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    toks = r[idx["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    byop[op] += ins
    tot += ins
    samples[op] += int(r[idx["# Samples"]] or 0)
print("total warp instr", tot, "per CTA", tot / nblocks)
ts = sum(samples.values())
for op, c in byop.most_common(22):
    print("%-10s %10d %5.1f%%   stall samples %5.1f%%" % (op, c, 100 * c / tot, 100 * samples[op] / max(1, ts)))
