"""Group the SASS page of an ncu report (ncu -i rep --page source --csv --kernel-name ...) into runs of
equal execution count: shows which loop bodies the dynamic instructions and stall samples come from."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] if len(his) > 1 else len(rows)   # first launch only
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
runs = []
for n, r in enumerate(rows[hi + 1:end]):
    if len(r) < len(hdr):
        continue
    try:
        ins = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    toks = r[idx["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    samp = int(r[idx["# Samples"]] or 0)
    if runs and runs[-1][0] == ins:
        runs[-1][1] += 1
        runs[-1][2][op] += 1
        runs[-1][3] += samp
    else:
        runs.append([ins, 1, collections.Counter({op: 1}), samp, n])
tot = sum(a * b for a, b, _, _, _ in runs)
ts = sum(r[3] for r in runs)
print("total warp instr", tot, "samples", ts)
for ins, cnt, ops, samp, n in runs:
    if ins * cnt > tot * 0.004:
        print("@%-5d exec %-9d x %-4d = %5.1f%%  samples %5.1f%%  %s" % (n, ins, cnt, 100 * ins * cnt / tot, 100 * samp / max(ts, 1), dict(ops.most_common(9))))
