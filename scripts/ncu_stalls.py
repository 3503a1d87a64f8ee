"""Stall-reason totals per execution-count region of an ncu SASS page CSV (first launch only)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] if len(his) > 1 else len(rows)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions = []
for r in rows[hi + 1:end]:
    if len(r) < len(hdr): continue
    try: ins = int(r[idx["Instructions Executed"]] or 0)
    except ValueError: continue
    if not regions or regions[-1][0] != ins: regions.append([ins, 0, collections.Counter()])
    regions[-1][1] += 1
    for h in stall_cols:
        v = int(r[idx[h]] or 0)
        if v: regions[-1][2][h] += v
tot = sum(sum(c.values()) for _, _, c in regions)
for ins, n, c in regions:
    s = sum(c.values())
    if s > 0.02 * tot:
        print("exec %-8d x %-4d samples %5.1f%%: %s" % (ins, n, 100 * s / tot, ", ".join("%s %.0f%%" % (k[6:], 100 * v / s) for k, v in c.most_common(6))))
