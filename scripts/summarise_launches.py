"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a markdown table
(per-kernel launches, total time, share).  Usage: python scripts/summarise_launches.py in.csv out.md "title" """
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
    ns = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault((name, r["Grid Size"], r["Block Size"]), [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
out = ["# " + title, "", "ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and "
       "serialised: compare SHARES, not absolutes.", "", "| kernel | grid | block | launches | total us | share |", "|---|---|---|---|---|---|"]
for (n, g, b), (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %s | %s | %d | %.1f | %.1f%% |" % (n, g, b, c, ns / 1e3, 100 * ns / tot))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:18]))
