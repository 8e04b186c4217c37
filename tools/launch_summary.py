"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and the
last launches in order.  usage: python tools/launch_summary.py gpurun_out/<file>.csv [tail]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 16
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
agg, seq = collections.OrderedDict(), []
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row["Metric Unit"]
    t = t / 1e3 if unit == "ns" else t * 1e3 if unit == "ms" else t
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    seq.append((name, t, row.get("Grid Size"), row.get("Block Size")))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(v[1] for v in agg.values())
print("%-66s %5s %12s %10s %6s" % ("kernel", "n", "total us", "avg us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-66s %5d %12.1f %10.2f %5.1f%%" % (k, n, t, t / n, 100 * t / total))
print("--- last %d launches" % tail)
for s in seq[-tail:]:
    print("%-66s %10.2f us  grid %s block %s" % s)
