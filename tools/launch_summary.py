"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share.  Usage: launch_summary.py raw.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    a = agg.setdefault(r[ik][:100], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
tot = sum(t for _, t in agg.values())
print("# total %.1f ms over %d launches (cold-cache, serialised under the profiler: compare SHARES)" % (tot / 1e6, sum(n for n, _ in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-102s n=%4d total=%9.3f ms avg=%9.1f us share=%5.1f%%" % (k, n, t / 1e6, t / n / 1e3, 100 * t / tot))
