"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (development tool;
the committed copies live under profiles/)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def short(n):
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"^void ", "", n)
    n = n.replace("qr::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<.*)?", n)
    return m.group(1) if m else n


names = [r["Kernel Name"] for r in rows]
first = next((i for i, n in enumerate(names) if "rank_kernel" in n), 0)
tr = rows[first:]
agg = collections.OrderedDict()
for r in tr:
    k = short(r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(u, 1.0)
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
ntrees = max(1, sum(1 for r in tr if "lambda_kernel" in r["Kernel Name"]))
tot = sum(a[1] for a in agg.values())
print("launches after init: %d, boosting iterations: %d" % (len(tr), ntrees))
print("%-30s %10s %12s %10s %10s %7s" % ("kernel", "n/tree", "us/tree", "avg us", "max us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-30s %10.1f %12.1f %10.2f %10.1f %6.1f%%" % (k, a[0] / ntrees, a[1] / ntrees / 1e3, a[1] / a[0] / 1e3,
                                                          a[2] / 1e3, 100 * a[1] / tot))
print("total kernel time per tree: %.1f us" % (tot / ntrees / 1e3))
