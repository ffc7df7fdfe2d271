"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and optionally grid)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        out.append((name, row.get("Grid Size"), row.get("Block Size"), v))
    return out


def main():
    rows = load(sys.argv[1])
    base = load(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, b, v in rows:
        if n.startswith("at::"):
            n = "torch:" + n[4:60]
        agg[n][0] += 1
        agg[n][1] += v
    if base is not None:
        for n, g, b, v in base:
            if n.startswith("at::"):
                n = "torch:" + n[4:60]
            agg[n][0] -= 1
            agg[n][1] -= v
    tot = sum(v[1] for v in agg.values())
    print("total %.1f us over %d launches%s" % (tot, sum(v[0] for v in agg.values()), " (difference)" if base else ""))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        if v[0]:
            print("%10.1f us %5.1f%% %6d  avg %8.1f  %s" % (v[1], 100 * v[1] / tot, v[0], v[1] / max(v[0], 1), k[:90]))


main()
