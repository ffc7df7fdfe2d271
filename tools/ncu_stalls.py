"""Top stall locations (SASS) of an .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = rows[2:]
ia, isamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", ", ".join("%s %.1f%%" % (k, 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  %-60s %s" % (int(r[isamp]), 100 * int(r[isamp]) / tot, r[ia][:60], st))
