"""Blackwell-native evidence: per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass).

    python tools/sass_histogram.py > profiles/sass_histogram_r2.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, FFMA2 / FADD2 = packed f32x2 (B200_PROFILING.md "What proves a
Blackwell-native kernel")."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "unirestore_b200", "libunirestore_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "ELECT", "SYNCS", "MUFU",
        "FFMA2", "FADD2", "FMNMX3", "HMMA", "REDG", "ATOMG", "ATOMS", "STS", "LDS", "LDG", "STG"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*$", "", name)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        hist[cur][base] += 1
        hist[cur]["__total__"] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            hist[cur]["UTCHMMA.2CTA"] += 1
print("# SASS opcode histogram per kernel of unirestore_b200/libunirestore_b200.so (cuobjdump -sass, sm_100a)")
print("# %-72s %7s  %s" % ("kernel", "instrs", "  ".join(KEYS)))
tot = collections.Counter()
for k, h in hist.items():
    if not k.startswith("void ur::") and not k.startswith("ur::"):
        continue
    cols = ["%s=%d" % (key, h[key]) for key in KEYS if h[key]]
    print("%-74s %7d  %s" % (k.replace("void ", "")[:74], h["__total__"], " ".join(cols)))
    tot.update(h)
print("# library totals: " + " ".join("%s=%d" % (key, tot[key]) for key in KEYS if tot[key]))
