#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass on every object of scpp_b200/_obj): the evidence for which data-movement and
FP64 instructions each kernel really contains (DMMA = mma.sync.m8n8k4.f64, UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ...)."""
import collections, glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
want = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "SHFL", "BAR", "ATOMS", "ATOMG", "RED", "LDC", "IMAD", "HMMA", "UTCHMMA"]
rows = []
for obj in sorted(glob.glob(os.path.join(root, "scpp_b200", "_obj", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kern = None; cnt = None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if kern: rows.append((kern, cnt))
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]; cnt = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cnt is not None:
            cnt[m.group(1).split(".")[0]] += 1; cnt["total"] += 1
    if kern: rows.append((kern, cnt))
print("%-64s %8s " % ("kernel", "total") + " ".join("%7s" % w for w in want))
for k, c in rows:
    if c["total"] < 200: continue
    print("%-64s %8d " % (k[:64], c["total"]) + " ".join("%7d" % c[w] for w in want))
