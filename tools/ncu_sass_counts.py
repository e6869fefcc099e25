#!/usr/bin/env python3
"""Per-SASS-instruction execution counts of one kernel launch from an .ncu-rep (source page).
  python tools/ncu_sass_counts.py <report.ncu-rep> <kernel-regex> [min_Minst]
Prints loads/stores/branches and every instruction executed more than min_Minst million times (warp level),
with the average number of active threads -- the raw data behind the per-phase tables in profiles/."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1:3]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, ie, it, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        if data:
            break
        continue
    data.append(r)
tot = sum(int(r[ie]) for r in data)
print("instructions %d, warp-level executed %.1f M" % (len(data), tot / 1e6))
for k, r in enumerate(data):
    s, e = r[isrc].strip(), int(r[ie])
    if e / 1e6 >= thr or any(t in s for t in ("LDG", "LDL", "STL", "STG", "ATOM", "CALL")) and e > 0:
        print("%5d %-72s %9.1f M  lanes %4.1f  samples %s" % (k, s[:72], e / 1e6, int(r[it]) / max(1, e), r[iS]))
