#!/usr/bin/env python3
"""Attribute ncu per-instruction samples of one kernel to source lines / inlined functions.
  python tools/ncu_hotspots.py <report.ncu-rep> <kernel-name> <lib.so> [topN [section-substring]]
kernel-name selects the launches in the report (ncu --kernel-name: base name or regex:...); section-substring selects the ONE
.text section of the cubin whose line table is used (mangled name, e.g. k_shadeILi128ELi1ELi2E; default: the kernel name).
Joins `ncu --page source --csv` (SASS view: address, #samples, instructions executed) with
`nvdisasm --print-line-info` of the cubin extracted from the library (needs -lineinfo)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kernel, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
section = sys.argv[5] if len(sys.argv) > 5 else kernel
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
if not dis:
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# address -> (file:line, inline chain) inside the kernel's section
amap = {}
cur = None
chain = ""
insec = False
prev_annot = False
for line in dis.splitlines():
    if line.startswith("//--------------------- .text."):
        insec = section in line
        continue
    if not insec:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        if not prev_annot:  # a block of annotations lists the innermost location first, the kernel's own line last
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            chain = m.group(3)
        prev_annot = True
        continue
    prev_annot = False
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
    if m and cur:
        amap[int(m.group(1), 16)] = (cur, chain, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
per_line = collections.Counter()
per_line_ex = collections.Counter()
per_line_thr = collections.Counter()
tot = totex = 0
for r in rows[hi + 1:]:
    if len(r) <= ithr or not r[ia]:
        continue
    try:
        a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    off = a - base
    s = int(r[isamp] or 0)
    e = int(r[iex] or 0)
    t = int(r[ithr] or 0)
    key = amap.get(off, (("?", 0), "", ""))[0]
    per_line[key] += s
    per_line_ex[key] += e
    per_line_thr[key] += t
    tot += s
    totex += e
print("kernel %s: %d samples, %d warp instructions" % (kernel, tot, totex))
print("%-28s %8s %6s %12s %6s %6s" % ("file:line", "samples", "%", "warp-inst", "%", "lanes"))
for key, s in per_line.most_common(top):
    e = per_line_ex[key]
    print("%-28s %8d %6.2f %12d %6.2f %6.1f" % ("%s:%d" % key, s, 100.0 * s / max(tot, 1), e, 100.0 * e / max(totex, 1), per_line_thr[key] / max(e, 1)))
# per file totals
pf = collections.Counter()
pfe = collections.Counter()
for key, s in per_line.items():
    pf[key[0]] += s
    pfe[key[0]] += per_line_ex[key]
print("-- per file")
for f, s in pf.most_common():
    print("%-28s %8d %6.2f %12d %6.2f" % (f, s, 100.0 * s / max(tot, 1), pfe[f], 100.0 * pfe[f] / max(totex, 1)))
