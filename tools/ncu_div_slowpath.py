#!/usr/bin/env python3
"""How often does a kernel leave the IEEE division sequence through its out-of-line slow path?
  python tools/ncu_div_slowpath.py <report.ncu-rep> <kernel-regex>
Reads the SASS page of the first matching launch; a division slow path is a local subroutine (ends in RET.REL.NODEC) that tests its
operands against +INF (FSETP ... +INF) and extracts exponents (SHF.R.U32.HI ..., 0x17); prints calls per warp, the lanes that
took it, the share of the kernel's executed instructions, and where the calls come from (needs --import-source / -lineinfo for
nothing: works on SASS alone)."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one block per launch: a "Kernel Name" row, a header row, the instructions
blocks = []
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        blocks.append([r[1] if len(r) > 1 else kern, i + 1, len(rows)])
        if len(blocks) > 1:
            blocks[-2][2] = i
seen = set()
for name, h, end in blocks:
    if name in seen or h >= len(rows):
        continue
    seen.add(name)
    print(name)
    hdr = rows[h]
    isrc, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = []
    for r in rows[h + 1:end]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        data.append((r[isrc].strip(), int(r[ie]), int(r[it])))
    if not data:
        continue
    tot = sum(d[1] for d in data)
    warps = data[0][1]
    start = 0
    subs = []
    for i, d in enumerate(data):  # split into subroutines at RET / EXIT
        if d[0].startswith("RET.REL") or d[0].startswith("EXIT") or " EXIT" in d[0]:
            subs.append((start, i))
            start = i + 1
    found = 0
    for a, b in subs:
        body = data[a:b + 1]
        if not body[-1][0].startswith("RET.REL"):
            continue
        text = " ".join(x[0] for x in body)
        if "+INF" in text and "0x17" in text and "MUFU.RCP" in text and len(body) < 120:
            calls = body[0][1]
            ex = sum(x[1] for x in body)
            lanes = body[0][2] / max(1, body[0][1])
            if calls:
                found += 1
                print("  division slow path at SASS %d..%d: %d calls (%.2f per warp; the kernel's first instruction ran %d times), %.1f lanes, %.2f %% of the executed instructions"
                      % (a, b, calls, calls / max(1, warps), warps, lanes, 100.0 * ex / max(1, tot)))
    if not found:
        print("  no division slow path executed")
    print("  kernel: %d static instructions, %.1f M executed" % (len(data), tot / 1e6))
