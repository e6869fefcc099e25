#!/usr/bin/env python3
"""Static SASS size of a kernel attributed to the source FUNCTION each instruction was inlined from.
  python tools/sass_size_by_function.py <lib.so> <kernel-substring> [topN]"""
import bisect
import collections
import os
import re
import subprocess
import sys
import tempfile

lib, kernel = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
funcs = {}


def func_of(path, line):
    if path not in funcs:
        starts, names = [], []
        try:
            for i, l in enumerate(open(path), 1):
                m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:PRB_DEV|__device__|__global__|inline|static)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$", l)
                if m and not l.strip().startswith("//") and m.group(1) not in ("if", "for", "while", "return", "defined"):
                    starts.append(i)
                    names.append(m.group(1))
        except OSError:
            pass
        funcs[path] = (starts, names)
    starts, names = funcs[path]
    k = bisect.bisect_right(starts, line) - 1
    return names[k] if k >= 0 else "?"


cnt = collections.Counter()
insec = False
prev_annot = False
cur = None
total = 0
for line in dis.splitlines():
    if line.startswith("//--------------------- .text."):
        insec = kernel in line
        continue
    if not insec:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        if not prev_annot:
            cur = (m.group(1), int(m.group(2)))
        prev_annot = True
        continue
    prev_annot = False
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line) and cur:
        cnt[(os.path.basename(cur[0]), func_of(cur[0], cur[1]))] += 1
        total += 1
print("kernel %s: %d SASS instructions (%.0f KB)" % (kernel, total, total * 16 / 1024))
for (f, fn), n in cnt.most_common(top):
    print("%-22s %-34s %8d %6.2f%%" % (f, fn, n, 100.0 * n / total))
