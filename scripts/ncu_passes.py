#!/usr/bin/env python3
"""Aggregate an ncu report of k_fused by kernel pass / helper function.
Kernel body lines are attributed to the pass whose marker comment ("// ---- ... name") precedes them; lines of helper functions
(before the kernel in gg_kernels.cu, and in gg_device.cuh) to the enclosing __device__ function.
usage: scripts/ncu_passes.py gpurun_out/X.ncu-rep"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def index(path):
    marks = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.search(r"// -{4,} (.*)", l)
        if m and "k_fused" not in l: marks.append((i, "pass: " + m.group(1)[:48])); continue
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", l)
        if m: marks.append((i, "fn: " + m.group(1)))
        m = re.match(r"\s*(?:struct)\s+(\w+)\s*\{", l)
        if m: marks.append((i, "fn: " + m.group(1)))
    return marks
idx = {"gg_kernels.cu": index("gargammel_b200/csrc/gg_kernels.cu"), "gg_device.cuh": index("gargammel_b200/csrc/gg_device.cuh")}
def where(f, ln):
    if f in idx:
        name = "top"
        for i, n in idx[f]:
            if ln >= i: name = n
        return name
    return "[" + f + "]"
cur = None; agg = {}
rows = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
    try: n = int(r[7]); s = int(r[6])
    except ValueError: continue
    rows.append((cur, int(r[0]), n, s))
tot = sum(r[2] for r in rows) or 1; ts = sum(r[3] for r in rows) or 1
for f, ln, n, s in rows:
    a = agg.setdefault(where(f, ln), [0, 0]); a[0] += n; a[1] += s
print("%-62s %8s %8s" % ("where", "inst %", "samples %"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 1000 > tot: print("%-62s %7.2f%% %7.2f%%" % (k, 100 * v[0] / tot, 100 * v[1] / ts))
