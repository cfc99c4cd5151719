#!/usr/bin/env python3
"""Aggregate an ncu report of k_fused by kernel pass (source-line ranges found from the pass marker comments).
usage: scripts/ncu_passes.py gpurun_out/X.ncu-rep"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = open("gargammel_b200/csrc/gg_kernels.cu").read().split("\n")
marks = []
for i, l in enumerate(lines, 1):
    m = re.search(r"// -{20,} (.*)", l)
    if m: marks.append((i, m.group(1)[:60]))
def pass_of(f, ln, txt):
    if f.endswith("gg_kernels.cu"):
        name = "prologue/other"
        for i, n in marks:
            if ln >= i: name = n
        return name
    return None
cur = None; agg = {}; byfile = {}
rows = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
    try: n = int(r[7]); s = int(r[6])
    except ValueError: continue
    rows.append((cur, int(r[0]), r[1].strip(), n, s))
tot = sum(r[3] for r in rows) or 1; ts = sum(r[4] for r in rows) or 1
for f, ln, txt, n, s in rows:
    key = pass_of(f, ln, txt) or ("[" + f + "]")
    a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += s
print("%-62s %8s %8s" % ("pass", "inst %", "samples %"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-62s %7.2f%% %7.2f%%" % (k, 100 * v[0] / tot, 100 * v[1] / ts))
print("-- device-header lines (inlined helpers), top by instructions")
hl = {}
for f, ln, txt, n, s in rows:
    if not f.endswith("gg_kernels.cu"):
        a = hl.setdefault((f, ln, txt[:90]), [0, 0]); a[0] += n; a[1] += s
for k, v in sorted(hl.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / ts, k[0], k[1], k[2]))
