#!/usr/bin/env python3
"""L1 data-pipe load by source line (k_fused): shared-memory wavefronts (and the excess over ideal) and global tag requests per tile.
usage: scripts/ncu_wavefronts.py X.ncu-rep [min_per_tile] [tiles]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0; tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 312500.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}; col = None
def num(x):
    try: return float(x)
    except ValueError: return 0.0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 20 and r[0] == 'Line No':
        col = {n: i for i, n in enumerate(r)}; continue
    if col is None or len(r) < 24 or not r[0].isdigit(): continue
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:90]), [0.0, 0.0, 0.0, 0.0])
    a[0] += num(r[col['L1 Wavefronts Shared']]); a[1] += num(r[col['L1 Wavefronts Shared Excessive']])
    a[2] += num(r[col['L1 Tag Requests Global']]); a[3] += num(r[col['Instructions Executed']])
tot = [sum(v[i] for v in agg.values()) / tiles for i in range(4)]
print("per tile: shared wavefronts %.1f (excess %.1f), global tag requests %.1f, instructions %.1f" % tuple(tot))
print("%-14s %5s %8s %8s %8s %8s" % ("file", "line", "sh.wave", "excess", "gl.tags", "inst"))
for k in sorted(agg, key=lambda k: -(agg[k][0] + agg[k][2])):
    v = [x / tiles for x in agg[k]]
    if v[0] + v[2] >= thr: print("%-14s %5d %8.1f %8.1f %8.1f %8.1f  %s" % (k[0][:14], k[1], v[0], v[1], v[2], v[3], k[2]))
