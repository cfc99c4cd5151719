#!/usr/bin/env python3
"""Executed warp-instructions per tile by source line (k_fused): scripts/ncu_perline.py X.ncu-rep [min_per_tile] [tiles]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0; tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 312500.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
    try: n = int(r[7]); s = int(r[6])
    except ValueError: continue
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:110]), [0, 0]); a[0] += n; a[1] += s
ts = sum(v[1] for v in agg.values()) or 1
tot = 0
for k in sorted(agg):
    tot += agg[k][0] / tiles
    if agg[k][0] / tiles >= thr: print("%-14s %4d %7.1f %5.2f%%smp  %s" % (k[0][:14], k[1], agg[k][0] / tiles, 100.0 * agg[k][1] / ts, k[2]))
print("total per tile", tot)
