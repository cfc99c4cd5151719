#!/usr/bin/env python3
"""Hot code footprint of a kernel: static SASS instructions executed at least `thr` times per tile, by source line.
(The L1.5 instruction cache is 32 KB = 2048 instructions; warps that run asynchronously keep the whole hot path live.)
usage: scripts/ncu_hotcode.py X.ncu-rep [thr_per_tile] [tiles] [min_static]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5; tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 312500.0
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; key = None; agg = collections.Counter(); tot = 0; hot = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8: continue
    if r[0].isdigit(): key = (cur, int(r[0]), r[1].strip()[:100]); continue
    if r[0] == '' and key is not None:
        try: n = int(r[7])
        except ValueError: continue
        tot += 1
        if n >= thr * tiles: hot += 1; agg[key] += 1
print("static instructions %d, hot (>= %.1f per tile) %d = %.1f KB" % (tot, thr, hot, hot * 16 / 1024.0))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v >= mins: print("%5d  %-14s %4d  %s" % (v, k[0][:14], k[1], k[2]))
