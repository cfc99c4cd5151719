#!/usr/bin/env python3
"""SASS instructions attributed to one source line, with execution counts per tile: scripts/ncu_sass_of_line.py X.ncu-rep file.cu LINE [tiles]"""
import csv, io, subprocess, sys
rep, fname, line = sys.argv[1], sys.argv[2], int(sys.argv[3]); tiles = float(sys.argv[4]) if len(sys.argv) > 4 else 312500.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; on = False
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8: continue
    if r[0].isdigit(): on = (cur == fname and int(r[0]) == line); continue
    if on and r[0] == '':
        try: n = int(r[7])
        except ValueError: continue
        print("%8.2f  %s" % (n / tiles, r[3].strip()))
