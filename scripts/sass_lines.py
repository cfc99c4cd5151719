#!/usr/bin/env python3
"""Static SASS instruction counts per source line of one kernel: nvdisasm -g -c X.cubin > all.sass; scripts/sass_lines.py all.sass <kernel substring> [file lo hi]"""
import re, sys, collections
path, kern = sys.argv[1], sys.argv[2]
flt = (sys.argv[3], int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else None
cur = None; on = False; cnt = collections.Counter()
for l in open(path):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: on = kern in m.group(1); cur = None; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): cnt[cur] += 1
print("total", sum(cnt.values()))
if flt:
    sel = {k: v for k, v in cnt.items() if k and k[0] == flt[0] and flt[1] <= k[1] <= flt[2]}
    print("range", sum(sel.values()))
    for k in sorted(sel): print(k[1], sel[k])
