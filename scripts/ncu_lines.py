#!/usr/bin/env python3
"""Summarise an ncu report of k_fused: headline metrics + executed instructions aggregated by source line.
usage: scripts/ncu_lines.py gpurun_out/X.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'launch__shared_mem_per_block_dynamic', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print("%-70s %s %s" % (w, vals[i], units[i]))
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.15:
        print("%-70s %s" % (h.replace('smsp__average_warps_issue_stalled_', 'stall: ').replace('_per_issue_active.ratio', ''), vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''): continue
    try: n = int(r[7]); s = int(r[6])
    except ValueError: continue
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:100]), [0, 0]); a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print("total warp-instructions (by line)", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / ts, k[0], k[1], k[2]))
