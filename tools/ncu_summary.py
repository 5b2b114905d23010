#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): headline metrics, stall mix, executed instructions per code region.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [warp_tiles]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
wt = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'smsp__warps_eligible.avg.per_cycle_active']
for i, h in enumerate(hdr):
    if h in want or ('smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.2):
        print(f"{h:85s} {units[i]:10s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ia]) for r in data)
print("total warp instructions", tot, "static", len(data))
if wt:
    print("per warp-tile", tot / wt)
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and data[j + 1][ia] == data[i][ia]: j += 1
        n = j - i + 1; ex = int(data[i][ia]) / wt; st = sum(int(data[k][ist]) for k in range(i, j + 1))
        if ex * n > 4 or st > 300:
            ops = {}
            for k in range(i, j + 1):
                t = data[k][isrc].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op] = ops.get(op, 0) + 1
            print(f"[{i:5d}-{j:5d}] n={n:4d} x{ex:6.3f} dyn={n*ex:7.1f} stall={st:6d} {sorted(ops.items(), key=lambda x: -x[1])[:5]}")
        i = j + 1
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
t = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in cols}
print("stall samples:", sorted(t.items(), key=lambda x: -x[1])[:8])
