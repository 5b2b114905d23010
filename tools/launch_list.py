#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log of a bench run into the committed launch list:
profiles/<prefix>.csv (one line per library kernel launch) and profiles/<prefix>_summary.txt (per kernel totals and shares).
usage: tools/launch_list.py gpurun_out/r2_launches.csv profiles/r2_bench_launches"""
import csv, io, re, sys
src, prefix = sys.argv[1], sys.argv[2]
text = open(src, errors="replace").read()
start = text.index('"ID"')
rows = list(csv.reader(io.StringIO(text[start:])))
hdr = rows[0]
ik, ig, ib, im, iv, iu = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value", "Metric Unit"))
ours = re.compile(r"sg1d|sg2d|sep_kernel|flush_kernel|state_append|scatter_kernel|add_rows|compat_inplace|direct_kernel")
launches, total = [], 0
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    total += 1
    name = re.sub(r"^(void )?(sg::|sg2d::\(anonymous namespace\)::|sg2d::|\(anonymous namespace\)::)*", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    if not ours.search(name):
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1e3 if r[iu] in ("nsecond", "ns") else v * 1e3 if r[iu] in ("msecond", "ms") else v
    launches.append((name, us, r[ig], r[ib]))
with open(prefix + ".csv", "w") as f:
    f.write("kernel,duration_us,grid,block\n")
    for n, us, g, b in launches:
        f.write(f'"{n}",{us:.2f},"{g}","{b}"\n')
agg = {}
for n, us, g, b in launches:
    a = agg.setdefault(n, [0, 0.0, g, b])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open(prefix + "_summary.txt", "w") as f:
    f.write("Launch list of the DEFAULT bench command (all configs), `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained` (tools/r2_profile_final.sh, condensed by tools/launch_list.py).\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.  Library kernels only (torch's RNG / fill / copy kernels that build the synthetic inputs are left out);\n")
    f.write(f"captured launches: {total} of which library kernels {len(launches)}.\n\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s}  share of library time  grid x block\n")
    for n, a in agg.items():
        f.write(f"{n:60s} {a[0]:8d} {a[1]:12.1f} {a[1] / a[0]:10.1f} {100 * a[1] / tot:21.1f}%  {a[2]} x {a[3]}\n")
print(open(prefix + "_summary.txt").read())
