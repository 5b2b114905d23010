#!/bin/bash
# round-2 ncu evidence: launch list of the default bench command + full captures of the dominant kernels.
#   usage: gpurun -- 'bash tools/r2_profile.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) every launch of the default (all-config) bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > gpurun_out/r2_launch_run.log 2>&1
tail -c 400 gpurun_out/r2_launch_run.log
# (2) full captures (one launch each, steady state)
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2_prof_$1 \
      python bench.py --workload $2 --steps 3 --warmup 3 --no-e2e --no-cpu --no-sustained > gpurun_out/r2_ncu_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_$1.log | cut -c1-160
}
cap c2 c2 sg1d_tma 4
cap c5 c5 sg1d_tma 2
cap c3 c3 sg1d_kernel 3
SG_C4_IMAGES=256 cap c4 c4 sep_kernel 3
# (3) shapes
timeout 600 python tools/perf_shapes.py > gpurun_out/r2_shapes_1d.txt 2>&1; tail -3 gpurun_out/r2_shapes_1d.txt
timeout 600 python tools/perf_shapes2d.py > gpurun_out/r2_shapes_2d.txt 2>&1; tail -3 gpurun_out/r2_shapes_2d.txt
timeout 600 python tools/r2_sweep1d.py > gpurun_out/r2_sweep1d.txt 2>&1; tail -3 gpurun_out/r2_sweep1d.txt
timeout 600 python tools/r2_wrappers.py > gpurun_out/r2_wrappers.txt 2>&1; tail -3 gpurun_out/r2_wrappers.txt
