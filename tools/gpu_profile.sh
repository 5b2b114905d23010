#!/bin/bash
# ncu captures used for profiles/: launch list of the bench command + full set of the dominant kernel.
#   usage: gpurun -- 'bash tools/gpu_profile.sh c2 sg1d_kernel'   (workload, kernel-name regex)
cd "$(dirname "$0")/.."
W=${1:-c2}; K=${2:-sg1d_kernel}; S=${3:-4}
mkdir -p gpurun_out
[ -n "$NOLIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 5 --warmup 3 --no-cpu > gpurun_out/launch_run_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$W \
    python bench.py --workload $W --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$W.log 2>&1
tail -1 gpurun_out/ncu_$W.log | cut -c1-120
