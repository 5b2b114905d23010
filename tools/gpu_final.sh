#!/bin/bash
# Round-end evidence pass (one GPU): bench lines for every workload + ncu launch list and full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for W in c2 c1 c3 c5 c4; do
  python bench.py --workload $W --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_$W.json
done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
bash tools/gpu_profile.sh c2 sg1d_kernel 4
NOLIST=1 bash tools/gpu_profile.sh c3 sg1d_kernel 4
NOLIST=1 bash tools/gpu_profile.sh c5 sg1d_kernel 7
NOLIST=1 bash tools/gpu_profile.sh c4 sep_kernel 3
python tools/perf_shapes.py > gpurun_out/shapes1d.txt 2>&1
python tools/perf_shapes2d.py > gpurun_out/shapes2d.txt 2>&1
