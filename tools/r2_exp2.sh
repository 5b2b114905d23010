#!/bin/bash
# round-2 experiment 2: additive 2D kernel + TMA 1D kernels -- parity, A/B, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
echo "== tests tma"; timeout 600 python -m pytest tests/test_gpu_tma.py -x -q -m gpu 2>&1 | tail -15
echo "== tests 1d/stream"; timeout 900 python -m pytest tests/test_gpu_1d.py tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -6
echo "== tests 2d/sweep"; timeout 900 python -m pytest tests/test_gpu_2d.py tests/test_gpu_random_sweep.py -x -q -m gpu 2>&1 | tail -6
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity"), d["clocks"])'
for W in c2 c5 c3 c1; do
  for NT in 0 1; do
    echo "== $W NO_TMA=$NT"
    SAVGOL_B200_NO_TMA=$NT timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
  done
done
for NA in 0 1; do
  echo "== c4/64 NO_ADDITIVE=$NA"
  SAVGOL_B200_NO_ADDITIVE=$NA SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
done
echo "== full c4"
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
timeout 300 python tools/perf_shapes2d.py 2>&1 | head -8
SG_C4_IMAGES=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sep_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_add python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4_add.log 2>&1
tail -1 gpurun_out/ncu_c4_add.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg1d_tma -s 4 -c 1 -f -o gpurun_out/prof_c2_tma python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2_tma.log 2>&1
tail -1 gpurun_out/ncu_c2_tma.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg1d_tma -s 4 -c 1 -f -o gpurun_out/prof_c5_tma python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c5_tma.log 2>&1
tail -1 gpurun_out/ncu_c5_tma.log | cut -c1-200
