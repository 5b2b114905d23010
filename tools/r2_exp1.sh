#!/bin/bash
# round-2 experiment 1: additive 2D kernel -- parity + A/B against the generic rank-2 kernel + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
python -m pytest tests/test_gpu_2d.py tests/test_gpu_random_sweep.py -x -q -m gpu 2>&1 | tail -4
for NA in 0 1; do
  echo "== NO_ADDITIVE=$NA"
  SAVGOL_B200_NO_ADDITIVE=$NA SG_C4_IMAGES=64 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('parity'), d['clocks'])"
done
echo "== full c4"
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
python tools/perf_shapes2d.py 2>&1 | head -8
SG_C4_IMAGES=16 ncu --set full --clock-control none --import-source on -k regex:sep_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_add python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4_add.log 2>&1
tail -1 gpurun_out/ncu_c4_add.log | cut -c1-200
