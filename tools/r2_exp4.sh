#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu test suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["fp32"]["frac"], d.get("parity"), d["clocks"], d.get("sustained"))'
for NT in 0 1; do
echo "== c4/64 NO_TMA2D=$NT"; SAVGOL_B200_NO_TMA2D=$NT SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
done
echo "== full c4"; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
timeout 300 python tools/perf_shapes2d.py 2>&1 | head -8
echo "== c3 / c1 / c5 e2e"
for W in c3 c1 c5; do timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel"], d.get("parity",{}).get("ok"), d.get("e2e"))'; done
SG_C4_IMAGES=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sep_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_add3 python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-sustained > gpurun_out/ncu_c4_add3.log 2>&1
tail -1 gpurun_out/ncu_c4_add3.log | cut -c1-200
