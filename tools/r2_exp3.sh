#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests 2d/sweep"; timeout 900 python -m pytest tests/test_gpu_2d.py tests/test_gpu_random_sweep.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -6
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["fp32"]["frac"], d.get("parity"), d["clocks"], d.get("sustained"))'
echo "== c4/64"; SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
echo "== full c4"; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
timeout 300 python tools/perf_shapes2d.py 2>&1 | head -8
timeout 600 python tools/r2_sweep1d.py 2>&1
SG_C4_IMAGES=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sep_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_add2 python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu --no-sustained > gpurun_out/ncu_c4_add2.log 2>&1
tail -1 gpurun_out/ncu_c4_add2.log | cut -c1-200
echo "== default bench (all configs)"
( time timeout 800 python bench.py > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err ) 2>&1 | tail -3
tail -c 3000 gpurun_out/bench_all.json; tail -5 gpurun_out/bench_all.err
echo "== reference arm"
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
