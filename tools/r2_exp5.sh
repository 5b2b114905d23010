#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu test suite"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["fp32"]["frac"], d.get("parity",{}).get("ok"), d["clocks"], d.get("sustained"))'
echo "== c4/64"; SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
echo "== full c4"; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
echo "== wrappers"; timeout 600 python tools/r2_wrappers.py 2>&1 | tail -12
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
