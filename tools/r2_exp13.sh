#!/bin/bash
cd "$(dirname "$0")/.."
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"].get("frac_best_step"), d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"])'
for rep in 1 2 3; do
for W in c2 c5; do
for L in - variants/libsavgol_b200_prev.so; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  echo -n "$W $L: "; timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"
done
done
done
unset SAVGOL_B200_LIB
timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_stream.py tests/test_gpu_1d.py -x -q -m gpu 2>&1 | tail -3
