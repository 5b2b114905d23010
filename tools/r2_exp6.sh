#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/r2_c4_error_stats.py 256 2>&1 | tail -4
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"])'
for rep in 1 2; do
for L in - variants/libsavgol_b200_d34.so; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  for NT in 0 1; do
  echo -n "$L NO_TMA2D=$NT: "; SAVGOL_B200_NO_TMA2D=$NT SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"
  done
done
done
unset SAVGOL_B200_LIB
echo "== remaining gpu tests"; timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_host_multi.py tests/test_gpu_multi.py tests/test_gpu_random_sweep.py tests/test_gpu_stream.py tests/test_gpu_threads.py tests/test_gpu_tma.py -x -q -m gpu 2>&1 | tail -6
