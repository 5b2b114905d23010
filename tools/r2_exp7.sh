#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"])'
{
echo "C4 (64 images of 4096^2, 15x15 order 3 constant), 20 steps, device resident; columns: ms/step, Gpixel/s, fraction of the HBM roofline, parity, SM MHz"
for rep in 1 2; do
for L in - variants/libsavgol_b200_tma2d.so variants/libsavgol_b200_d34.so; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  echo -n "$L: "; SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"
done
done
} 2>&1 | tee gpurun_out/r2_c4_tma_experiment.txt
unset SAVGOL_B200_LIB
echo "== remaining gpu tests"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_host_multi.py tests/test_gpu_multi.py tests/test_gpu_random_sweep.py tests/test_gpu_stream.py tests/test_gpu_threads.py tests/test_gpu_tma.py tests/test_gpu_2d.py -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/r2_profile.sh
