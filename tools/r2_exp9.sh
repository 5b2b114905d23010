#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"].get("frac_best_step"), d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"])'
{
echo "1D TMA kernels: three boxes per segment (body + halo rows; in-tree at the time) vs ONE box per segment (four tensor maps); 20 steps, device resident"
echo "columns: ms/step, Gsamples/s, fraction of the HBM roofline, best single step, parity, SM MHz"
for rep in 1 2 3; do
for W in c2 c5; do
for L in - variants/libsavgol_b200_tma1box.so; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  echo -n "$W $L: "; timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"
done
done
done
} 2>&1 | tee gpurun_out/r2_tma_1box_experiment.txt
export SAVGOL_B200_LIB=$PWD/variants/libsavgol_b200_tma1box.so
echo "== tests on the one-box variant"; timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_1d.py tests/test_gpu_stream.py tests/test_gpu_random_sweep.py -x -q -m gpu 2>&1 | tail -4
