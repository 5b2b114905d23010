#!/bin/bash
cd "$(dirname "$0")/.."
echo "== 2D tests"; timeout 900 python -m pytest tests/test_gpu_2d.py tests/test_gpu_random_sweep.py -x -q -m gpu 2>&1 | tail -4
for NA in 0 1; do echo "== shapes NO_ADDITIVE=$NA"; SAVGOL_B200_NO_ADDITIVE=$NA timeout 300 python tools/perf_shapes2d.py 2>&1 | grep -E "19x19|25x25|15x15 order 3 images 16 x 4096x4096|33x33 order 3|17x17 order 3|21x21 order 3"; done
