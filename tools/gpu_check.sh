#!/bin/bash
# One gpurun-able pass: GPU test suite + headline bench line.   usage: gpurun -- 'bash tools/gpu_check.sh'
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 2>&1 | tail -1
