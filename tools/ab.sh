#!/bin/bash
# A/B a set of library builds on one workload: usage  bash tools/ab.sh <workload> lib1.so lib2.so ...   ("-" = the in-tree build)
cd "$(dirname "$0")/.."
W=$1; shift
for L in "$@"; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  echo -n "$L: "; python bench.py --workload $W --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('parity'))"
done
