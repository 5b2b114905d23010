#!/bin/bash
# full validation of HEAD on one B200: GPU test suite, smoke, sanitizer, default bench + reference arm, shape sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu test suite"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
echo "== default bench"
( time timeout 800 python bench.py > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err ) 2>&1 | tail -3
tail -1 gpurun_out/bench_all.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
def show(k,r): print(k, r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline'].get('frac_best_step'), r['roofline']['fp32']['frac'], 'parity', r['parity'].get('ok'), 'e2e', r.get('e2e',{}).get('value'), 'sust', (r.get('sustained') or {}).get('frac_hbm'), 'cpu', (r.get('cpu_baseline') or {}).get('value'), r['clocks']['reasons'])
show('c2', d)
for k,v in d['configs'].items(): show(k,v)
"
tail -3 gpurun_out/bench_all.err
echo "== reference arm"; ( time timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/bench_ref.json
timeout 300 python tools/perf_shapes.py > gpurun_out/r2_shapes_1d.txt 2>&1; sed -n 1,5p gpurun_out/r2_shapes_1d.txt
timeout 600 python tools/r2_wrappers.py > gpurun_out/r2_wrappers.txt 2>&1; sed -n 1,4p gpurun_out/r2_wrappers.txt
timeout 600 python tools/perf_shapes2d.py > gpurun_out/r2_shapes_2d.txt 2>&1; tail -2 gpurun_out/r2_shapes_2d.txt
