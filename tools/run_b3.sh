cd /root/repo
python -m pytest tests/test_gpu_stream.py tests/test_gpu_1d.py -x -q 2>&1 | tail -3
for w in c2 c3 c5; do python bench.py --workload $w --steps 20 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"; done
