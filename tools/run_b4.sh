cd /root/repo
python -m pytest tests/test_gpu_2d.py -x -q 2>&1 | tail -12
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'])"
