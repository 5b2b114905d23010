cd /root/repo
SAVGOL_B200_LIB=/root/repo/$(ls variants_*.so | head -1) timeout 300 python -m pytest tests/test_gpu_2d.py -x -q 2>&1 | tail -1
for v in variants_*.so; do
  SAVGOL_B200_LIB=/root/repo/$v python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"
done
