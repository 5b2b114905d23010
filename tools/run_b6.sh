cd /root/repo
timeout 300 python -m pytest tests/test_gpu_2d.py -x -q 2>&1 | tail -2
for cfg in "64 4096" "1 4096"; do set -- $cfg
SG_C4_IMAGES=$1 SG_C4_SIZE=$2 timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 x $2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"
done
