cd /root/repo
timeout 300 python -m pytest tests/test_gpu_2d.py -x -q 2>&1 | tail -3
for cfg in "64 4096" "1 4096"; do set -- $cfg
SG_C4_IMAGES=$1 SG_C4_SIZE=$2 timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tile  $1 x $2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'])"
done
SAVGOL_B200_2D=stream SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stream 64 x 4096', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"
