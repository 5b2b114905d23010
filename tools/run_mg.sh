cd /root/repo
nvidia-smi -L | head -4
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29510 tools/mg_p2p_check.py 2>&1 | tail -3
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']['ok'])"
for H in p2p nccl; do
SG_C3_HALO=$H timeout 300 $TR --master-port 29512 bench.py --gpus $N --workload c3 --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $H', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'], d['config'].get('halo'))"
done
for W in c4 c5; do
timeout 300 $TR --master-port 29514 bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$W', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'])"
done
timeout 300 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
