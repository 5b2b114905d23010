cd /root/repo
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29510 tools/mg_p2p_check.py 2>&1 | tail -1
timeout 200 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']['ok'])"
for H in p2p nccl; do
SG_C3_HALO=$H timeout 200 $TR --master-port 29512 bench.py --gpus $N --workload c3 --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $H', d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity']['ok'])"
done
