#!/bin/bash
# round-2 multi-GPU pass: usage  gpurun --gpus N -- 'bash tools/r2_mg.sh N'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_multi.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -5
echo "== c_multi_gpu on $N devices"; timeout 300 oracle/_ref/c_multi_gpu_b200 $N 2>&1 | tail -6
echo "== peer ring check"; timeout 300 $TR --master-port 29510 tools/mg_p2p_check.py 2>&1 | tail -3
echo "== default bench at N=$N"
( time timeout 850 $TR --master-port 29511 bench.py --gpus $N > gpurun_out/bench_all_n$N.json 2> gpurun_out/bench_all_n$N.err ) 2>&1 | tail -3
tail -1 gpurun_out/bench_all_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
def show(k,r): print(k, r['value'], r['ms_per_step'], r['roofline']['frac'], 'parity', r['parity'].get('ok'), r['parity'].get('ranks_ok'), 'e2e', r.get('e2e',{}).get('value'), r.get('e2e',{}).get('h2d_gbs_per_rank'), r.get('config',{}).get('halo','')[:40], 'sust', r.get('sustained',{}).get('frac_hbm') if r.get('sustained') else None)
show('c2', d)
for k,v in d['configs'].items(): show(k,v)
print('numa', d.get('numa_node'))
"
tail -3 gpurun_out/bench_all_n$N.err
echo "== reference arm under torchrun"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
