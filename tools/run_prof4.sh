cd /root/repo
mkdir -p gpurun_out
python bench.py --workload c5 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'])"
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 4 -c 1 -f -o gpurun_out/prof_c2_r1d python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 5 -c 1 -f -o gpurun_out/prof_c5_r1d python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-100
