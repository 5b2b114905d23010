cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_1d.py -x -q 2>&1 | tail -2
python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 4 -c 1 -f -o gpurun_out/prof_c2_r1a python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log | cut -c1-300
