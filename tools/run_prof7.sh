cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sep_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_r1a python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log | cut -c1-200
