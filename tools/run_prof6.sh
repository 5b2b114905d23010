cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 5 -c 1 -f -o gpurun_out/prof_c5_r1f python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c5.log 2>&1
