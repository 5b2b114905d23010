cd /root/repo
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -f -o gpurun_out/prof_c4_tile python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log | cut -c1-100
