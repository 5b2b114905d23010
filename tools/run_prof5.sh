cd /root/repo
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -1 gpurun_out/bench_c2.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 4 -c 1 -f -o gpurun_out/prof_c2_r1e python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 4 -c 1 -f -o gpurun_out/prof_c3_r1e python bench.py --workload c3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg1d_kernel -s 5 -c 1 -f -o gpurun_out/prof_c5_r1e python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out | tail -8
