cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > gpurun_out/r2_launch_run.log 2>&1
tail -c 300 gpurun_out/r2_launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sg1d_packed -s 3 -c 1 -f -o gpurun_out/r2_prof_l64 python tools/run_shape.py 64 > gpurun_out/r2_ncu_l64.log 2>&1; tail -1 gpurun_out/r2_ncu_l64.log
