cd /root/repo
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 2>&1 | tail -3
python bench.py --workload c1 --steps 20 --no-e2e --no-cpu 2>&1 | tail -2
python bench.py --workload c3 --steps 10 --no-e2e --no-cpu 2>&1 | tail -2
python bench.py --workload c5 --steps 10 --no-e2e --no-cpu 2>&1 | tail -2
python bench.py --workload c4 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
