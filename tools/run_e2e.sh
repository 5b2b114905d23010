cd /root/repo
for mib in 64 16 8 4; do
SAVGOL_B200_CHUNK_MIB=$mib python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 8 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk MiB $mib', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
