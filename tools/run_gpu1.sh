set -x
cd /root/repo
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_1d.py -x -q 2>&1 | tail -30
