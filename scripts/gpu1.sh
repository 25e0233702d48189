set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40
