set -x
timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q 2>&1 | tail -15
