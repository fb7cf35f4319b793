set -x
timeout 900 python -m pytest tests/test_heads_gpu.py -x -q -s 2>&1 | grep -E "passed|failed|pca tf32|precision|Error|assert" | head
