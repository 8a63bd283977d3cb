timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "Error|error|assert|failed|passed|raise" | head -20
