timeout 600 python -m pytest tests/test_cli_bam.py -x -q -m gpu 2>&1 | grep -E "Error|error|assert|failed|passed" | head -20
