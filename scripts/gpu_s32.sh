#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gat.py -m gpu -q 2>&1 | grep -E "^E  |assert|Error|passed|failed" | cut -c1-260 | head -30
