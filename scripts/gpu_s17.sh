#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/s17_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/s17_pytest.log | cut -c1-250
