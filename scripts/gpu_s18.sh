#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/s18_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/s18_pytest.log | cut -c1-250
for v in "" _nohalf; do KAGNN_LIB=kagnn_b200/lib/libkagnn_b200$v.so timeout 300 python scripts/layer_probe.py "h$v" 2>/dev/null | cut -c1-100; done
