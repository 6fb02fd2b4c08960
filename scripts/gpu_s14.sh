#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/s14_pytest.log | cut -c1-200
timeout 300 python bench.py --config rmat > $OUT/s14_rmat.json 2> $OUT/s14_rmat.err; cut -c260-420 $OUT/s14_rmat.json
