#!/bin/bash
# compute-sanitizer memcheck over the tensor-core gradient kernels' parity tests
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck_bwd.log \
    python -m pytest tests/test_gpu_backward_tc.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
echo "rc=$?"; tail -5 $OUT/memcheck_bwd.log
