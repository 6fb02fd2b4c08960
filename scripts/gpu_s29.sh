#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_hostside.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_gat.py -m gpu -q 2>&1 | tail -6 | cut -c1-250
timeout 300 python bench.py --config cora > $OUT/s29_cora.json 2> $OUT/s29_cora.err; echo "cora rc=$?"; tail -2 $OUT/s29_cora.err; python -c "
import json
j=json.loads(open('$OUT/s29_cora.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['value'], j.get('e2e'), j.get('extra') or j.get('config'))" | cut -c1-900
