#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/s11_pytest.log | cut -c1-200
timeout 300 python bench.py --config mutag > $OUT/s11_mutag.json 2> $OUT/s11_mutag.err; cat $OUT/s11_mutag.json; tail -3 $OUT/s11_mutag.err
timeout 300 python bench.py --config rmat > $OUT/s11_rmat.json 2> $OUT/s11_rmat.err; cat $OUT/s11_rmat.json | cut -c1-900; tail -3 $OUT/s11_rmat.err
timeout 300 python scripts/layer_probe.py "v" > $OUT/s11_probe.jsonl 2>$OUT/s11_probe.err; cat $OUT/s11_probe.jsonl | cut -c1-100
