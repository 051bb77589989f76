#!/bin/bash
set -u
OUT=gpurun_out/call14
mkdir -p $OUT
echo "== new GPU tests: model parity + ring geometries" | tee $OUT/summary.txt
timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_engine.py -m gpu -q -k "models or geometr or llama or resnet" > $OUT/pytest_new.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest_new.txt | cut -c1-300 | tee -a $OUT/summary.txt
