#!/bin/bash
set -u
OUT=gpurun_out/call15
mkdir -p $OUT
echo "== model parity tests" | tee $OUT/summary.txt
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q > $OUT/pytest_models.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -30 $OUT/pytest_models.txt | cut -c1-300 | tee -a $OUT/summary.txt
echo "== config #3 shape: matmul(ones, ones), ours then reference, 10% of HBM, TQ 5" | tee -a $OUT/summary.txt
timeout 900 python bench.py --kind matmul --pattern ones --hbm-fraction 0.1 --tq 5 --steps 3 --warmup 3 > $OUT/matmul_ours.json 2> $OUT/matmul_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/matmul_ours.json | cut -c1-1500 | tee -a $OUT/summary.txt
timeout 900 python bench.py --impl reference --kind matmul --pattern ones --hbm-fraction 0.1 --tq 5 --steps 3 --warmup 3 > $OUT/matmul_ref.json 2> $OUT/matmul_ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/matmul_ref.json | cut -c1-1500 | tee -a $OUT/summary.txt
