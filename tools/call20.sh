#!/bin/bash
set -u
OUT=gpurun_out/call20
mkdir -p $OUT
echo "== the driver's command on the committed tree: python bench.py (full BASELINE scale)" | tee $OUT/summary.txt
timeout 440 python bench.py > $OUT/bench_final.json 2> $OUT/bench_final.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_final.json | cut -c1-3000 | tee -a $OUT/summary.txt
