#!/bin/bash
# Final confirmation of the committed state, invoked the way the driver does.
set -u
OUT=gpurun_out/call10
mkdir -p $OUT
echo "== build check (prebuilt artefacts) + smoke" | tee $OUT/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -2 $OUT/smoke.txt | tee -a $OUT/summary.txt
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.txt | tee -a $OUT/summary.txt
echo "== bench.py --impl reference (driver order: reference first)" | tee -a $OUT/summary.txt
timeout 1500 python bench.py --impl reference --gpus 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_ref.json | tee -a $OUT/summary.txt
echo "== bench.py (ours)" | tee -a $OUT/summary.txt
timeout 1500 python bench.py --gpus 1 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_ours.json | tee -a $OUT/summary.txt
echo "== ncu launch list of the bench command (shares of the step)" | tee -a $OUT/summary.txt
