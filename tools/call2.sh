#!/bin/bash
# GPU call 2: smoke, GPU tests, probes G/H, small-scale bench of both arms, ncu launch list.
set -u
OUT=gpurun_out/call2
mkdir -p $OUT
echo "== smoke" | tee $OUT/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.txt | tee -a $OUT/summary.txt
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest_gpu.txt | tee -a $OUT/summary.txt
echo "== probe H, G" | tee -a $OUT/summary.txt
timeout 300 ./tools/probe H > $OUT/probe_h.txt 2>&1; echo "probe H rc=$?" | tee -a $OUT/summary.txt
timeout 600 ./tools/probe G > $OUT/probe_g.txt 2>&1; echo "probe G rc=$?" | tee -a $OUT/summary.txt
cat $OUT/probe_h.txt $OUT/probe_g.txt | tee -a $OUT/summary.txt
echo "== bench (25% of HBM, rest ballast), ours then reference" | tee -a $OUT/summary.txt
timeout 900 python bench.py --hbm-fraction 0.25 --tq 4 --steps 4 --warmup 3 --keep $OUT/bench_ours > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "bench ours rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_ours.json | tee -a $OUT/summary.txt
timeout 1200 python bench.py --impl reference --hbm-fraction 0.25 --tq 4 --steps 4 --warmup 3 --keep $OUT/bench_ref > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_ref.json | tee -a $OUT/summary.txt
echo "== ncu launch list (smoke)" | tee -a $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.txt 2>&1; echo "ncu rc=$?" | tee -a $OUT/summary.txt
grep -c nvs_slab $OUT/launches.csv | tee -a $OUT/summary.txt
# keep the merged payload small
find $OUT -name "*.jsonl" -size +2M -delete
du -sh $OUT | tee -a $OUT/summary.txt
