#!/bin/bash
# GPU call 5: probe K (hand-off with HBM full), ncu captures of the kernels.
set -u
OUT=gpurun_out/call5
mkdir -p $OUT
echo "== probe K" | tee $OUT/summary.txt
timeout 900 ./tools/probe K > $OUT/probe_k.txt 2>&1; echo "probe K rc=$?" | tee -a $OUT/summary.txt
cat $OUT/probe_k.txt | cut -c1-300 | tee -a $OUT/summary.txt
echo "== ncu launch list of the profiling target" | tee -a $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_target_launches.csv python tools/ncu_target.py > $OUT/ncu_target_list.txt 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
grep -c nvs_slab $OUT/ncu_target_launches.csv | tee -a $OUT/summary.txt
echo "== ncu --set full: TMA copy kernel (first eviction launch = D2H, and the D2D launch)" | tee -a $OUT/summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nvs_slab_copy_tma -c 10 -o $OUT/prof_tma python tools/ncu_target.py > $OUT/ncu_full.txt 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nvs_slab_scan|nvs_slab_splat" -c 4 -o $OUT/prof_scan python tools/ncu_target.py > $OUT/ncu_full_scan.txt 2>&1; echo "rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT | tee -a $OUT/summary.txt
