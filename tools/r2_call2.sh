#!/bin/bash
# round 2, GPU call 2 (1 GPU): full bench of our arm (with the same-scale sub-run), the reference arm,
# BASELINE config #3 (matmul, TQ 30) for both arms, ncu captures of the scan/hash kernel and a launch list
O=gpurun_out/r2c2; mkdir -p $O
timeout 1200 python bench.py --steps 6 --warmup 5 --keep $O/ours > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"; tail -c 1500 $O/bench_ours.json; tail -3 $O/bench_ours.err
timeout 1200 python bench.py --impl reference --steps 4 --warmup 3 --keep $O/ref > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 1200 $O/bench_ref.json; tail -3 $O/bench_ref.err
timeout 900 python bench.py --kind matmul --pattern pos --steps 3 --warmup 2 --no-extras --keep $O/mm_ours > $O/mm_ours.json 2> $O/mm_ours.err; echo "matmul ours rc=$?"; tail -c 1500 $O/mm_ours.json; tail -3 $O/mm_ours.err
timeout 900 python bench.py --kind matmul --pattern pos --steps 3 --warmup 2 --no-extras --hbm-fraction 0.6 --keep $O/mm_ours06 > $O/mm_ours06.json 2> $O/mm_ours06.err; echo "matmul ours 0.6 rc=$?"; tail -c 600 $O/mm_ours06.json
timeout 1200 python bench.py --kind matmul --pattern pos --impl reference --steps 3 --warmup 2 --keep $O/mm_ref > $O/mm_ref.json 2> $O/mm_ref.err; echo "matmul ref rc=$?"; tail -c 1200 $O/mm_ref.json; tail -3 $O/mm_ref.err
# ncu: never a bench value
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nvs_slab_scan -c 4 -o $O/prof_scan python tools/ncu_target_r2.py > $O/ncu_scan.txt 2>&1; echo "ncu scan rc=$?"
ncu -i $O/prof_scan.ncu-rep --page raw --csv > $O/prof_scan_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/target_launches.csv python tools/ncu_target_r2.py > $O/ncu_list.txt 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/bench_launches.csv \
    python bench.py --hbm-fraction 0.1 --tq 3 --steps 2 --warmup 3 --no-extras > $O/ncu_bench.json 2> $O/ncu_bench.err; echo "ncu bench rc=$?"
python tools/ncu_summarise.py $O/prof_scan_raw.csv > $O/prof_scan_summary.json 2>&1
rm -rf $O/*/main/sock; du -sh $O; ls $O
