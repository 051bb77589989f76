#!/bin/bash
# round 2, GPU call 4 (1 GPU): bursts that follow the releaser's progress words; launch list of a bench run
# whose steps hold few kernels (matmul clients), so that every kernel's share of a step can be read off it
O=gpurun_out/r2c4; mkdir -p $O
timeout 600 python bench.py --steps 6 --warmup 5 --no-extras --keep $O/ours > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"; python tools/brief.py $O/bench_ours.json
timeout 500 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/bench_matmul_launches.csv \
    python bench.py --kind matmul --hbm-fraction 0.1 --tq 4 --steps 2 --warmup 2 --no-extras > $O/ncu_bench.json 2> $O/ncu_bench.err; echo "ncu bench rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c4/bench_matmul_launches.csv")) if len(r) > 14 and r[0].isdigit()]
t = collections.Counter(); n = collections.Counter()
for r in rows:
    k = r[4].split("(")[0][:70]; t[k] += float(r[14]); n[k] += 1
tot = sum(t.values())
for k, v in t.most_common(12):
    print(f"{v/1e6:10.2f} ms {100*v/tot:5.1f}% {n[k]:6d}  {k}")
PY
rm -rf $O/*/main/sock; du -sh $O
