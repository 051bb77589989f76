#!/bin/bash
# GPU call 11: (1) ncu launch list of bench.py itself (small scale so that the serialised run stays short),
# (2) sweep of the fetch burst size at full scale.
set -u
OUT=gpurun_out/call11
mkdir -p $OUT
echo "== ncu launch list of the bench command (clients are child processes)" | tee $OUT/summary.txt
timeout 1200 ncu --target-processes all -k regex:nvs_slab --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/bench_launches.csv python bench.py --hbm-fraction 0.1 --tq 3 --steps 3 --warmup 3 --keep $OUT/ncu_bench > $OUT/ncu_bench.json 2> $OUT/ncu_bench.err; echo "rc=$?" | tee -a $OUT/summary.txt
grep -c nvs_slab $OUT/bench_launches.csv | tee -a $OUT/summary.txt
tail -1 $OUT/ncu_bench.json | cut -c1-600 | tee -a $OUT/summary.txt
for B in 2048 4096 16384; do
  echo "== bench ours full scale, NVSHARE_BURST_MIB=$B" | tee -a $OUT/summary.txt
  NVSHARE_BURST_MIB=$B timeout 1200 python bench.py --keep $OUT/burst_$B > $OUT/burst_$B.json 2> $OUT/burst_$B.err; echo "rc=$?" | tee -a $OUT/summary.txt
  python - <<PY | tee -a $OUT/summary.txt
import json
r=json.loads(open('$OUT/burst_$B.json').read().strip().splitlines()[-1])
print('burst $B:', 'verified', r.get('verified'), 'stall', round(r.get('stall_ms_per_handoff',0)), 'ms iter/s', round(r.get('iter_per_s',0),2), 'e2e', round(r['e2e']['value'],1), 'device', r.get('device',{}).get('evict_GBps'), r.get('device',{}).get('fetch_GBps'), 'wait', r.get('device',{}).get('wait_ms_mean'))
PY
done
find $OUT -name "client*.jsonl" -size +2M -exec truncate -s 2M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
