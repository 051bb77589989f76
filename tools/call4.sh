#!/bin/bash
# GPU call 4: unmap/release investigation (probe J), shm provisioning (probe I), GPU tests,
# bench: ours full scale, reference at its largest feasible scale, ours at that same scale.
set -u
OUT=gpurun_out/call4
mkdir -p $OUT
( while true; do echo "$(date +%s) $(cat /sys/fs/cgroup/memory.current 2>/dev/null) $(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits 2>/dev/null)"; sleep 2; done ) > $OUT/memwatch.txt 2>&1 &
WATCH=$!
run_guarded() {
  local secs=$1; shift
  setsid timeout $secs "$@" &
  local BP=$!
  ( LIM=$(cat /sys/fs/cgroup/memory.max 2>/dev/null); case "$LIM" in max|"") LIM=0;; esac
    while [ "$LIM" -gt 0 ] && kill -0 $BP 2>/dev/null; do
      CUR=$(cat /sys/fs/cgroup/memory.current)
      if [ $((LIM-CUR)) -lt 8589934592 ]; then echo "WATCHDOG: memory $CUR near limit $LIM: stopping group $BP" >> $OUT/summary.txt; kill -KILL -- -$BP; break; fi
      sleep 1
    done ) &
  local DOG=$!
  wait $BP; local rc=$?
  kill $DOG 2>/dev/null
  return $rc
}
echo "== probe J" | tee $OUT/summary.txt
timeout 600 ./tools/probe J > $OUT/probe_j.txt 2>&1; echo "probe J rc=$?" | tee -a $OUT/summary.txt
cat $OUT/probe_j.txt | cut -c1-300 | tee -a $OUT/summary.txt
echo "== probe I" | tee -a $OUT/summary.txt
timeout 300 ./tools/probe I > $OUT/probe_i.txt 2>&1; echo "probe I rc=$?" | tee -a $OUT/summary.txt
cat $OUT/probe_i.txt | cut -c1-330 | tee -a $OUT/summary.txt
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest_gpu.txt | tee -a $OUT/summary.txt
echo "== bench ours (default = full scale)" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --keep $OUT/full_ours > $OUT/full_ours.json 2> $OUT/full_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ours.json | tee -a $OUT/summary.txt
echo "== bench reference (default = largest feasible scale)" | tee -a $OUT/summary.txt
run_guarded 2400 python bench.py --impl reference --keep $OUT/ref > $OUT/ref.json 2> $OUT/ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/ref.json | tee -a $OUT/summary.txt
FR=$(python -c "import json;print(json.load(open('$OUT/ref.json'))['config']['hbm_fraction_used'])" 2>/dev/null || echo 0.6)
echo "== bench ours at the reference's scale ($FR)" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --hbm-fraction $FR --keep $OUT/ours_same > $OUT/ours_same.json 2> $OUT/ours_same.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/ours_same.json | tee -a $OUT/summary.txt
echo "== bench ours, reference's literal data (all ones) at full scale" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --pattern ones --keep $OUT/full_ours_ones > $OUT/full_ours_ones.json 2> $OUT/full_ours_ones.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ours_ones.json | tee -a $OUT/summary.txt
kill $WATCH 2>/dev/null
awk '{print $2}' $OUT/memwatch.txt | sort -n | tail -1 | xargs echo "peak cgroup memory.current:" | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
