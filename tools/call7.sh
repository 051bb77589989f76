#!/bin/bash
# GPU call 7: effect of burst gating.  GPU tests, bench ours full scale (pos, then ones), ours at 0.6 scale.
set -u
OUT=gpurun_out/call7
mkdir -p $OUT
( while true; do echo "$(date +%s) $(cat /sys/fs/cgroup/memory.current 2>/dev/null)"; sleep 2; done ) > $OUT/memwatch.txt 2>&1 &
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
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.txt | tee -a $OUT/summary.txt
echo "== bench ours full scale" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --keep $OUT/full_ours > $OUT/full_ours.json 2> $OUT/full_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ours.json | tee -a $OUT/summary.txt
echo "== bench ours full scale, all-ones data (the reference's literal tensors)" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --pattern ones --keep $OUT/full_ours_ones > $OUT/full_ours_ones.json 2> $OUT/full_ours_ones.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ours_ones.json | tee -a $OUT/summary.txt
echo "== bench ours at 0.6" | tee -a $OUT/summary.txt
run_guarded 1500 python bench.py --hbm-fraction 0.6 --keep $OUT/ours_06 > $OUT/ours_06.json 2> $OUT/ours_06.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/ours_06.json | tee -a $OUT/summary.txt
echo "== bench reference at 0.6, all-ones data" | tee -a $OUT/summary.txt
run_guarded 2400 python bench.py --impl reference --hbm-fraction 0.6 --pattern ones --keep $OUT/ref_06_ones > $OUT/ref_06_ones.json 2> $OUT/ref_06_ones.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/ref_06_ones.json | tee -a $OUT/summary.txt
kill $WATCH 2>/dev/null
awk '{print $2}' $OUT/memwatch.txt | sort -n | tail -1 | xargs echo "peak cgroup memory.current:" | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
