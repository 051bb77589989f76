#!/bin/bash
# GPU call 3: GPU tests, shm-pool probe, 25% scale bench both arms, then FULL SCALE both arms with a memory watchdog.
set -u
OUT=gpurun_out/call3
mkdir -p $OUT
( while true; do echo "$(date +%s) $(cat /sys/fs/cgroup/memory.current 2>/dev/null) $(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits 2>/dev/null)"; sleep 2; done ) > $OUT/memwatch.txt 2>&1 &
WATCH=$!
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest_gpu.txt | tee -a $OUT/summary.txt
echo "== probe I" | tee -a $OUT/summary.txt
timeout 300 ./tools/probe I > $OUT/probe_i.txt 2>&1; echo "probe I rc=$?" | tee -a $OUT/summary.txt
cat $OUT/probe_i.txt | cut -c1-400 | tee -a $OUT/summary.txt
echo "== bench 25% ours / reference" | tee -a $OUT/summary.txt
timeout 900 python bench.py --hbm-fraction 0.25 --tq 4 --keep $OUT/b25_ours > $OUT/b25_ours.json 2> $OUT/b25_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/b25_ours.json | tee -a $OUT/summary.txt
timeout 900 python bench.py --impl reference --hbm-fraction 0.25 --tq 4 --keep $OUT/b25_ref > $OUT/b25_ref.json 2> $OUT/b25_ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/b25_ref.json | tee -a $OUT/summary.txt
echo "== FULL SCALE ours" | tee -a $OUT/summary.txt
# watchdog: if the cgroup gets within 12 GiB of its limit, stop the clients instead of being OOM-killed
# each bench runs in its own process group (setsid) so the watchdog can stop exactly that group by id
run_guarded() {  # $1 = seconds, rest = command
  local secs=$1; shift
  setsid timeout $secs "$@" &
  local BP=$!
  ( LIM=$(cat /sys/fs/cgroup/memory.max 2>/dev/null); case "$LIM" in max|"") LIM=0;; esac
    while [ "$LIM" -gt 0 ] && kill -0 $BP 2>/dev/null; do
      CUR=$(cat /sys/fs/cgroup/memory.current)
      if [ $((LIM-CUR)) -lt 12884901888 ]; then echo "WATCHDOG: memory $CUR near limit $LIM: stopping group $BP" >> $OUT/summary.txt; kill -KILL -- -$BP; break; fi
      sleep 1
    done ) &
  local DOG=$!
  wait $BP; local rc=$?
  kill $DOG 2>/dev/null
  return $rc
}
run_guarded 1500 python bench.py --tq 6 --keep $OUT/full_ours > $OUT/full_ours.json 2> $OUT/full_ours.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ours.json | tee -a $OUT/summary.txt
echo "== FULL SCALE reference" | tee -a $OUT/summary.txt
run_guarded 2400 python bench.py --impl reference --tq 6 --keep $OUT/full_ref > $OUT/full_ref.json 2> $OUT/full_ref.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/full_ref.json | tee -a $OUT/summary.txt
kill $WATCH 2>/dev/null
awk '{print $2}' $OUT/memwatch.txt | sort -n | tail -1 | xargs echo "peak cgroup memory.current:" | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
