#!/bin/bash
# First GPU call of round 1: box facts, the design probe, and a sanity run of the
# UNMODIFIED reference (oracle/_ref) with torch on this box at small scale.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/first_call
mkdir -p $OUT
{
  echo "== nproc: $(nproc)"; free -g; 
  echo "== cgroup mem:"; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null
  echo "== ulimit -l: $(ulimit -l)"; 
  echo "== /dev/shm:"; df -h /dev/shm | tail -1
  echo "== THP: $(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)"
  nvidia-smi; nvidia-smi topo -m
  nvidia-smi -q | grep -A12 "GPU Link Info" | head -40
  nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv
  ls -la /dev/nvidia* 2>/dev/null
  lscpu | head -25
  numactl -H 2>/dev/null | head -20
  ls /var/run/nvshare 2>/dev/null; id
} > $OUT/box.txt 2>&1

timeout 900 ./tools/probe ABCDE 8 > $OUT/probe.txt 2>&1
echo "probe rc=$?" >> $OUT/probe.txt

# --- reference sanity at small scale: 2 clients, n=14000 (pytorch-add-small), TQ=2s
export NVSHARE_DEBUG=1
( ./oracle/_ref/nvshare-scheduler > $OUT/ref_sched.log 2>&1 & echo $! > $OUT/sched.pid )
sleep 1
./oracle/_ref/nvsharectl -T 2 >> $OUT/ref_sched.log 2>&1
PIDS=""
for c in 1 2; do
  ( LD_PRELOAD=$PWD/oracle/_ref/libnvshare.so timeout 300 python -m nvshare_b200.workloads --kind add --n 14000 --iters 100000 --seconds 12 \
      --pattern pos --log $OUT/ref_small_c$c.jsonl --tag ref$c > $OUT/ref_small_c$c.out 2>&1 ; echo "rc=$?" >> $OUT/ref_small_c$c.out ) &
  PIDS="$PIDS $!"
done
wait $PIDS
kill $(cat $OUT/sched.pid) 2>/dev/null
unset NVSHARE_DEBUG
tail -3 $OUT/ref_small_c1.out $OUT/ref_small_c2.out
grep -c "DROP_LOCK" $OUT/ref_sched.log
tail -60 $OUT/probe.txt
