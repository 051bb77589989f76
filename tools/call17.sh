#!/bin/bash
set -u
OUT=gpurun_out/call17
mkdir -p $OUT
echo "== loading client beside a computing one (SURVEY 8f rank 3)" | tee $OUT/summary.txt
NVSHARE_DEBUG=1 timeout 600 python tools/loader_probe.py > $OUT/loader_probe.jsonl 2> $OUT/loader_probe.err; echo "rc=$?" | tee -a $OUT/summary.txt
cat $OUT/loader_probe.jsonl | tee -a $OUT/summary.txt
tail -5 $OUT/loader_probe.err | cut -c1-300 | tee -a $OUT/summary.txt
