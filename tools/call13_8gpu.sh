#!/bin/bash
set -u
OUT=gpurun_out/call13
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== bench N=8 (torchrun), TQ 5" | tee $OUT/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --tq 5 --steps 3 --warmup 3 --keep $OUT/bench_n8 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n8.json | cut -c1-3000 | tee -a $OUT/summary.txt
tail -3 $OUT/bench_n8.err | cut -c1-300 | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +2M -exec truncate -s 2M {} \;
