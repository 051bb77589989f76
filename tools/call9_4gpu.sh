#!/bin/bash
set -u
OUT=gpurun_out/call9
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
echo "== bench N=4 (torchrun)" | tee $OUT/summary.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --keep $OUT/bench_n4 > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n4.json | cut -c1-3000 | tee -a $OUT/summary.txt
tail -3 $OUT/bench_n4.err | cut -c1-300 | tee -a $OUT/summary.txt
nvidia-smi --query-gpu=index,memory.used --format=csv | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
