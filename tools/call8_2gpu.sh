#!/bin/bash
set -u
OUT=gpurun_out/call8
mkdir -p $OUT
echo "== pytest peer tests" | tee $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -s > $OUT/pytest_peer.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_peer.txt | tee -a $OUT/summary.txt
echo "== bench N=2 (torchrun)" | tee -a $OUT/summary.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --keep $OUT/bench_n2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n2.json | cut -c1-3000 | tee -a $OUT/summary.txt
echo "== bench N=2 reference arm (rank 0 only)" | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
