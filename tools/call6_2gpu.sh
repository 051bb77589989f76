#!/bin/bash
# GPU call 6 (2 GPUs): peer-HBM tier tests, probe F, bench at N=2 under torchrun.
set -u
OUT=gpurun_out/call6
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== probe F (peer tier raw bandwidth)" | tee $OUT/summary.txt
timeout 600 ./tools/probe AF 8 > $OUT/probe_f.txt 2>&1; echo "probe F rc=$?" | tee -a $OUT/summary.txt
grep -E '"section":"F"|error' $OUT/probe_f.txt | cut -c1-300 | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
rows=[json.loads(l[6:]) for l in open('gpurun_out/call6/probe_f.txt') if l.startswith('PROBE {"section":"D"')]
for r in rows:
    print(r['dir'], r['variant'], 'grid', r['grid'], 'warps', r['warps'], f"{r['GBps_per_dir']:.0f} GB/s", 'mismatch', r['mismatches'])
PY
echo "== pytest peer tests" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -s > $OUT/pytest_peer.txt 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_peer.txt | tee -a $OUT/summary.txt
echo "== bench N=2 (torchrun)" | tee -a $OUT/summary.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --keep $OUT/bench_n2 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?" | tee -a $OUT/summary.txt
tail -1 $OUT/bench_n2.json | cut -c1-3000 | tee -a $OUT/summary.txt
tail -5 $OUT/bench_n2.err | cut -c1-300 | tee -a $OUT/summary.txt
find $OUT -name "client*.jsonl" -size +3M -exec truncate -s 3M {} \;
du -sh $OUT | tee -a $OUT/summary.txt
