#!/bin/bash
# round 2, GPU call 6 (4 GPUs): BASELINE config #5 at its full per-client scale -- four Llama-7B decode clients at
# 3.0x HBM, backing striped over three peers -- as a direct bench line
O=gpurun_out/r2c6; mkdir -p $O
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --kind llama --steps 4 --warmup 2 --no-extras --keep $O/llama > $O/bench_llama_n4.json 2> $O/bench_llama_n4.err; echo "llama x4 full scale on 4 GPUs rc=$?"
python tools/brief.py $O/bench_llama_n4.json; tail -c 1200 $O/bench_llama_n4.json; tail -4 $O/bench_llama_n4.err
grep -h '"setup_done"\|"summary"' $O/llama/main/client*.jsonl | cut -c1-300
rm -rf $O/*/*/sock; du -sh $O
