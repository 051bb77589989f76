#!/bin/bash
# round 2, GPU call 5 (2 GPUs): peer-tier tests, the N=2 bench as the driver launches it (with BASELINE
# config #4 as a sub-run), kernel vs copy-engine fetch on NVLink, 4 Llama clients at reduced scale (the
# machinery of config #5), ncu capture of a peer-tier launch
O=gpurun_out/r2c5; mkdir -p $O
nvidia-smi -L > $O/box.txt
timeout 600 python -m pytest tests/test_gpu_peer.py -q -s > $O/pytest_peer.txt 2>&1; echo "peer tests rc=$?"; grep -E "peer tier|passed|failed" $O/pytest_peer.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 6 --warmup 5 --keep $O/n2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; python tools/brief.py $O/bench_n2.json; tail -3 $O/bench_n2.err
NVSHARE_PEER_FETCH_VARIANT=tma timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 2 --steps 4 --warmup 4 --no-extras > $O/bench_n2_tmafetch.json 2> $O/bench_n2_tmafetch.err; echo "bench n2 tma-fetch rc=$?"; python tools/brief.py $O/bench_n2_tmafetch.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --kind llama --hbm-fraction 0.4 --steps 4 --warmup 3 --no-extras --keep $O/llama4 > $O/bench_llama4.json 2> $O/bench_llama4.err; echo "llama x4 (0.4 HBM) rc=$?"; python tools/brief.py $O/bench_llama4.json; tail -c 1500 $O/bench_llama4.json; tail -5 $O/bench_llama4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nvs_slab_copy_tma -c 6 -o $O/prof_peer python tools/ncu_target_r2.py --peer > $O/ncu_peer.txt 2>&1; echo "ncu peer rc=$?"
ncu -i $O/prof_peer.ncu-rep --page raw --csv > $O/prof_peer_raw.csv 2>/dev/null
python tools/ncu_summarise.py $O/prof_peer_raw.csv > $O/prof_peer_summary.json 2>&1
ncu --query-metrics 2>/dev/null | grep -i -E "nvl|pcie" | head -60 > $O/ncu_link_metrics.txt
rm -rf $O/*/*/sock; du -sh $O
