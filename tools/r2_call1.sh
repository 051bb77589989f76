#!/bin/bash
# round 2, GPU call 1 (1 GPU): fd hand-over probe, the whole GPU suite, smoke, a short full-scale bench of our arm
O=gpurun_out/r2c1; mkdir -p $O
{ nvidia-smi -L; nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/memory.current; df -h /dev/shm | tail -1; } > $O/box.txt 2>&1
timeout 180 tools/probe L > $O/probe_l.txt 2>&1; echo "probe L rc=$?"; cat $O/probe_l.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.txt
timeout 1200 python bench.py --steps 4 --warmup 4 --no-extras --keep $O/bench > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 6000 $O/bench.json; tail -5 $O/bench.err
du -sh $O
