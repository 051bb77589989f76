#!/bin/bash
# round 2, GPU call 7 (1 GPU): the tree as it will be judged -- whole GPU suite, smoke(), the default bench with its sub-run
O=gpurun_out/r2c7; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 900 python bench.py --steps 4 --warmup 5 --keep $O/bench > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python tools/brief.py $O/bench.json; tail -3 $O/bench.err
rm -rf $O/*/*/sock; du -sh $O
