#!/bin/bash
# round 2, GPU call 8 (1 GPU): the whole GPU suite on the final tree
O=gpurun_out/r2c8; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
