#!/bin/bash
# round 2, GPU call 12: the GPU ledger on hardware (smoke, the account rows of a real engine, one hooked test).
mkdir -p gpurun_out
{
timeout 100 python - <<'PY'
import __graft_entry__ as g
g.smoke()
from nvshare_b200 import engine as E
with E.Engine() as e:
    p = e.alloc(512 << 20)
    print("ledger own GPU:", e.gpu_account(-1), "lent:", e.gpu_lent_bytes())
    e.free(p)
PY
echo "rc=$?"
timeout 100 python -m pytest tests/test_gpu_hooked.py -x -q -m gpu -k "no_swap_when_everything_fits" 2>&1 | tail -3
} > gpurun_out/call12.txt 2>&1
tail -12 gpurun_out/call12.txt
