#!/bin/bash
set -u
OUT=gpurun_out/call19
mkdir -p $OUT
echo "== NUMA placement of the pinned pool: full-scale bench, placed from the GPU's CPUs (default) vs anywhere" | tee $OUT/summary.txt
nvidia-smi topo -m 2>/dev/null | head -4 | cut -c1-200 | tee -a $OUT/summary.txt
for arm in 1 0; do  # default first
  NVSHARE_NUMA=$arm timeout 700 python bench.py > $OUT/bench_numa$arm.json 2> $OUT/bench_numa$arm.err; echo "NVSHARE_NUMA=$arm rc=$?" | tee -a $OUT/summary.txt
  python - $OUT/bench_numa$arm.json <<'PY' | tee -a $OUT/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    dev = d.get("device", {})
    print({k: d.get(k) for k in ("stall_ms_per_handoff", "iter_per_s", "value", "verified")}, "e2e", d.get("e2e", {}).get("value"),
          "evict/fetch GB/s", dev.get("evict_GBps"), dev.get("fetch_GBps"), "wall ms", dev.get("wall_ms_mean"),
          "wait", dev.get("wait_ms_mean"), "pages/node", dev.get("pool_pages_per_numa_node"), "near cpus", dev.get("pool_placed_from_cpus"),
          "frac", d.get("roofline", {}).get("frac"), "peak", d.get("link_peak_GBps_measured"))
except Exception as ex:
    print("parse error", ex)
PY
done
