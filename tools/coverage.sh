#!/bin/bash
# Builds the host side with gcov instrumentation into a scratch directory, runs the CPU test-suite
# against it (fake driver) and prints line coverage per source file plus the uncovered lines of the
# data path (engine.c, gpu_ledger.c).  Mirrors tools/sanitize.sh.
#   tools/coverage.sh [pytest args...]
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${COV_OUT:-/tmp/nvs_coverage}
SRC=$ROOT/nvshare_b200/csrc
CF="-O0 -g --coverage -std=gnu11 -fPIC -I$ROOT/include -fprofile-update=atomic"
rm -rf "$OUT"; mkdir -p "$OUT"
make -C "$SRC" -s
for f in engine gpu_ledger nvs_log hook client nvs_wire scheduler ctl; do gcc $CF -c "$SRC/$f.c" -o "$OUT/$f.o"; done
gcc -shared --coverage -Wl,-soname=libnvs_engine.so -Wl,--version-script="$SRC/libnvs_engine.ld" "$OUT"/{engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvs_engine.so" -ldl -lpthread
# (separate objects for the interposer: the same sources are linked into two libraries)
for f in hook client nvs_wire engine gpu_ledger nvs_log; do gcc $CF -c "$SRC/$f.c" -o "$OUT/pre_$f.o"; done
gcc -shared --coverage -Wl,-soname=libnvshare.so -Wl,--version-script="$SRC/libnvshare.ld" "$OUT"/pre_{hook,client,nvs_wire,engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvshare.so" -ldl -lpthread
gcc --coverage "$OUT"/{scheduler,nvs_wire,nvs_log}.o -o "$OUT/nvshare-scheduler"
gcc --coverage "$OUT"/{ctl,nvs_wire,nvs_log}.o -o "$OUT/nvsharectl"
cp "$ROOT/nvshare_b200/_build/slab_copy.cubin" "$OUT/" 2>/dev/null || true
cd "$ROOT"
export NVS_TEST_BUILD=$OUT NVS_ENGINE_LIB=$OUT/libnvs_engine.so
python -m pytest -q -p no:cacheprovider -m "not gpu" --deselect tests/test_abi_and_oracle.py "${@:-tests}" || true
cd "$OUT"
python - "$OUT" <<'PY'
import re, subprocess, sys
from collections import defaultdict
out = sys.argv[1]
report = {}
for f in ("engine", "gpu_ledger", "hook", "client", "scheduler", "nvs_wire", "ctl"):
    counts = defaultdict(int)          # line -> executions, summed over the libraries the file is linked into
    lines = {}
    for obj in (f"{f}.o", f"pre_{f}.o"):
        r = subprocess.run(["gcov", "-t", "-o", out, obj], capture_output=True, text=True, cwd=out)
        cur = None
        for l in r.stdout.splitlines():
            m = re.match(r"\s*([0-9]+\*?|#####|=====|-):\s*(\d+):(.*)", l)
            if not m:
                continue
            c, n, text = m.group(1), int(m.group(2)), m.group(3)
            if n == 0:
                if text.startswith("Source:"):      # (headers with inline functions get sections of their own)
                    cur = text.endswith(f"/{f}.c")
                continue
            if not cur or c == "-":
                continue
            lines[n] = text
            counts[n] += 0 if c in ("#####", "=====") else int(c.rstrip("*"))
    if lines:
        miss = sorted(n for n in lines if counts[n] == 0)
        report[f] = (len(lines), miss, lines)
        print(f"{f}.c: {100 * (len(lines) - len(miss)) // len(lines)} % of {len(lines)} lines")
for f in ("engine", "gpu_ledger"):
    if f in report:
        tot, miss, lines = report[f]
        print(f"== {f}.c: lines the suite never ran")
        for n in miss:
            print(f"{n:5d}: {lines[n].rstrip()[:150]}")
PY
