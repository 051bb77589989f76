#!/bin/bash
# ThreadSanitizer pass over the host side (fake driver): the engine C-ABI tests in-process (python + ctypes)
# and the multi-threaded interposed application (tests/apps/mt_app.c, 3 threads per client) under one scheduler
# with tight shared pools.  The interposer is built WITHOUT its dlsym export here (-DNVS_NO_DLSYM_EXPORT:
# TSan's start-up cannot live with an interposed dlsym; mt_app binds the driver entry points directly).
# Reads of the fake HBM by the fake driver's "kernels" while the application writes it are by design (the
# background pre-cleaner reads application data; the content hash decides whether the copy is kept) and are
# suppressed by library (called_from_lib:libcuda.so.1).  Exit status 1 when any report is left.
#   tools/tsan.sh
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TSAN_OUT:-/tmp/nvs_tsan}
SRC=$ROOT/nvshare_b200/csrc
TS="-fsanitize=thread -fno-omit-frame-pointer"
CF="-O1 -g -std=gnu11 -fPIC -I$ROOT/include $TS -DNVS_NO_DLSYM_EXPORT"
rm -rf "$OUT"; mkdir -p "$OUT/logs"
make -C "$SRC" -s
make -C "$ROOT/oracle" -s
for f in hook client nvs_wire engine gpu_ledger nvs_log; do gcc $CF -c "$SRC/$f.c" -o "$OUT/$f.o"; done
gcc -shared $TS -Wl,-soname=libnvshare.so "$OUT"/{hook,client,nvs_wire,engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvshare.so" -ldl -lpthread
gcc -shared $TS -Wl,-soname=libnvs_engine.so -Wl,--version-script="$SRC/libnvs_engine.ld" "$OUT"/{engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvs_engine.so" -ldl -lpthread
cp "$ROOT/nvshare_b200/_build/slab_copy.cubin" "$OUT/" 2>/dev/null || true
echo "called_from_lib:libcuda.so.1" > "$OUT/supp.txt"
LIBTSAN=$(gcc -print-file-name=libtsan.so)
cd "$ROOT"
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 suppressions=$OUT/supp.txt log_path=$OUT/logs/t"
echo "== engine C-ABI tests under TSan"
NVS_ENGINE_LIB=$OUT/libnvs_engine.so LD_PRELOAD=$LIBTSAN python -m pytest -q -p no:cacheprovider \
    --deselect tests/test_engine_fake.py::test_evicted_memory_is_unmapped \
    tests/test_engine_fake.py tests/test_retention_fake.py tests/test_policy_fake.py tests/test_engine_property.py tests/test_advice_r1.py || true
echo "== interposed multi-threaded clients under TSan"
NVS_TSAN_DIR=$OUT NVS_TSAN_LIB=$LIBTSAN python - <<'PY'
import os, sys, subprocess, tempfile, pathlib
sys.path.insert(0, "tests")
from nvs_testlib import ORACLE, Daemon, fake_env
out = os.environ["NVS_TSAN_DIR"]
def run(n_clients, pool_mib, policy, seconds=6, total=200, threads=3, mib=12):
    tmp = pathlib.Path(tempfile.mkdtemp()); sd = tmp / "nvs"; sd.mkdir()
    d = Daemon("ours", sd, log_path=tmp / "sched.log")
    try:
        d.ctl("-T", "1")
        procs = []
        for i in range(n_clients):
            env = fake_env(total_mib=total, ledger=tmp / "ledger", extra={
                "NVSHARE_HOST_ARENA_MIB": 64, "NVSHARE_CHUNK_MIB": 8, "NVSHARE_BATCH_MIB": 32, "NVSHARE_SOCK_DIR": sd,
                "NVSHARE_POOL_MIB": pool_mib, "NVSHARE_EVICT_POLICY": policy, "NVSHARE_OOM_WAIT_MS": 15000})
            env["LD_PRELOAD"] = os.environ["NVS_TSAN_LIB"] + ":" + out + "/libnvshare.so"
            procs.append(subprocess.Popen([str(ORACLE / "mt_app"), str(mib), str(seconds), str(i + 1), str(threads), "1"],
                                          env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=900) for p in procs]
        ok = all("RESULT PASS" in o for o, _ in outs)
        print(f"clients={n_clients} pool={pool_mib} policy={policy}: data {'ok' if ok else 'MISMATCH'}", flush=True)
        return ok
    finally:
        d.stop()
ok = True
for cfg in [(2, 128, "need"), (3, 64, "all"), (3, 128, "need"), (4, 192, "need")]:
    ok &= run(*cfg)
sys.exit(0 if ok else 1)
PY
n=$(cat "$OUT"/logs/t.* 2>/dev/null | grep -c "^SUMMARY: ThreadSanitizer" || true)
# the one deliberate crash test (a touch of unmapped memory) leaves a DEADLYSIGNAL line, not a report
echo "ThreadSanitizer reports: $n"
cat "$OUT"/logs/t.* 2>/dev/null | grep "^SUMMARY: ThreadSanitizer" | sort | uniq -c | sort -rn | head -20
[ "$n" = 0 ]
