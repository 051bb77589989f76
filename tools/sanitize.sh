#!/bin/bash
# Builds the host side with AddressSanitizer + UBSan into a scratch directory and runs the
# CPU test-suite against it (fake driver).  The embedded cubin is reused from the normal build.
#   tools/sanitize.sh [pytest args...]       e.g. tools/sanitize.sh tests/test_e2e_fake.py -x
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${SAN_OUT:-/tmp/nvs_sanitized}
SRC=$ROOT/nvshare_b200/csrc
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
CF="-O1 -g -Wall -Wextra -std=gnu11 -fPIC -I$ROOT/include $SAN"
mkdir -p "$OUT"
make -C "$SRC" -s            # makes sure slab_copy_cubin.h exists
for f in engine gpu_ledger nvs_log hook client nvs_wire scheduler ctl; do gcc $CF -c "$SRC/$f.c" -o "$OUT/$f.o"; done
gcc -shared $SAN -Wl,-soname=libnvs_engine.so -Wl,--version-script="$SRC/libnvs_engine.ld" "$OUT"/{engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvs_engine.so" -ldl -lpthread
# the interposer exports dlsym itself, which AddressSanitizer's start-up cannot live with:
# libnvshare.so gets UBSan only (its engine code is covered by ASan through libnvs_engine.so)
UB="-fsanitize=undefined -fno-omit-frame-pointer"
for f in hook client nvs_wire engine gpu_ledger nvs_log; do gcc -O1 -g -Wall -Wextra -std=gnu11 -fPIC -I$ROOT/include $UB -c "$SRC/$f.c" -o "$OUT/ub_$f.o"; done
gcc -shared $UB -Wl,-soname=libnvshare.so -Wl,--version-script="$SRC/libnvshare.ld" "$OUT"/ub_{hook,client,nvs_wire,engine,gpu_ledger,nvs_log}.o -o "$OUT/libnvshare.so" -ldl -lpthread
gcc $SAN "$OUT"/{scheduler,nvs_wire,nvs_log}.o -o "$OUT/nvshare-scheduler"
gcc $SAN "$OUT"/{ctl,nvs_wire,nvs_log}.o -o "$OUT/nvsharectl"
cp "$ROOT/nvshare_b200/_build/slab_copy.cubin" "$OUT/" 2>/dev/null || true
ASAN_LIB=$(gcc -print-file-name=libasan.so)
cd "$ROOT"
# python itself and the intentionally crashing / exiting test programs are not leak-clean
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
export NVS_TEST_BUILD=$OUT NVS_ENGINE_LIB=$OUT/libnvs_engine.so
LD_PRELOAD=$ASAN_LIB exec python -m pytest -q -p no:cacheprovider -m "not gpu" \
    --deselect tests/test_engine_fake.py::test_evicted_memory_is_unmapped \
    --deselect tests/test_abi_and_oracle.py "${@:-tests}"
