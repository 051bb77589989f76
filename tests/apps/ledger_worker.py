"""One swap engine in its own process, driven line by line over stdin (tests/test_gpu_ledger_fake.py):
the fake driver is libcuda here (LD_LIBRARY_PATH), FAKE_CUDA_DEVICE / FAKE_CUDA_VISIBLE say which
"GPU" this process computes on and how it numbers the GPUs.  Every command answers with one JSON line.

  alloc <MiB>        allocate, make resident, fill with the position-dependent pattern (seed = pid)
  evict              evict everything; answers with the transfer report
  fetch              fetch everything back and verify every allocation bit for bit
  account <which>    nvs_gpu_account_query (-1 = own GPU, i = peer i)
  free               free every allocation
  die                exit without closing the engine (its ledger claims stay behind, its "HBM" does not)
  quit
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
MiB = 1 << 20


def main():
    fake = C.CDLL("libcuda.so.1", mode=C.RTLD_GLOBAL)
    assert fake.cuInit(0) == 0
    dev = C.c_int()
    assert fake.cuCtxGetDevice(C.byref(dev)) == 0
    ctx = C.c_void_p()
    assert fake.cuDevicePrimaryCtxRetain(C.byref(ctx), dev.value) == 0 and fake.cuCtxSetCurrent(ctx) == 0
    from nvshare_b200 import engine as E
    wp = os.environ.get("WORKER_PEERS", "")
    peers = "auto" if wp == "auto" else [int(x) for x in wp.split(",") if x]
    e = E.Engine(chunk_bytes=8 * MiB, host_arena_bytes=32 * MiB, batch_bytes=32 * MiB, peers=peers, prepin=0,
                 elide_constant=0, retain=0, oom_wait_ms=2000)
    allocs = []
    seed = os.getpid()
    print(json.dumps({"ready": True, "device": dev.value, "pid": os.getpid()}), flush=True)
    for line in sys.stdin:
        cmd = line.split()
        if not cmd:
            continue
        if cmd[0] == "alloc":
            n = int(cmd[1]) * MiB
            p = e.alloc(n)
            e.fetch_all()
            e.pattern_fill(p, n // 8, seed=seed + len(allocs))
            allocs.append((p, n))
            out = {"ok": True}
        elif cmd[0] == "evict":
            out = e.evict(0)
        elif cmd[0] == "fetch":
            out = e.fetch_all()
            out["mismatches"] = sum(e.pattern_verify(p, n // 8, seed=seed + i) for i, (p, n) in enumerate(allocs))
        elif cmd[0] == "account":
            out = e.gpu_account(int(cmd[1]))
        elif cmd[0] == "free":
            for p, _ in allocs:
                e.free(p)
            allocs = []
            out = {"ok": True}
        elif cmd[0] == "die":
            # the process ends without the engine being closed: libc's exit() runs the fake driver's destructor
            # (a real driver frees a dead process's HBM too) but nothing of ours
            sys.stdout.flush()
            C.CDLL(None).exit(0)
        elif cmd[0] == "quit":
            break
        else:
            out = {"error": "unknown command"}
        print(json.dumps(out), flush=True)
    e.close()


if __name__ == "__main__":
    main()
