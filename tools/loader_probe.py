"""SURVEY 8f rank 3 measured on the real driver: client A (restated pytorch-add)
computes inside a long quantum; client B (tests/apps/load_app.c, a plain
driver-API binary) allocates 8 x 1 GiB, uploads and reads everything back.
Three arms: B un-hooked (raw driver speed, shares the GPU freely), B under
libnvshare.so with the lock-free copy path, B under libnvshare.so with every
copy gated like the reference.  Prints one JSON line per arm with B's wall time
and A's iteration rate while B was loading."""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
from nvs_testlib import ORACLE, Daemon, preload  # noqa: E402

TQ = 20
MIB, NBUF = 1024, 8


def arm(name, hooked, lockfree):
    tmp = Path(tempfile.mkdtemp(prefix="nvs_loader_"))
    sock = tmp / "nvs"
    sock.mkdir()
    d = Daemon("ours", sock, log_path=tmp / "sched.log")
    try:
        d.ctl("-T", str(TQ))
        env_a = dict(os.environ, LD_PRELOAD=preload("ours"), NVSHARE_SOCK_DIR=str(sock), PYTHONPATH=str(ROOT))
        a = subprocess.Popen([sys.executable, "-m", "nvshare_b200.workloads", "--kind", "add", "--n", "16000", "--iters",
                              "100000000", "--seconds", "45", "--pattern", "ones", "--log", str(tmp / "a.jsonl"), "--tag", "a"],
                             env=env_a, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        t_wait = time.time()
        while "Sent LOCK_OK" not in d.read_log():
            assert time.time() - t_wait < 180 and a.poll() is None
            time.sleep(0.2)
        time.sleep(3.0)
        env_b = dict(os.environ)
        if hooked:
            env_b.update(LD_PRELOAD=preload("ours"), NVSHARE_SOCK_DIR=str(sock), NVSHARE_LOCKFREE_COPY=str(int(lockfree)))
        t0 = time.time()
        b = subprocess.run([str(ORACLE / "load_app"), str(MIB), str(NBUF), "5"], env=env_b, capture_output=True, text=True,
                           timeout=600)
        t1 = time.time()
        out_a, err_a = a.communicate(timeout=600)
        iters = [json.loads(l) for l in (tmp / "a.jsonl").read_text().splitlines()]
        ts = [r["t"] for r in iters if r.get("event") == "iter"]
        during = [t for t in ts if t0 <= t <= t1]
        before = [t for t in ts if t0 - 3.0 <= t < t0]
        print(json.dumps({
            "arm": name, "b_result": b.stdout.strip().splitlines()[-1] if b.stdout.strip() else b.stderr[-300:],
            "b_wall_s": round(t1 - t0, 3), "b_bytes_each_way": MIB * NBUF << 20,
            "b_GBps_each_way_incl_startup": round(2 * (MIB * NBUF << 20) / 1e9 / (t1 - t0), 2),
            "a_iter_per_s_before": round(len(before) / 3.0, 1),
            "a_iter_per_s_while_b_loads": round(len(during) / max(t1 - t0, 1e-9), 1),
            "a_ok": a.returncode == 0 and out_a.startswith("PASS"),
            "b_asked_for_lock": "Sent REQ_LOCK" in b.stderr if hooked else None}), flush=True)
    finally:
        d.stop()


if __name__ == "__main__":
    arm("unhooked", False, False)
    arm("hooked_lockfree", True, True)
    arm("hooked_gated_like_reference", True, False)
