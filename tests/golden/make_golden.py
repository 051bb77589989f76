#!/usr/bin/env python3
"""Regenerate tests/golden/*.json from the UNMODIFIED reference binaries in
oracle/_ref (built from /root/reference by oracle/Makefile).  Run it in the
build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures pin what the reference *does* (not what its source seems to say):
  ctl_golden.json        nvsharectl: stdout/stderr/exit code per argument list with
                         the daemon down, and the exact 537-byte frames it sends
                         when something is listening
  hook_golden.json       libnvshare.so under a fake CUDA driver: what the
                         application observes (return codes, reserve) and the
                         filtered driver-call trace (launch / sync pattern)
  scheduler_golden.json  a scripted three-client scenario: which message types
                         each client receives, in order
The reference pins no numerical result for the data path (its tests print PASS
unconditionally), so there is no data fixture from it; data parity is defined
in tests/ against oracle/nvshare_oracle.c.
"""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))

from nvs_testlib import (DROP_LOCK, LOCK_OK, LOCK_RELEASED, MSG_SIZE, ORACLE, REQ_LOCK, TYPE_NAMES, Daemon,  # noqa: E402
                         MockClient, fake_env, preload, unpack)

CTL_ARGS = [[], ["-h"], ["--help"], ["-S", "maybe"], ["-T", "-3"], ["--bogus"], ["-x"], ["foo"], ["-T", "abc"],
            ["-T", "0"], ["-T"], ["-S"], ["-T5"], ["--help=1"], ["-hT"], ["--set-tq=abc"], ["--anti-thrash"],
            ["-T", "1.5"], ["-T", "5"], ["--set-tq=7"]]


def ctl_golden():
    exe = ORACLE / "nvsharectl"
    out = {"daemon_down": [], "frames": []}
    sock_path = Path("/var/run/nvshare/scheduler.sock")
    if sock_path.exists():
        sock_path.unlink()
    for args in CTL_ARGS:
        r = subprocess.run([str(exe), *args], capture_output=True, text=True)
        out["daemon_down"].append({"args": args, "rc": r.returncode, "stdout": r.stdout, "stderr": r.stderr})
    # a listener that just records frames
    os.makedirs(sock_path.parent, exist_ok=True)
    for args in (["-T", "5"], ["--set-tq=7"], ["-S", "on"], ["-S", "off"], ["-S", "off", "-T", "9"]):
        if sock_path.exists():
            sock_path.unlink()
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(str(sock_path))
        srv.listen(4)
        srv.settimeout(2)
        p = subprocess.Popen([str(exe), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        frames = []
        try:
            while True:
                conn, _ = srv.accept()
                buf = b""
                conn.settimeout(2)
                while len(buf) < MSG_SIZE:
                    chunk = conn.recv(MSG_SIZE - len(buf))
                    if not chunk:
                        break
                    buf += chunk
                m = unpack(buf)
                frames.append({"type": m["type"], "id": m["id"], "data": m["data"].decode(),
                               "pod_name": m["pod_name"].decode(), "pod_namespace": m["pod_namespace"].decode()})
                conn.close()
        except socket.timeout:
            pass
        so, se = p.communicate(timeout=5)
        srv.close()
        out["frames"].append({"args": args, "rc": p.returncode, "stderr": se, "frames": frames})
    if sock_path.exists():
        sock_path.unlink()
    return out


FILTER = ("cuInit", "cuLaunchKernel", "cuCtxSynchronize", "cuMemcpyHtoD", "cuMemcpyDtoH")


def run_trace_app(lib_impl, sched_dir, tmp, alloc_mib=100 * 1024, total_mib=192 * 1024, extra_env=None):
    trace = Path(tmp) / f"trace_{lib_impl}.txt"
    if trace.exists():
        trace.unlink()
    env = fake_env(total_mib=total_mib, trace=trace, extra=extra_env)
    env["LD_PRELOAD"] = preload(lib_impl)
    if lib_impl == "ours":
        env["NVSHARE_SOCK_DIR"] = str(sched_dir)
    r = subprocess.run([str(ORACLE / "trace_app"), str(alloc_mib), "3"], env=env, capture_output=True, text=True,
                       timeout=120)
    # the engine's own kernels (background pre-cleaning, scans) are not the application's calls
    names = [ln.split()[0] for ln in trace.read_text().splitlines()
             if ln.split() and not (len(ln.split()) > 1 and ln.split()[1].startswith("nvs_slab"))]
    # The real cuMemFree (what the reference's hook calls) synchronises INSIDE the driver, where no trace
    # sees it; our engine frees VMM memory with cuMemUnmap, which does not, so it drains the context
    # itself first.  That explicit call stands for the implicit one and is not part of the comparison.
    calls = [n for i, n in enumerate(names)
             if n in FILTER and not (n == "cuCtxSynchronize" and i + 1 < len(names) and names[i + 1] == "cuMemUnmap")]
    return {"rc": r.returncode, "stdout": r.stdout.splitlines(), "calls": calls, "stderr": r.stderr}


def hook_golden(tmp):
    d = Daemon("reference", Path("/var/run/nvshare"))
    try:
        g = run_trace_app("reference", d.sock_dir, tmp)
    finally:
        d.stop()
    g.pop("stderr")
    return g


def scheduler_scenario(d):
    """Three clients, TQ=1: returns {client: [message type names received]}."""
    d.ctl("-T", "1")
    cs = {n: MockClient(d.sock_path, n) for n in "abc"}
    got = {n: [] for n in cs}
    for c in cs.values():
        got[c.name].append(TYPE_NAMES[c.register()["type"]])
    for n in "abc":
        cs[n].send(REQ_LOCK)
        time.sleep(0.05)
    deadline = time.time() + 6
    released = set()
    while time.time() < deadline and len(released) < 3:
        for n, c in cs.items():
            m = c.recv(0.05)
            if m in (None, b""):
                continue
            got[n].append(TYPE_NAMES[m["type"]])
            if m["type"] == DROP_LOCK:
                c.send(LOCK_RELEASED)
                released.add(n)
    for c in cs.values():
        c.close()
    return got


def scheduler_golden():
    d = Daemon("reference", Path("/var/run/nvshare"))
    try:
        return scheduler_scenario(d)
    finally:
        d.stop()


def main():
    import tempfile
    if not (ORACLE / "nvshare-scheduler").exists():
        sys.exit("oracle/_ref is not built; run `make -C oracle` where /root/reference exists")
    with tempfile.TemporaryDirectory() as tmp:
        json.dump(ctl_golden(), open(HERE / "ctl_golden.json", "w"), indent=1)
        json.dump(hook_golden(tmp), open(HERE / "hook_golden.json", "w"), indent=1)
        json.dump(scheduler_golden(), open(HERE / "scheduler_golden.json", "w"), indent=1)
    print("wrote", [p.name for p in HERE.glob("*.json")])


if __name__ == "__main__":
    main()
