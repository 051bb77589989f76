"""Regression tests for the round-1 advisor findings, CPU only:
free-with-work-pending synchronises first, the shared pool refuses files it did
not create itself, a pressure hint can never make the lock holder give the GPU
away, the daemon forwards pressure only from the holder and never lets
advisory frames pile up on a client that is not reading."""
from __future__ import annotations

import os
import socket
import struct
import subprocess
import sys
import textwrap
import time
from pathlib import Path

import pytest

from nvs_testlib import (DROP_LOCK, FAKE_DIR, LOCK_OK, LOCK_RELEASED, MSG_SIZE, ORACLE, REGISTER, REQ_LOCK, ROOT,
                         SCHED_ON, Daemon, MockClient, fake_env, pack, preload, unpack)

MiB = 1 << 20

PRELUDE = textwrap.dedent(f"""
    import ctypes as C, os, sys
    sys.path.insert(0, {str(ROOT)!r})
    fake = C.CDLL({str(FAKE_DIR / 'libcuda.so.1')!r}, mode=C.RTLD_GLOBAL)
    fake.cuInit(0); ctx = C.c_void_p(); fake.cuDevicePrimaryCtxRetain(C.byref(ctx), 0); fake.cuCtxSetCurrent(ctx)
    from nvshare_b200 import engine as E
    MiB = 1 << 20
""")


def test_free_of_resident_memory_synchronises_first(artefacts, tmp_path):
    """cuMemFree synchronises implicitly, cuMemUnmap/cuMemRelease do not: the engine
    must drain the context before it unmaps chunks that are on the GPU."""
    trace = tmp_path / "trace"
    code = PRELUDE + textwrap.dedent("""
        e = E.Engine(chunk_bytes=8 * MiB, host_arena_bytes=64 * MiB, prepin=0)
        p = e.alloc(16 * MiB); e.fetch_all()
        q = e.alloc(16 * MiB); e.evict(0)           # p and q swapped out
        e.fetch_all()
        print("MARK", flush=True); fake.cuMemGetInfo_v2(C.byref(C.c_size_t()), C.byref(C.c_size_t()))
        e.free(p)                                    # resident: must synchronise, then unmap
        e.evict(0)
        e.free(q)                                    # swapped out: nothing on the GPU, no sync needed
    """)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FAKE_CUDA_TRACE=str(trace)),
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = trace.read_text().splitlines()
    mark = max(i for i, l in enumerate(lines) if l.startswith("cuMemGetInfo"))
    tail = lines[mark + 1:]
    first_unmap = next(i for i, l in enumerate(tail) if l.startswith("cuMemUnmap"))
    assert any(l.startswith("cuCtxSynchronize") for l in tail[:first_unmap]), tail[:first_unmap + 1]
    # the free of swapped-out memory that follows the second eviction does not synchronise
    last_evict_unmap = max(i for i, l in enumerate(tail) if l.startswith("cuMemUnmap"))
    assert not any(l.startswith("cuCtxSynchronize") for l in tail[last_evict_unmap:]), tail[last_evict_unmap:]


@pytest.mark.parametrize("how", ["mode", "symlink", "header"])
def test_pool_file_that_is_not_ours_is_refused(artefacts, tmp_path, how):
    pool = tmp_path / "pool"
    victim = tmp_path / "victim"
    if how == "mode":
        pool.write_bytes(b"\0" * 4096)
        os.chmod(pool, 0o666)                       # planted by "somebody else": world-writable
    elif how == "symlink":
        victim.write_bytes(b"precious")
        os.symlink(victim, pool)
    else:
        # right owner and mode, but a header whose geometry does not match the file
        hdr = struct.pack("<QIIQQ", 0x6e767368504f4f4c, 3, 0, 1 << 40, 0)
        pool.write_bytes(hdr + b"\0" * ((8 << 20) + (64 << 20) - len(hdr)))
        os.chmod(pool, 0o600)
    code = PRELUDE + textwrap.dedent(f"""
        e = E.Engine(chunk_bytes=8 * MiB, host_arena_bytes=64 * MiB, shared_pool_path={str(pool)!r},
                     shared_pool_bytes=1 << 30, prepin=0)
        p = e.alloc(16 * MiB); e.fetch_all(); e.pattern_fill(p, 16 * MiB // 8, seed=3); e.evict(0)
        e.fetch_all(); print("BAD", e.pattern_verify(p, 16 * MiB // 8, seed=3))
        e.free(p); e.close()
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "BAD 0" in r.stdout, r.stdout + r.stderr      # private pool instead: still works
    assert "private pinned pool" in r.stderr
    if how == "symlink":
        assert victim.read_bytes() == b"precious"
    if how == "mode":
        assert pool.stat().st_size == 4096                                     # untouched, not unlinked


def test_pool_of_another_pid_namespace_is_not_attached(artefacts, tmp_path):
    """Units are owned by pid and dead owners' units are reaped with kill(pid, 0): that means nothing across pid
    namespaces (containers sharing /dev/shm), where every live owner would look dead.  Such a client must fall
    back to a private pool and leave the owner's units alone."""
    import shutil
    if not shutil.which("unshare") or subprocess.run(["unshare", "--pid", "--fork", "true"], capture_output=True).returncode != 0:
        pytest.skip("cannot create a pid namespace here")
    pool = tmp_path / "pool"
    owner_code = PRELUDE + textwrap.dedent(f"""
        e = E.Engine(chunk_bytes=8 * MiB, host_arena_bytes=64 * MiB, shared_pool_path={str(pool)!r},
                     shared_pool_bytes=256 * MiB, prepin=0, retain=0)
        p = e.alloc(32 * MiB); e.fetch_all(); e.pattern_fill(p, 32 * MiB // 8, seed=5); e.evict(0)   # 32 MiB live in the pool
        print("READY", flush=True); sys.stdin.readline()
        e.fetch_all(); print("BAD", e.pattern_verify(p, 32 * MiB // 8, seed=5), flush=True)
        e.free(p); e.close()
    """)
    guest_code = PRELUDE + textwrap.dedent(f"""
        e = E.Engine(chunk_bytes=8 * MiB, host_arena_bytes=64 * MiB, shared_pool_path={str(pool)!r},
                     shared_pool_bytes=256 * MiB, prepin=0, retain=0)
        p = e.alloc(64 * MiB); e.fetch_all(); e.pattern_fill(p, 64 * MiB // 8, seed=6); e.evict(0)
        e.fetch_all(); print("BAD", e.pattern_verify(p, 64 * MiB // 8, seed=6))
        e.free(p); e.close()
    """)
    owner = subprocess.Popen([sys.executable, "-c", owner_code], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
    try:
        assert "READY" in owner.stdout.readline()
        g = subprocess.run(["unshare", "--pid", "--fork", sys.executable, "-c", guest_code], capture_output=True, text=True, timeout=120)
        assert g.returncode == 0 and "BAD 0" in g.stdout, g.stdout + g.stderr
        assert "another pid namespace" in g.stderr and "private pinned pool" in g.stderr
        # the same program inside the owner's namespace does attach
        g2 = subprocess.run([sys.executable, "-c", guest_code], capture_output=True, text=True, timeout=120)
        assert g2.returncode == 0 and "BAD 0" in g2.stdout and "another pid namespace" not in g2.stderr, g2.stdout + g2.stderr
        out, err = owner.communicate("go\n", timeout=60)
        assert "BAD 0" in out, out + err                     # the owner's units were not reaped under it
    finally:
        if owner.poll() is None:
            owner.kill()


class ScriptedDaemon:
    """Just enough of nvshare-scheduler to drive one real client library by hand."""

    def __init__(self, sock_dir: Path):
        self.path = sock_dir / "scheduler.sock"
        self.l = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.l.bind(str(self.path))
        self.l.listen(4)
        self.c = None

    def accept_and_register(self, timeout=20):
        self.l.settimeout(timeout)
        self.c, _ = self.l.accept()
        m = self.recv()
        assert m["type"] == REGISTER
        data = b"%016x" % 0xabc + b"\0" + b"2"     # id + capability marker at data[17]
        self.c.sendall(pack(SCHED_ON, 7331, data))

    def recv(self, timeout=20):
        self.c.settimeout(timeout)
        buf = b""
        while len(buf) < MSG_SIZE:
            chunk = self.c.recv(MSG_SIZE - len(buf))
            assert chunk, "client closed the connection"
            buf += chunk
        return unpack(buf)

    def try_recv(self, timeout):
        try:
            return self.recv(timeout)
        except (socket.timeout, TimeoutError):
            return None

    def send(self, mtype, data=b"", msg_id=7331):
        self.c.sendall(pack(mtype, msg_id, data))

    def close(self):
        for s in (self.c, self.l):
            if s:
                s.close()


def test_pressure_hint_never_makes_the_holder_release(artefacts, sock_dir, tmp_path):
    """ADVICE: an "e<MiB>" DROP_LOCK that reaches the client while it holds the lock (a frame
    queued during its fetch) must not be taken for a quantum expiry."""
    d = ScriptedDaemon(sock_dir)
    env = fake_env(total_mib=400, extra={"NVSHARE_HOST_ARENA_MIB": 64, "NVSHARE_CHUNK_MIB": 8, "NVSHARE_DEBUG": 1,
                                         "NVSHARE_SOCK_DIR": sock_dir, "NVSHARE_POOL": "private"})
    env["LD_PRELOAD"] = preload("ours")
    p = subprocess.Popen([str(ORACLE / "driver_app"), "16", "3.0", "1"], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    try:
        d.accept_and_register()
        m = d.recv()
        assert m["type"] == REQ_LOCK
        d.send(LOCK_OK)
        time.sleep(0.5)
        d.send(DROP_LOCK, b"e100", msg_id=1337)                 # stale pressure hint to the HOLDER
        assert d.try_recv(1.0) is None                          # no LOCK_RELEASED: it keeps the GPU
        d.send(DROP_LOCK, b"w1n10", msg_id=1337)                # the real thing
        m = d.recv()
        assert m["type"] == LOCK_RELEASED
        m = d.recv()                                            # it still has work: asks again
        assert m["type"] == REQ_LOCK
        d.send(LOCK_OK)
        out, err = p.communicate(timeout=60)
    finally:
        if p.poll() is None:
            p.kill()
        d.close()
    assert p.returncode == 0 and "RESULT PASS" in out, out + err[-2000:]
    assert "pressure hint reached the lock holder: ignored" in err


def test_daemon_forwards_pressure_only_from_the_holder_and_coalesces(artefacts, tmp_path):
    sock_dir = tmp_path / "nvs"; sock_dir.mkdir()
    d = Daemon("ours", sock_dir, log_path=tmp_path / "sched.log")
    try:
        d.ctl("-T", "30")
        a, b, c = (MockClient(d.sock_path, n) for n in "abc")
        for x in (a, b, c):
            x.register()
        a.send(REQ_LOCK, data=b"n10"); a.expect(LOCK_OK)
        b.send(REQ_LOCK, data=b"n10")
        b.send(REQ_LOCK, data=b"p50")                          # b waits: it is in no position to press
        a.expect_nothing(0.4); c.expect_nothing(0.1)
        # the holder presses three times while c does not read its socket: c gets ONE frame, and stays connected
        for _ in range(3):
            a.send(REQ_LOCK, data=b"p70")
            time.sleep(0.05)
        assert c.expect(DROP_LOCK)["data"] == b"e70"
        c.expect_nothing(0.3)
        assert b.expect(DROP_LOCK)["data"] == b"e70"
        b.expect_nothing(0.3)
        a.send(REQ_LOCK, data=b"p80")                          # both have drained: the next hint goes through
        assert c.expect(DROP_LOCK)["data"] == b"e80"
        assert b.expect(DROP_LOCK)["data"] == b"e80"
        for x in (a, b, c):
            x.close()
    finally:
        d.stop()
    assert "does not hold the lock: ignored" in d.read_log()
