"""INTEGRATION.md section B, executed: the UNMODIFIED reference hook.c + client.c with the binding that section
shows (oracle/bind_reference.py applies it in a temporary directory and leaves only oracle/_ref/libnvshare_bound.so)
linked against our C-ABI library libnvs_engine.so.  Two oversubscribed clients under the REFERENCE daemon: the
reference's interposer, gate, protocol and threads, our data path -- allocations come from the engine (VMM chunks,
not managed memory), every hand-off evicts and fetches through nvs_evict / nvs_fetch_all, and every word comes
back (an access to an evicted slab is a SIGSEGV on the fake driver)."""
from __future__ import annotations

import json
import re
import subprocess
import sys

import pytest

from nvs_testlib import ORACLE, ROOT, Daemon, fake_env

pytestmark = pytest.mark.reference
BOUND = ORACLE / "libnvshare_bound.so"


@pytest.fixture(scope="module")
def bound(artefacts):
    r = subprocess.run([sys.executable, str(ROOT / "oracle" / "bind_reference.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not BOUND.exists():
        pytest.skip("reference sources not present and no prebuilt bound library")
    return BOUND


def test_the_bound_library_is_the_references_plus_our_c_abi(bound):
    dyn = subprocess.run(["readelf", "-d", str(bound)], capture_output=True, text=True).stdout
    assert "libnvs_engine.so" in dyn and "Library soname: [libnvshare.so]" in dyn
    syms = subprocess.run(["nm", "-D", "--undefined-only", str(bound)], capture_output=True, text=True).stdout
    used = set(re.findall(r"\bnvs_\w+", syms))
    # the whole binding: six calls of include/nvshare_engine.h (+ the config default)
    assert used == {"nvs_engine_default_config", "nvs_engine_create", "nvs_alloc", "nvs_free", "nvs_fetch_all", "nvs_evict",
                    "nvs_set_resident_mode"}, used


def test_two_oversubscribed_clients_under_the_reference_daemon(bound, default_sock_lock, tmp_path):
    d = Daemon("reference", default_sock_lock, log_path=tmp_path / "sched.log")
    try:
        procs = []
        d.ctl("-T", "1")
        for i in (1, 2):
            env = fake_env(total_mib=200, ledger=tmp_path / "hbm", trace=tmp_path / f"trace{i}.txt",
                           extra={"NVSHARE_HOST_ARENA_MIB": 64, "NVSHARE_CHUNK_MIB": 8, "NVSHARE_BATCH_MIB": 32, "NVSHARE_DEBUG": 1,
                                  "NVSHARE_STATS_FILE": tmp_path / f"stats{i}.jsonl"})
            env["LD_PRELOAD"] = str(bound)
            procs.append(subprocess.Popen([str(ORACLE / "driver_app"), "40", "4.0", str(i), "3"], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=120) for p in procs]
    finally:
        for p in procs:                     # (a client that is still there after a failure must not outlive the test)
            if p.poll() is None:
                p.kill()
        d.stop()
    for i, (p, (out, err)) in enumerate(zip(procs, outs), 1):
        assert p.returncode == 0 and re.search(r"RESULT PASS iters=\d+ mismatches=0", out), out + err[-2500:]
        trace = (tmp_path / f"trace{i}.txt").read_text()
        assert "cuMemAllocManaged" not in trace and "cuMemCreate" in trace        # the engine's memory, not UVM
        ops = [json.loads(l)["op"] for l in (tmp_path / f"stats{i}.jsonl").read_text().splitlines()]
        assert ops.count("evict") >= 1 and ops.count("fetch") >= 2                # hand-offs went through the C-ABI
        assert "[NVSHARE][DEBUG]: Received LOCK_OK" in err                        # the reference's own client code ran
