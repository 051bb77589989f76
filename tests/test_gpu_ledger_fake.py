"""Per-GPU accounting across processes and schedulers (nvshare_b200/csrc/gpu_ledger.c; SURVEY 8e:
"peers' HBM used as backing must be accounted in those GPUs' caps"), on CPU with the fake driver's
several "GPUs".  The reference counts per process and knows device 0 only (src/hook.c:77-78, 662;
src/client.c:386): there is nothing of its own to compare with, so these tests pin the rules
gpu_ledger.h states -- and that the data still comes back bit for bit when a GPU refuses to lend."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

import pytest

from nvs_testlib import ORACLE, ROOT, Daemon, fake_env, preload

MiB = 1 << 20
WORKER = ROOT / "tests" / "apps" / "ledger_worker.py"


class Worker:
    def __init__(self, tmp_path, peers=(), device=0, visible=None, total_mib=256, reserve_mib=16, extra=None):
        env = fake_env(total_mib=total_mib, ledger=tmp_path / "hbm", devices=4,
                       extra={"NVSHARE_GPU_LEDGER": tmp_path / "gpus", "NVSHARE_GPU_RESERVE_MIB": reserve_mib,
                              "WORKER_PEERS": peers if peers == "auto" else ",".join(map(str, peers)), "FAKE_CUDA_DEVICE": device,
                              "NVSHARE_POOL": "private", **(extra or {})})
        if visible:
            env["FAKE_CUDA_VISIBLE"] = visible
        self.p = subprocess.Popen([sys.executable, str(WORKER)], env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True)
        self.hello = self._read()

    def _read(self):
        line = self.p.stdout.readline()
        assert line, "worker died: " + self.p.stderr.read()[-3000:]
        return json.loads(line)

    def __call__(self, *cmd):
        self.p.stdin.write(" ".join(map(str, cmd)) + "\n")
        self.p.stdin.flush()
        return self._read()

    def quit(self):
        if self.p.poll() is None:
            self.p.stdin.write("quit\n")
            self.p.stdin.flush()
            self.p.wait(timeout=30)

    def die(self):
        self.p.stdin.write("die\n")
        self.p.stdin.flush()
        self.p.wait(timeout=10)


@pytest.fixture()
def workers(artefacts):
    made = []

    def make(*a, **kw):
        w = Worker(*a, **kw)
        made.append(w)
        return w
    yield make
    for w in made:
        try:
            w.quit()
        except Exception:
            w.p.kill()


def test_lending_is_budgeted_across_processes(workers, tmp_path):
    """Two clients of GPU 0 back their slabs on GPU 1 (256 MiB, 16 MiB reserve): together they may put
    240 MiB there, whoever comes second spills the rest to pinned host memory -- decided from the ledger,
    before the driver is asked -- and both get their bytes back."""
    a, b = workers(tmp_path, peers=[1]), workers(tmp_path, peers=[1])
    assert a("alloc", 160)["ok"]
    ra = a("evict")                                   # (GPU 0 itself only has room for one of them at a time)
    assert ra["peer_bytes"] == 160 * MiB and ra["host_bytes"] == 0
    assert b("alloc", 160)["ok"]
    rb = b("evict")
    assert rb["peer_bytes"] + rb["host_bytes"] == 160 * MiB and rb["host_bytes"] >= 80 * MiB
    acc_a, acc_b = a("account", 0), b("account", 0)
    assert acc_a["tracked"] == acc_b["tracked"] == 1
    assert acc_a["lent_bytes"] == acc_b["lent_bytes"] == acc_a["my_lent_bytes"] + acc_b["my_lent_bytes"]
    assert acc_a["lent_bytes"] <= (256 - 16) * MiB
    assert acc_b["refusals"] >= 1 and acc_a["refusals"] == 0
    # GPU 0 itself: nobody lends from it, and both clients' own footprints are on record there
    own = a("account", -1)
    assert own["lent_bytes"] == 0 and own["max_own_bytes"] == 160 * MiB and own["my_own_bytes"] == 160 * MiB
    assert a("fetch")["mismatches"] == 0
    # empty arenas go back at once: a's share is free again, b's is still out there
    assert a("account", 0)["lent_bytes"] == acc_b["my_lent_bytes"]
    assert a("evict")["peer_bytes"] == 160 * MiB       # 240 - b's 80: all of a fits again
    assert b("fetch")["mismatches"] == 0
    a("free")
    assert b("account", 0)["lent_bytes"] == 0


def test_a_gpu_does_not_lend_what_its_own_clients_need(workers, tmp_path):
    """GPU 1 has a client of its own with a 128 MiB footprint: of its 256 MiB, 16 are reserved, 128 are
    that client's whenever it holds GPU 1's lock, 112 are left for clients of GPU 0 to use as backing."""
    tenant = workers(tmp_path, device=1)
    assert tenant("alloc", 128)["ok"]
    assert tenant("account", -1)["my_own_bytes"] == 128 * MiB
    tenant("evict")                                    # swapped out or not: the claim is its footprint
    guest = workers(tmp_path, peers=[1])
    assert guest("alloc", 160)["ok"]
    rep = guest("evict")
    assert rep["peer_bytes"] <= 112 * MiB and rep["peer_bytes"] >= 96 * MiB      # whole 32 MiB arenas
    assert rep["peer_bytes"] + rep["host_bytes"] == 160 * MiB
    acc = guest("account", 0)
    assert acc["max_own_bytes"] == 128 * MiB and acc["lent_bytes"] == rep["peer_bytes"]
    # the tenant sees what its GPU has lent
    assert tenant("account", -1)["lent_bytes"] == rep["peer_bytes"]
    assert tenant("fetch")["mismatches"] == 0 and guest("fetch")["mismatches"] == 0
    # the tenant frees its memory: its claim goes, GPU 1 lends more
    tenant("free")
    assert guest("evict")["peer_bytes"] == 160 * MiB
    assert guest("fetch")["mismatches"] == 0


def test_gpus_are_identified_by_uuid_not_by_ordinal(workers, tmp_path):
    """CUDA_VISIBLE_DEVICES renumbers the GPUs per process: a client that sees physical GPU 1 as its
    device 0 and a client of GPU 0 that uses "device 1" as a peer are talking about the same HBM."""
    tenant = workers(tmp_path, device=0, visible="1,0")          # its device 0 is physical GPU 1
    guest = workers(tmp_path, peers=[1])                         # identity numbering
    assert tenant("alloc", 64)["ok"] and guest("alloc", 96)["ok"]
    rep = guest("evict")
    assert rep["peer_bytes"] == 96 * MiB
    assert tenant("account", -1)["lent_bytes"] == 96 * MiB
    assert guest("account", 0)["max_own_bytes"] == 64 * MiB
    assert guest("account", -1)["lent_bytes"] == 0               # physical GPU 0 lends nothing
    assert guest("fetch")["mismatches"] == 0


def test_claims_of_a_dead_client_are_dropped(workers, tmp_path):
    a = workers(tmp_path, peers=[1])
    assert a("alloc", 160)["ok"]
    assert a("evict")["peer_bytes"] == 160 * MiB
    a.die()                                                       # its arenas die with it (the driver frees them)
    b = workers(tmp_path, peers=[1])
    assert b("alloc", 200)["ok"]
    rep = b("evict")
    assert rep["peer_bytes"] == 200 * MiB, rep                    # the dead client's 160 MiB did not count
    acc = b("account", 0)
    assert acc["lent_bytes"] == acc["my_lent_bytes"] == 224 * MiB  # 7 arenas of 32 MiB
    assert b("fetch")["mismatches"] == 0


def test_a_foreign_ledger_file_is_not_used(workers, tmp_path):
    """Same rule as for the shared host pool: only a regular 0600 file of this very user.  Otherwise the
    accounting is off (loudly) and the data path is what it was."""
    (tmp_path / "gpus").write_bytes(b"\0" * 4096)
    os.chmod(tmp_path / "gpus", 0o666)
    a = workers(tmp_path, peers=[1])
    assert a("alloc", 64)["ok"]
    assert a("evict")["peer_bytes"] == 64 * MiB
    acc = a("account", 0)
    assert acc["tracked"] == 0 and acc["lent_bytes"] == 0
    assert a("fetch")["mismatches"] == 0
    a.quit()
    assert "not a ledger of ours" in a.p.stderr.read()


def test_ledger_can_be_switched_off(workers, tmp_path):
    a = workers(tmp_path, peers=[1], extra={"NVSHARE_GPU_LEDGER": "off"})
    assert a("alloc", 64)["ok"] and a("evict")["peer_bytes"] == 64 * MiB
    assert a("account", 0)["tracked"] == 0 and not (tmp_path / "gpus").exists()
    assert a("fetch")["mismatches"] == 0


def test_hooked_client_sees_what_its_gpu_has_lent(artefacts, workers, sock_dir, tmp_path):
    """The application's view (the hook): cuMemGetInfo's `free` and the cuMemAlloc cap of a client that
    computes on GPU 1 shrink by what GPU 1 has lent to clients of GPU 0, and grow back when it is returned.
    With nothing lent the numbers are the reference's (total - 1536 MiB, tests/golden/hook_golden.json)."""
    d = Daemon("ours", sock_dir)
    try:
        env = fake_env(total_mib=4096, ledger=tmp_path / "hbm", devices=4,
                       extra={"NVSHARE_GPU_LEDGER": tmp_path / "gpus", "FAKE_CUDA_DEVICE": 1, "NVSHARE_SOCK_DIR": sock_dir,
                              "NVSHARE_POOL": "private"})
        env["LD_PRELOAD"] = preload("ours")
        app = subprocess.Popen([str(ORACLE / "cap_app")], env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True)

        def ask(cmd):
            app.stdin.write(cmd + "\n")
            app.stdin.flush()
            return app.stdout.readline().strip()
        assert app.stdout.readline().startswith("ready device=1")
        assert ask("info") == "free_mib=2560 total_mib=4096 rc=0"
        assert ask("alloc 2500") == "alloc rc=0"
        guest = workers(tmp_path, peers=[1], total_mib=4096, reserve_mib=1536)
        assert guest("alloc", 512)["ok"] and guest("evict")["peer_bytes"] == 512 * MiB
        assert ask("info") == "free_mib=2048 total_mib=4096 rc=0"
        assert ask("alloc 2500") == "alloc rc=2"                 # CUDA_ERROR_OUT_OF_MEMORY: the cap moved
        assert ask("alloc 2000") == "alloc rc=0"
        assert guest("fetch")["mismatches"] == 0                  # arenas returned
        assert ask("info") == "free_mib=2560 total_mib=4096 rc=0"
        assert ask("alloc 2500") == "alloc rc=0"
        # and the other way round: what the application holds is not there to be lent
        assert ask("hold 2100") == "hold rc=0"
        rep = guest("evict")
        assert rep["peer_bytes"] + rep["host_bytes"] == 512 * MiB
        assert rep["peer_bytes"] <= (4096 - 1536 - 2100) * MiB and rep["host_bytes"] > 0
        assert guest("fetch")["mismatches"] == 0
        ask("quit")
        app.wait(timeout=30)
    finally:
        d.stop()


# layout of the ledger file (gpu_ledger.c): header 24 bytes, pthread_mutex_t 40, 32 devices x 32 bytes, then the
# claims: {i32 pid, u16 dev, u16 kind, u64 bytes, u64 start time of the pid}
_CLAIMS_AT, _CLAIM, _DEVS_AT, _DEV = 24 + 40 + 32 * 32, 24, 24 + 40, 32


def _start_time(pid):
    return int(Path(f"/proc/{pid}/stat").read_text().rsplit(")", 1)[1].split()[19])


def _plant_claim(path, gpu, kind, nbytes, pid, start):
    import struct
    raw = bytearray(Path(path).read_bytes())
    dev = next(i for i in range(32) if raw[_DEVS_AT + i * _DEV + 15] == gpu and raw[_DEVS_AT + i * _DEV + 24])
    slot = next(i for i in range(2048) if struct.unpack_from("<i", raw, _CLAIMS_AT + i * _CLAIM)[0] == 0)
    with open(path, "r+b") as f:
        f.seek(_CLAIMS_AT + slot * _CLAIM)
        f.write(struct.pack("<iHHQQ", pid, dev, kind, nbytes, start))


def test_a_recycled_pid_does_not_keep_a_dead_clients_claim(workers, tmp_path):
    """Owners are pids, and pids are recycled: a claim whose pid is alive again but started at another time
    than the claim says belongs to nobody.  Planted by hand: one claim of a live pid (this test process) with
    the wrong start time, one with the right one.  Only the second is counted."""
    a = workers(tmp_path, peers=[1])
    assert a("alloc", 64)["ok"] and a("evict")["peer_bytes"] == 64 * MiB
    me = os.getpid()
    _plant_claim(tmp_path / "gpus", 1, 1, 100 * MiB, me, _start_time(me) + 12345)      # kind 1 = lent
    time.sleep(1.1)                                                                     # (the dead are looked for once a second)
    assert a("account", 0)["lent_bytes"] == 64 * MiB
    _plant_claim(tmp_path / "gpus", 1, 1, 50 * MiB, me, _start_time(me))
    time.sleep(1.1)
    assert a("account", 0)["lent_bytes"] == (64 + 50) * MiB
    assert a("fetch")["mismatches"] == 0


def test_transfer_records_carry_the_ledgers_view_of_the_peers(workers, tmp_path):
    """What bench.py puts into the N >= 2 line (`device.gpu_ledger`): every evict / fetch record of an engine with a
    peer tier says what the ledger holds for its peers -- all clients' arenas, its own claim -- next to what the
    engine itself has mapped there."""
    import json
    a = workers(tmp_path, peers=[1], extra={"NVSHARE_STATS_FILE": tmp_path / "a.jsonl"})
    b = workers(tmp_path, peers=[1], extra={"NVSHARE_STATS_FILE": tmp_path / "b.jsonl"})
    assert a("alloc", 96)["ok"] and a("evict")["peer_bytes"] == 96 * MiB
    assert b("alloc", 64)["ok"] and b("evict")["peer_bytes"] == 64 * MiB
    ra = [json.loads(l) for l in (tmp_path / "a.jsonl").read_text().splitlines() if '"op":"evict"' in l][-1]
    rb = [json.loads(l) for l in (tmp_path / "b.jsonl").read_text().splitlines() if '"op":"evict"' in l][-1]
    assert ra["gl_tracked_peers"] == rb["gl_tracked_peers"] == 1
    assert ra["gl_my_lent"] == ra["peer_pool_bytes"] == 96 * MiB and ra["gl_lent"] == 96 * MiB      # a was alone then
    assert rb["gl_my_lent"] == rb["peer_pool_bytes"] == 64 * MiB and rb["gl_lent"] == (96 + 64) * MiB
    assert ra["gl_refusals"] == rb["gl_refusals"] == 0
    assert a("fetch")["mismatches"] == 0 and b("fetch")["mismatches"] == 0
    fa = [json.loads(l) for l in (tmp_path / "a.jsonl").read_text().splitlines() if '"op":"fetch"' in l][-1]
    assert fa["gl_my_lent"] == fa["peer_pool_bytes"] == 0                                            # arenas went back


def test_peers_auto_uses_every_other_reachable_gpu_and_respects_their_tenants(workers, tmp_path):
    """NVSHARE_PEERS=auto (NVS_PEERS_AUTO): every other visible GPU the engine's GPU can reach backs its slabs,
    striped; the ledger is what makes "all of them" safe -- a GPU with a tenant lends only what the tenant leaves."""
    tenant = workers(tmp_path, device=2)                           # physical GPU 2 has a client of its own
    assert tenant("alloc", 192)["ok"]                              # of 256 MiB: 16 reserved, 192 its own, 48 to lend
    guest = workers(tmp_path, peers="auto")                        # computes on GPU 0; peers 1, 2, 3 (four fake GPUs)
    assert guest("alloc", 240)["ok"]
    rep = guest("evict")
    assert rep["peer_bytes"] == 240 * MiB and rep["host_bytes"] == 0      # the three peers hold it all
    accs = [guest("account", i) for i in range(3)]
    assert [a["device"] for a in accs] == [1, 2, 3] and all(a["tracked"] == 1 for a in accs)
    assert sum(a["my_lent_bytes"] for a in accs) >= 240 * MiB
    assert accs[1]["my_lent_bytes"] <= 48 * MiB and accs[1]["max_own_bytes"] == 192 * MiB   # GPU 2 kept its tenant's room
    assert accs[0]["my_lent_bytes"] >= 64 * MiB and accs[2]["my_lent_bytes"] >= 64 * MiB    # striping used the others
    assert guest("fetch")["mismatches"] == 0 and tenant("fetch")["mismatches"] == 0


def test_a_ledger_of_another_layout_is_left_alone(workers, tmp_path):
    """An explicitly named ledger (NVSHARE_GPU_LEDGER) that is ours by owner and mode but has another layout's size
    -- an older client's -- is not used and not waited for: the accounting is off, loudly, at once; the data path works.
    (The default name carries the layout version, so two versions on one node keep two files.)"""
    old = tmp_path / "gpus"
    old.write_bytes(b"\0" * 33856)
    os.chmod(old, 0o600)
    os.utime(old, (time.time() - 60, time.time() - 60))
    t0 = time.time()
    a = workers(tmp_path, peers=[1])
    assert time.time() - t0 < 1.9                                 # (a young short file is waited for up to 2 s: its creator is sizing it)
    assert a("alloc", 64)["ok"] and a("evict")["peer_bytes"] == 64 * MiB
    assert a("account", 0)["tracked"] == 0
    assert a("fetch")["mismatches"] == 0
    a.quit()
    assert "unexpected size" in a.p.stderr.read()
    assert old.stat().st_size == 33856                            # untouched
