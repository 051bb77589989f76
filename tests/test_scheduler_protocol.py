"""BASELINE config #1: nvshare-scheduler + scripted clients over the Unix socket,
CPU only.  Every behavioural golden of SURVEY section 4 is asserted against
OUR daemon and, with the same test body, against the REFERENCE daemon compiled
into oracle/_ref -- so the expectations below are pinned by the reference
binary itself, not by our reading of its source.

Reference lines: src/scheduler.c:159-206 (register), :123-155 (queue),
:295-316 (try_schedule), :329-390 (timer), :393-501 (process_msg),
:641-663 (client death).
"""
from __future__ import annotations

import time

import pytest

from nvs_testlib import (DROP_LOCK, LOCK_OK, LOCK_RELEASED, MSG_SIZE, REGISTER, REQ_LOCK, SCHED_OFF, SCHED_ON,
                         SET_TQ, Daemon, MockClient, pack)

IMPLS = ["ours", pytest.param("reference", marks=pytest.mark.reference)]


@pytest.fixture(params=IMPLS)
def daemon(request, artefacts, tmp_path):
    impl = request.param
    if impl == "reference":
        sock_dir = request.getfixturevalue("default_sock_lock")
    else:
        sock_dir = tmp_path / "nvs"
        sock_dir.mkdir()
    d = Daemon(impl, sock_dir, log_path=tmp_path / "sched.log")
    yield d
    d.stop()


def client(d, name="c"):
    return MockClient(d.sock_path, name)


def test_register_reply(daemon):
    c = client(daemon)
    m = c.register(pod_name=b"pod-a", pod_namespace=b"ns-a")
    assert m["type"] == SCHED_ON            # scheduler starts ON (src/scheduler.c:550)
    assert m["id"] == 7331                  # src/scheduler.c:591
    assert len(m["data"]) == 16 and m["data"] == m["data"].lower()
    int(m["data"], 16)
    assert m["pod_name"] == b"" and m["pod_namespace"] == b""
    c2 = client(daemon)
    m2 = c2.register()
    assert m2["data"] != m["data"]          # ids are unique
    c.close(); c2.close()


def test_unregistered_req_lock_closes_connection(daemon):
    c = client(daemon)
    c.send(REQ_LOCK)
    c.expect_closed()


def test_unregistered_lock_released_closes_connection(daemon):
    c = client(daemon)
    c.send(LOCK_RELEASED)
    c.expect_closed()


def test_second_register_closes_connection(daemon):
    c = client(daemon)
    c.register()
    c.send(REGISTER)
    c.expect_closed()


def test_short_frame_closes_connection(daemon):
    c = client(daemon)
    c.register()
    c.send_raw(pack(REQ_LOCK, c.id)[:100])
    c.expect_closed()


def test_unknown_type_is_ignored(daemon):
    c = client(daemon)
    c.register()
    c.send(42)
    c.expect_nothing()
    c.send(REQ_LOCK)
    c.expect(LOCK_OK)


def test_fcfs_and_duplicate_request(daemon):
    daemon.ctl("-T", "30")
    a, b, c = client(daemon, "a"), client(daemon, "b"), client(daemon, "c")
    for x in (a, b, c):
        x.register()
    a.send(REQ_LOCK)
    m = a.expect(LOCK_OK)
    assert m["id"] == 7331
    b.send(REQ_LOCK)
    b.send(REQ_LOCK)            # duplicate: warned about, ignored
    c.send(REQ_LOCK)
    b.expect_nothing()
    a.send(LOCK_RELEASED)
    b.expect(LOCK_OK)           # first come first served
    c.expect_nothing()
    b.send(LOCK_RELEASED)
    c.expect(LOCK_OK)
    c.send(LOCK_RELEASED)
    b.expect_nothing(0.2)       # the duplicate did not leave a second request behind


def test_waiter_can_cancel(daemon):
    daemon.ctl("-T", "30")
    a, b, c = client(daemon, "a"), client(daemon, "b"), client(daemon, "c")
    for x in (a, b, c):
        x.register()
    a.send(REQ_LOCK); a.expect(LOCK_OK)
    b.send(REQ_LOCK); c.send(REQ_LOCK)
    time.sleep(0.1)
    b.send(LOCK_RELEASED)       # from a waiter: cancels its request
    time.sleep(0.1)
    a.send(LOCK_RELEASED)
    c.expect(LOCK_OK)
    b.expect_nothing()


def test_holder_death_passes_the_lock_on(daemon):
    daemon.ctl("-T", "30")
    a, b = client(daemon, "a"), client(daemon, "b")
    a.register(); b.register()
    a.send(REQ_LOCK); a.expect(LOCK_OK)
    b.send(REQ_LOCK); b.expect_nothing()
    a.close()
    b.expect(LOCK_OK)


def test_time_quantum_sequence(daemon):
    """TQ=1: A LOCK_OK@0, A DROP_LOCK@~1 s, B LOCK_OK right after A's release,
    B DROP_LOCK ~1 s later although nobody is waiting (SURVEY section 4 probe)."""
    daemon.ctl("-T", "1")
    a, b = client(daemon, "a"), client(daemon, "b")
    a.register(); b.register()
    t0 = time.time()
    a.send(REQ_LOCK); a.expect(LOCK_OK)
    b.send(REQ_LOCK)
    m = a.expect(DROP_LOCK, timeout=3)
    t_drop_a = time.time() - t0
    assert m["id"] == 1337                      # src/scheduler.c:337
    assert 0.8 <= t_drop_a <= 1.6
    a.expect_nothing(0.3)                       # DROP_LOCK is sent once
    a.send(LOCK_RELEASED)
    t1 = time.time()
    b.expect(LOCK_OK)
    m = b.expect(DROP_LOCK, timeout=3)          # nobody waits, still dropped
    assert 0.8 <= time.time() - t1 <= 1.6
    b.send(LOCK_RELEASED)
    b.expect_nothing(0.3)


def test_set_tq_restarts_the_quantum(daemon):
    daemon.ctl("-T", "30")
    a = client(daemon, "a")
    a.register()
    a.send(REQ_LOCK); a.expect(LOCK_OK)
    a.expect_nothing(0.5)
    t0 = time.time()
    daemon.ctl("-T", "1")
    a.expect(DROP_LOCK, timeout=3)
    assert 0.8 <= time.time() - t0 <= 1.6
    log = daemon.read_log()
    assert "New TQ = 1" in log


def test_set_tq_accepts_base_prefixes_and_rejects_garbage(daemon):
    x = client(daemon, "ctl")
    x.send(SET_TQ, data=b"0x10", msg_id=0xBEEF)      # strtoll(base 0), src/scheduler.c:454
    x.send(SET_TQ, data=b"12abc", msg_id=0xBEEF)
    x.send(SET_TQ, data=b"", msg_id=0xBEEF)
    time.sleep(0.3)
    log = daemon.read_log()
    assert "New TQ = 16" in log
    assert log.count("Failed to parse new TQ from message") == 2


def test_sched_off_and_on(daemon):
    daemon.ctl("-T", "30")
    a, b = client(daemon, "a"), client(daemon, "b")
    a.register(); b.register()
    a.send(REQ_LOCK); a.expect(LOCK_OK)
    b.send(REQ_LOCK)
    time.sleep(0.1)
    daemon.ctl("-S", "off")
    a.expect(SCHED_OFF); b.expect(SCHED_OFF)       # broadcast to every registered client
    # the queue was emptied: requests are ignored while off
    b.send(REQ_LOCK)
    b.expect_nothing()
    late = client(daemon, "late")
    assert late.register()["type"] == SCHED_OFF    # newcomers learn the status
    daemon.ctl("-S", "off")                        # no change: no broadcast
    a.expect_nothing(0.2)
    daemon.ctl("-S", "on")
    a.expect(SCHED_ON); b.expect(SCHED_ON); late.expect(SCHED_ON)
    b.send(REQ_LOCK)
    b.expect(LOCK_OK)                              # the pre-OFF holder no longer holds anything


def test_many_clients_round_robin(daemon):
    daemon.ctl("-T", "30")
    cs = [client(daemon, f"c{i}") for i in range(8)]
    for c in cs:
        c.register()
    for c in cs:
        c.send(REQ_LOCK)
    for c in cs:
        c.expect(LOCK_OK)
        c.send(LOCK_RELEASED)


def test_frames_are_exactly_537_bytes(daemon):
    c = client(daemon)
    c.send(REGISTER)
    raw = b""
    c.s.settimeout(2)
    while len(raw) < MSG_SIZE:
        raw += c.s.recv(4096)
    assert len(raw) == MSG_SIZE
    c.s.settimeout(0.3)
    with pytest.raises(Exception):
        assert c.s.recv(1) == b"impossible"


def test_drop_lock_carries_waiter_hint_ours_only(artefacts, tmp_path):
    """Wire-compatible extension: our daemon puts "w<k>" into DROP_LOCK.data."""
    sock_dir = tmp_path / "nvs"; sock_dir.mkdir()
    d = Daemon("ours", sock_dir)
    try:
        d.ctl("-T", "1")
        a, b = MockClient(d.sock_path, "a"), MockClient(d.sock_path, "b")
        ra = a.register(); b.register()
        assert ra["raw_data"][17:18] == b"2" and ra["raw_data"][16:17] == b"\0"   # capability marker after the id
        a.send(REQ_LOCK, data=b"n1234"); a.expect(LOCK_OK)
        b.send(REQ_LOCK, data=b"n777")
        assert a.expect(DROP_LOCK, timeout=3)["data"] == b"w1n777"      # one waiter, who must map 777 MiB
        a.send(LOCK_RELEASED)
        b.expect(LOCK_OK)
        # pressure from the client being granted is forwarded to everybody else as an eviction request
        b.send(REQ_LOCK, data=b"p300")
        assert a.expect(DROP_LOCK, timeout=3)["data"] == b"e300"
        b.expect_nothing(0.2)
        assert b.expect(DROP_LOCK, timeout=3)["data"] == b"w0n0"        # nobody waits: the holder may stay resident
    finally:
        d.stop()


@pytest.mark.parametrize("seed", range(4))
def test_mutual_exclusion_under_random_traffic_with_hints(artefacts, tmp_path, seed):
    """Our daemon only (the extension frames have no counterpart): random register / request (with need hints) /
    pressure / release / disconnect traffic with a 1 s quantum running underneath.  Whatever arrives, a LOCK_OK is
    only ever sent while nobody else has been told it holds the lock, pressure never grants anything, and the daemon
    survives."""
    import random
    rng = random.Random(1000 + seed)
    sock_dir = tmp_path / "nvs"; sock_dir.mkdir()
    d = Daemon("ours", sock_dir, log_path=tmp_path / "sched.log")
    clients, holder, grants = {}, None, 0
    pending = set()          # clients with a request the daemon may still answer
    try:
        d.ctl("-T", "1")

        def drain():
            nonlocal holder, grants
            time.sleep(0.02)
            for name, c in list(clients.items()):
                while True:
                    m = c.recv(0.005)
                    if m is None:
                        break
                    if m == b"":
                        c.close(); del clients[name]
                        if holder == name:
                            holder = None
                        break
                    if m["type"] == LOCK_OK:
                        assert holder is None, f"LOCK_OK to {name} while {holder} holds the lock"
                        holder = name
                        grants += 1
                        pending.discard(name)
                    elif m["type"] == DROP_LOCK and m["data"][:1] != b"e" and holder == name and rng.random() < 0.8:
                        c.send(LOCK_RELEASED)              # a well-behaved holder gives way at the end of its quantum
                        holder = None

        for step in range(150):
            op = rng.choice(["connect", "req", "req", "req", "press", "rel", "close", "wait"])
            if op == "connect" or not clients:
                if len(clients) < 5:
                    name = f"c{step}"
                    c = MockClient(d.sock_path, name)
                    c.register()
                    clients[name] = c
                drain()
                continue
            name = rng.choice(sorted(clients))
            c = clients[name]
            if op == "req":
                c.send(REQ_LOCK, data=b"n%d" % rng.randrange(0, 200000))
                pending.add(name)
            elif op == "press":
                c.send(REQ_LOCK, data=b"p%d" % rng.randrange(0, 200000))
            elif op == "rel":
                # from the holder, or from somebody who is not even asking (ignored by the daemon).  Not from a
                # waiter: its release could cross a LOCK_OK that is already on its way and give that lock back,
                # which is legal and which this model could not tell from a violation.
                if holder != name and name in pending:
                    continue
                c.send(LOCK_RELEASED)
                if holder == name:
                    holder = None
            elif op == "close":
                c.close(); del clients[name]
                pending.discard(name)
                if holder == name:
                    holder = None
            elif op == "wait":
                time.sleep(0.15)
            drain()
        assert grants >= 5
        # still alive and sane: a fresh client gets the lock once everybody else is gone
        for c in clients.values():
            c.close()
        time.sleep(0.1)
        z = MockClient(d.sock_path, "z"); z.register(); z.send(REQ_LOCK, data=b"n1"); z.expect(LOCK_OK, timeout=3)
        z.close()
    finally:
        d.stop()
