"""BASELINE config #1 (the reference's own CPU-runnable case) as a measurement: `harness.scheduler_pingpong`
runs tools/pingpong.c -- scripted clients doing REQ_LOCK -> LOCK_OK -> LOCK_RELEASED over the 537-byte
protocol -- against one arm's daemon with the placement of daemon and clients fixed, and `bench.py` puts the
record of its arm into the line (`configs.config1_scheduler_cpu_2_clients`).  Here: the record's shape,
that every cycle of every client was served by BOTH daemons (the behaviour itself is pinned frame by frame in
test_scheduler_protocol.py / test_scheduler_differential.py), and the wiring into the bench line.  No rates
are asserted: they belong to the box."""
from __future__ import annotations

import argparse
import importlib.util

import pytest

from nvshare_b200 import harness
from nvs_testlib import ROOT


def check(rec, impl, clients, cycles, trials):
    assert rec["impl"] == impl and rec["clients"] == clients and rec["cycles_per_client"] == cycles
    assert "same_core" in rec
    for name in ("same_core", "separate_cores"):
        if name not in rec:
            continue
        p = rec[name]
        assert len(p["handoffs_per_s_all_trials"]) == trials and min(p["handoffs_per_s_all_trials"]) > 0
        assert p["handoffs_per_s"] == sorted(p["handoffs_per_s_all_trials"])[trials // 2]
        assert 0 < p["req_to_lock_ok_us"]["p50"] <= p["req_to_lock_ok_us"]["p99"] <= p["req_to_lock_ok_us"]["max"]
        assert p["daemon_cpus"] and p["client_cpus"]
    assert rec["handoffs_per_s"] == rec["same_core"]["handoffs_per_s"]


def test_pingpong_against_our_daemon(artefacts, tmp_path):
    check(harness.scheduler_pingpong("ours", tmp_path, clients=3, cycles=500, trials=3), "ours", 3, 500, 3)
    log = (tmp_path / "scheduler_same_core.log").read_text()
    assert log.count("Sent LOCK_OK") == 3 * 3 * 500         # trials x clients x cycles: nobody was skipped


@pytest.mark.reference
def test_pingpong_against_the_reference_daemon(artefacts, default_sock_lock, tmp_path):
    check(harness.scheduler_pingpong("reference", tmp_path, clients=3, cycles=500, trials=3), "reference", 3, 500, 3)
    log = (tmp_path / "scheduler_same_core.log").read_text()
    assert log.count("Sent LOCK_OK") == 3 * 3 * 500


def test_bench_line_carries_config1_of_its_own_arm(artefacts, tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location("bench_under_test_c1", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    seen = {}

    def fake(impl, out_dir, **kw):
        seen.update(impl=impl, **kw)
        return {"impl": impl, "handoffs_per_s": 1.0}
    monkeypatch.setattr(bench.harness, "scheduler_pingpong", fake)
    for impl in ("ours", "reference"):
        line = {}
        bench.config1(argparse.Namespace(impl=impl, kind="add"), line, 1, tmp_path)
        assert line["configs"]["config1_scheduler_cpu_2_clients"]["impl"] == impl and seen["clients"] == 2
    line = {}
    bench.config1(argparse.Namespace(impl="ours", kind="add"), line, 2, tmp_path)      # N > 1: the host cores are busy
    assert line == {}
    monkeypatch.setattr(bench.harness, "scheduler_pingpong", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("boom")))
    line = {"configs": {"other": 1}}
    bench.config1(argparse.Namespace(impl="ours", kind="add"), line, 1, tmp_path)      # a failure never costs the line
    assert "boom" in line["configs"]["config1_scheduler_cpu_2_clients"]["error"] and line["configs"]["other"] == 1
