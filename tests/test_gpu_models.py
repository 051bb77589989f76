"""BASELINE configs #4/#5 as parity cases at small scale: a ResNet-50 training
loop and a Llama-architecture decode, each run (a) un-hooked and (b) as two
co-located clients under libnvshare.so with the swap forced on every hand-off
(evict-all policy, TQ = 1 s).  cuDNN / cuBLAS / attention kernels, memsets,
cluster launches and stream-ordered copies all go through the gate while their
memory is unmapped and re-mapped under them; results must agree with the
un-hooked run within north_star's 1e-5 relative."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

from nvs_testlib import ROOT, Daemon, preload

pytestmark = pytest.mark.gpu


def result_of(stdout):
    for line in stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise AssertionError("no RESULT line:\n" + stdout[-2000:])


def run_plain(kind, steps):
    r = subprocess.run([sys.executable, "-m", "nvshare_b200.workloads_models", "--kind", kind, "--steps", str(steps)],
                       cwd=ROOT, env=dict(os.environ, PYTHONPATH=str(ROOT)), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return result_of(r.stdout)


def run_hooked_pair(tmp_path, kind, steps, seconds):
    sock_dir = tmp_path / "nvs"
    sock_dir.mkdir(exist_ok=True)
    d = Daemon("ours", sock_dir, log_path=tmp_path / "sched.log")
    try:
        d.ctl("-T", "1")
        procs = []
        for i in (1, 2):
            env = dict(os.environ, LD_PRELOAD=preload("ours"), NVSHARE_SOCK_DIR=str(sock_dir), PYTHONPATH=str(ROOT),
                       NVSHARE_EVICT_POLICY="all", NVSHARE_STATS_FILE=str(tmp_path / f"stats{i}.jsonl"))
            procs.append(subprocess.Popen([sys.executable, "-m", "nvshare_b200.workloads_models", "--kind", kind,
                                           "--steps", str(steps), "--seconds", str(seconds)], cwd=ROOT, env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=1200) for p in procs]
        for p, (o, e) in zip(procs, outs):
            assert p.returncode == 0, e[-3000:]
        return d.read_log(), [result_of(o) for o, _ in outs]
    finally:
        d.stop()


def swaps(tmp_path):
    n = 0
    for i in (1, 2):
        f = tmp_path / f"stats{i}.jsonl"
        if f.exists():
            n += sum(1 for l in f.read_text().splitlines() if '"op":"evict"' in l)
    return n


def close(a, b, rel=1e-5):
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-30)


def test_resnet50_training_matches_unhooked(artefacts, tmp_path):
    pytest.importorskip("torchvision")
    plain = run_plain("resnet", 6)
    log, hooked = run_hooked_pair(tmp_path, "resnet", 6, seconds=10)
    assert log.count("Sent DROP_LOCK") >= 2 and swaps(tmp_path) >= 1          # weights, grads and momenta were swapped
    for h in hooked:
        assert len(h["losses"]) == 6
        assert all(close(x, y) for x, y in zip(h["losses"], plain["losses"])), (h["losses"], plain["losses"])
        assert close(h["param_checksum"], plain["param_checksum"])


def test_llama_decode_matches_unhooked(artefacts, tmp_path):
    pytest.importorskip("transformers")
    plain = run_plain("llama", 8)
    log, hooked = run_hooked_pair(tmp_path, "llama", 8, seconds=12)
    assert log.count("Sent DROP_LOCK") >= 2 and swaps(tmp_path) >= 1
    for h in hooked:
        assert h["tokens"] == plain["tokens"]                                  # identical greedy continuation
        assert len(h["logit_sums"]) == 8
        assert all(close(x, y) for x, y in zip(h["logit_sums"], plain["logit_sums"])), (h["logit_sums"], plain["logit_sums"])
