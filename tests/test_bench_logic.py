"""bench.py's host-side decisions, CPU only: geometry of the BASELINE
configuration and the host-memory-aware scale selection."""
from __future__ import annotations

import importlib.util
import math
import types

from nvs_testlib import ROOT


def load_bench(shm_bytes=1 << 40):
    """bench.py as a module, with /dev/shm reporting `shm_bytes` of free space."""
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    real_os = m.os

    class FakeOS:
        def __getattr__(self, name):
            return getattr(real_os, name)

        @staticmethod
        def statvfs(path):
            return types.SimpleNamespace(f_bavail=shm_bytes // 4096, f_frsize=4096)

    m.os = FakeOS()
    return m


def args(**kw):
    d = dict(hbm_fraction=0.0, impl="ours", oversub=1.5, clients=2)
    d.update(kw)
    return types.SimpleNamespace(**d)


HBM = 191_503_007_744          # what cuMemGetInfo reports on the r01 B200


def test_full_scale_fits_for_ours_but_not_for_the_reference(monkeypatch):
    b = load_bench()
    monkeypatch.setattr(b, "host_memory_budget", lambda: 213_000_000_000)     # the r01 box: 200 GiB cgroup
    assert b.pick_fraction(args(impl="ours"), HBM) == (1.0, None)
    frac, note = b.pick_fraction(args(impl="reference"), HBM)
    assert 0.5 <= frac <= 0.65 and "host RAM budget" in note
    # an explicit fraction is never overridden
    assert b.pick_fraction(args(impl="reference", hbm_fraction=0.25), HBM) == (0.25, None)


def test_no_limit_means_full_scale(monkeypatch):
    b = load_bench()
    monkeypatch.setattr(b, "host_memory_budget", lambda: None)
    assert b.pick_fraction(args(impl="reference"), HBM) == (1.0, None)
    monkeypatch.setattr(b, "host_memory_budget", lambda: 2_000_000_000_000)
    assert b.pick_fraction(args(impl="reference"), HBM) == (1.0, None)


def test_host_memory_model():
    b = load_bench()
    f = 0.75 * HBM
    # the reference keeps every client's pages on the host; ours only what is swapped out (+ pinned windows ahead)
    assert b.host_memory_needed("reference", 2, f, HBM) > 2 * f
    assert b.host_memory_needed("ours", 2, f, HBM) < 1.0 * HBM
    assert b.host_memory_needed("ours", 2, 0.4 * HBM, HBM) < 20e9          # fits: nothing to swap


def test_small_dev_shm_means_private_pools(monkeypatch):
    # docker's default 64 MiB /dev/shm: every client pins its own arenas, which no longer fits at full scale
    b = load_bench(shm_bytes=64 << 20)
    f = 0.75 * HBM
    assert b.host_memory_needed("ours", 2, f, HBM) > 2 * (2 * f - HBM)
    monkeypatch.setattr(b, "host_memory_budget", lambda: 213_000_000_000)
    frac, note = b.pick_fraction(args(impl="ours"), HBM)
    assert frac < 1.0 and note


def test_baseline_geometry():
    # four live n^2 fp32 blocks per client (x, y, z and the allocator's cached block): SURVEY 8d
    n = int(math.floor(math.sqrt(1.5 * HBM / 2 / 16)))
    assert n == 94745
    assert abs(4 * 4 * n * n - 0.75 * HBM) / (0.75 * HBM) < 1e-4
