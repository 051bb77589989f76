"""Property test of the swap engine's host logic (fake driver): arbitrary
interleavings of alloc / write / evict(partial or all) / fetch / free / lock-free
host copies (nvs_host_io) never lose or mix up a byte, and the accounting identities always hold.  The model is a
plain dict of numpy arrays; contents are checked by direct reads (the fake
driver's "HBM" is host-addressable while mapped)."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from nvs_testlib import FAKE_DIR

MiB = 1 << 20
SLAB = 2 * MiB


@pytest.fixture(scope="module")
def fake(artefacts):
    lib = C.CDLL(str(FAKE_DIR / "libcuda.so.1"), mode=C.RTLD_GLOBAL)
    lib.fake_cuda_phys_used.restype = C.c_uint64
    assert lib.cuInit(0) == 0
    ctx = C.c_void_p()
    assert lib.cuDevicePrimaryCtxRetain(C.byref(ctx), 0) == 0
    assert lib.cuCtxSetCurrent(ctx) == 0
    return lib


def view(ptr, nbytes):
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr))


ops = st.lists(
    st.one_of(
        st.tuples(st.just("alloc"), st.integers(1, 9), st.booleans()),          # size in slabs (+ ragged bytes), constant fill?
        st.tuples(st.just("write"), st.integers(0, 7), st.integers(0, 255)),    # which allocation, seed
        st.tuples(st.just("const"), st.integers(0, 7), st.integers(0, 255)),    # make one slab same-filled
        st.tuples(st.just("poke"), st.integers(0, 7), st.integers(0, 10**7)),   # flip ONE byte (clean-slab detection)
        st.tuples(st.just("evict"), st.integers(0, 12), st.just(0)),            # 0 = all, else MiB
        st.tuples(st.just("fetch"), st.just(0), st.just(0)),
        st.tuples(st.just("free"), st.integers(0, 7), st.just(0)),
        st.tuples(st.just("cold"), st.integers(1, 9), st.integers(0, 255)),     # allocate + load without the GPU
        st.tuples(st.just("hio_w"), st.integers(0, 7), st.integers(0, 10**6)),  # host -> swapped-out range
        st.tuples(st.just("hio_r"), st.integers(0, 7), st.integers(0, 10**6)),  # swapped-out range -> host
        st.tuples(st.just("idle"), st.integers(1, 6), st.just(0)),              # the owner computes: the pre-cleaner gets a few ms
    ),
    min_size=4, max_size=28)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(script=ops, elide=st.booleans(), chunk_slabs=st.sampled_from([1, 2, 4]), retain=st.booleans())
def test_any_interleaving_preserves_contents(fake, script, elide, chunk_slabs, retain):
    _run_model(fake, script, elide, chunk_slabs, retain=int(retain))


def _run_model(fake, script, elide, chunk_slabs, **engine_kw):
    from nvshare_b200 import engine as E
    base = fake.fake_cuda_phys_used()
    e = E.Engine(chunk_bytes=chunk_slabs * SLAB, host_arena_bytes=16 * MiB, batch_bytes=8 * MiB,
                 oom_wait_ms=200 if set(engine_kw) <= {"retain"} else 5000, elide_constant=int(elide), prepin=0, **engine_kw)
    model = {}          # ptr -> expected bytes
    order = []
    resident = True     # what the owner believes: we fetch before touching memory
    try:
        e.fetch_all()
        for op, a, b in script:
            if op == "alloc":
                size = a * SLAB - (123 if b and a > 1 else 0)
                p = e.alloc(size)
                if not resident:
                    e.fetch_all(); resident = True
                else:
                    e.fetch_all()                       # a fresh allocation is UNBACKED until mapped
                data = np.full(size, 7 if b else 0, dtype=np.uint8)
                view(p, size)[:] = data
                model[p] = data
                order.append(p)
            elif op == "poke" and order:
                p = order[a % len(order)]
                if not resident:
                    e.fetch_all(); resident = True
                at = (b * 2654435761) % len(model[p])
                model[p][at] ^= 1 + (b & 0x7f)
                view(p, len(model[p]))[at] = model[p][at]
            elif op in ("write", "const") and order:
                p = order[a % len(order)]
                if not resident:
                    e.fetch_all(); resident = True
                n = len(model[p])
                if op == "write":
                    rng = np.random.default_rng(b)
                    model[p][:] = rng.integers(0, 256, n, dtype=np.uint8)
                else:
                    s0 = (b % max(1, n // SLAB)) * SLAB
                    model[p][s0:min(n, s0 + SLAB)] = b
                view(p, n)[:] = model[p]
            elif op == "evict":
                rep = e.evict(a * MiB)
                st_ = e.stats()
                if a == 0:
                    assert st_["resident_bytes"] == 0
                resident = False
            elif op == "fetch":
                e.fetch_all(); resident = True
            elif op == "cold":
                size = a * SLAB - (77 if a > 2 else 0)
                p = e.alloc(size)
                data = np.random.default_rng(b).integers(0, 256, size, dtype=np.uint8)
                lo = (b * 4099) % size                                  # load only [lo, size): the rest must read 0
                rc = e.host_io(p + lo, data[lo:].ctypes.data, size - lo, True)
                assert rc in (0, -9) and (rc == -9 or not resident)   # after a partial eviction the engine still maps new memory
                if rc != 0:
                    e.fetch_all()
                    view(p, size)[lo:] = data[lo:]
                    view(p, size)[:lo] = 0
                data[:lo] = 0
                model[p] = data
                order.append(p)
            elif op in ("hio_w", "hio_r") and order:
                p = order[a % len(order)]
                n = len(model[p])
                lo = (b * 7919) % n
                ln = 1 + (b * 104729) % min(n - lo, 5 * MiB)
                if op == "hio_w":
                    src = np.random.default_rng(b).integers(0, 256, ln, dtype=np.uint8)
                    rc = e.host_io(p + lo, src.ctypes.data, ln, True)
                    if rc == 0:
                        model[p][lo:lo + ln] = src
                else:
                    dst = np.full(ln, 0xEE, dtype=np.uint8)
                    rc = e.host_io(p + lo, dst.ctypes.data, ln, False)
                    if rc == 0:
                        assert np.array_equal(dst, model[p][lo:lo + ln])
                assert rc in (0, -9)
                if resident:
                    assert rc == -9                                     # never served while the data is in HBM
            elif op == "idle":
                if resident:
                    e.set_resident_mode(True)                           # what a lock grant does: wakes the pre-cleaner
                time.sleep(a * 1e-3)                                    # (the next op may well find it in mid-copy)
            elif op == "free" and order:
                p = order.pop(a % len(order))
                e.free(p)
                del model[p]
            st_ = e.stats()
            va = sum((len(d) + SLAB - 1) // SLAB * SLAB for d in model.values())
            assert st_["va_bytes"] == va
            assert st_["resident_bytes"] + st_["swapped_bytes"] + st_["unbacked_bytes"] == va
            assert st_["n_allocs"] == len(model)
        e.fetch_all()
        for p, want in model.items():
            assert np.array_equal(view(p, len(want)), want)
        for p in list(model):
            e.free(p)
        st_ = e.stats()
        assert st_["host_pool_used"] == 0 and st_["va_bytes"] == 0
    finally:
        e.close()
    assert fake.fake_cuda_phys_used() == base          # every physical byte went back to the "driver"


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(script=ops, elide=st.booleans(), chunk_slabs=st.sampled_from([1, 2, 4]))
def test_any_interleaving_with_overflow_arenas(fake, tmp_path_factory, monkeypatch, script, elide, chunk_slabs):
    """The first property again, with a shared pool of 8 slabs that is far too small and no
    grace period: ordinary evictions spill into private overflow arenas, so backing units
    come from both kinds of arena (and host copies read and write both)."""
    monkeypatch.setenv("NVSHARE_POOL_GRACE_MS", "0")
    pool = tmp_path_factory.mktemp("pool") / "pool"
    _run_model(fake, script, elide, chunk_slabs, shared_pool_path=str(pool), shared_pool_bytes=16 * MiB)


ops_tight = st.lists(
    st.one_of(
        st.tuples(st.just("alloc"), st.integers(1, 6), st.booleans()),
        st.tuples(st.just("write"), st.integers(0, 7), st.integers(0, 255)),
        st.tuples(st.just("const"), st.integers(0, 7), st.integers(0, 255)),
        st.tuples(st.just("evict_be"), st.integers(0, 12), st.just(0)),         # best effort: 0 = all, else MiB
        st.tuples(st.just("fetch"), st.just(0), st.just(0)),
        st.tuples(st.just("free"), st.integers(0, 7), st.just(0)),
        st.tuples(st.just("hio_r"), st.integers(0, 7), st.integers(0, 10**6)),
    ),
    min_size=4, max_size=24)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(script=ops_tight, elide=st.booleans(), chunk_slabs=st.sampled_from([1, 2, 4]))
def test_best_effort_evictions_into_a_tight_shared_pool(fake, tmp_path_factory, script, elide, chunk_slabs):
    """Same model, but the backing store is a shared pool of only 8 slabs and every eviction
    is the best-effort flavour: it may move less than asked (never more than the pool holds),
    never blocks, and whatever it leaves behind stays resident and intact."""
    import time
    from nvshare_b200 import engine as E
    base = fake.fake_cuda_phys_used()
    pool = tmp_path_factory.mktemp("pool") / "pool"
    e = E.Engine(chunk_bytes=chunk_slabs * SLAB, host_arena_bytes=16 * MiB, batch_bytes=8 * MiB, oom_wait_ms=5000,
                 elide_constant=int(elide), prepin=0, shared_pool_path=str(pool), shared_pool_bytes=16 * MiB)
    model, order = {}, []
    try:
        e.fetch_all()
        for op, a, b in script:
            if op == "alloc":
                size = a * SLAB - (99 if b and a > 1 else 0)
                p = e.alloc(size)
                e.fetch_all()
                data = np.full(size, 5 if b else 0, dtype=np.uint8)
                view(p, size)[:] = data
                model[p] = data
                order.append(p)
            elif op in ("write", "const") and order:
                p = order[a % len(order)]
                e.fetch_all()
                n = len(model[p])
                if op == "write":
                    model[p][:] = np.random.default_rng(b).integers(0, 256, n, dtype=np.uint8)
                else:
                    s0 = (b % max(1, n // SLAB)) * SLAB
                    model[p][s0:min(n, s0 + SLAB)] = b
                view(p, n)[:] = model[p]
            elif op == "evict_be":
                t0 = time.time()
                rep = e.evict_best_effort(a * MiB)
                assert time.time() - t0 < 1.0                            # never the 2 s grace period, never oom_wait
                assert e.stats()["host_pool_used"] <= 16 * MiB
                assert rep["bytes"] <= 16 * MiB
            elif op == "fetch":
                e.fetch_all()
            elif op == "free" and order:
                p = order.pop(a % len(order))
                e.free(p)
                del model[p]
            elif op == "hio_r" and order:
                p = order[a % len(order)]
                n = len(model[p])
                lo = (b * 7919) % n
                ln = 1 + (b * 104729) % min(n - lo, 3 * MiB)
                dst = np.full(ln, 0xEE, dtype=np.uint8)
                rc = e.host_io(p + lo, dst.ctypes.data, ln, False)
                assert rc in (0, -9)
                if rc == 0:
                    assert np.array_equal(dst, model[p][lo:lo + ln])
            st_ = e.stats()
            va = sum((len(d) + SLAB - 1) // SLAB * SLAB for d in model.values())
            assert st_["va_bytes"] == va
            assert st_["resident_bytes"] + st_["swapped_bytes"] + st_["unbacked_bytes"] == va
        e.fetch_all()
        for p, want in model.items():
            assert np.array_equal(view(p, len(want)), want)
        for p in list(model):
            e.free(p)
        st_ = e.stats()
        assert st_["host_pool_used"] == 0 and st_["va_bytes"] == 0
    finally:
        e.close()
    assert fake.fake_cuda_phys_used() == base
