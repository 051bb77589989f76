"""Test configuration.

Markers
  gpu   needs a real B200 (the driver runs `-m gpu` on the GPU box; everything
        else must pass on a CPU-only container against oracle/fake_cuda.c).

Everything under oracle/ is test infrastructure: the unmodified reference built
into oracle/_ref (when /root/reference exists, or prebuilt on the GPU box), the
fake CUDA driver, the C restatement.  Product code never imports it.
"""
from __future__ import annotations

import fcntl
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from nvshare_b200 import build  # noqa: E402

# The cross-process GPU ledger (nvshare_b200/csrc/gpu_ledger.c) is one file per user and machine: the
# suite keeps its own, so that it neither sees nor leaves claims in the one real clients of this box use.
_LEDGER = Path(f"/dev/shm/nvshare-gpus-test-{os.getpid()}")
os.environ.setdefault("NVSHARE_GPU_LEDGER", str(_LEDGER))


def pytest_sessionfinish(session, exitstatus):
    _LEDGER.unlink(missing_ok=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 GPU")
    config.addinivalue_line("markers", "reference: needs the compiled reference in oracle/_ref")


@pytest.fixture(scope="session")
def artefacts():
    """Build (or locate) the product and the test infrastructure once."""
    prod = build.build_product()
    try:
        orc = build.build_oracle()
    except Exception:
        orc = build.ORACLE_OUT  # GPU box without gcc headers etc.: use what travelled
    return {"build": prod, "oracle": orc, "root": ROOT}


@pytest.fixture(scope="session")
def have_reference(artefacts):
    return build.reference_available()


@pytest.fixture()
def sock_dir(tmp_path):
    d = tmp_path / "nvs"
    d.mkdir()
    return d


@pytest.fixture()
def default_sock_lock():
    """The reference binaries can only use /var/run/nvshare: serialise its users."""
    lock_path = Path("/tmp/nvshare_default_sock.lock")
    f = open(lock_path, "w")
    fcntl.flock(f, fcntl.LOCK_EX)
    try:
        yield Path("/var/run/nvshare")
    finally:
        fcntl.flock(f, fcntl.LOCK_UN)
        f.close()


def pytest_collection_modifyitems(config, items):
    have_ref = build.reference_available() or Path("/root/reference/src/hook.c").exists()
    skip_ref = pytest.mark.skip(reason="compiled reference (oracle/_ref) not available")
    # `pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not errors
    have_gpu = os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")
    skip_gpu = pytest.mark.skip(reason="needs a real GPU (no /dev/nvidia* here)")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)
