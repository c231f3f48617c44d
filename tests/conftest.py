import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _ensure_built():
    import __graft_entry__
    _build = __graft_entry__.load_build_module()
    if _build.stale():
        _build.build()
    from oracle import oracle
    oracle.build()


_ensure_built()


def gpu_available() -> bool:
    from lattice_qft_b200.engine import device_count
    return device_count() > 0


@pytest.fixture(scope="session")
def have_gpu():
    return gpu_available()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail (not skip) on a box without a device only if explicitly selected;
    # under a plain `pytest` run on a CPU box they are skipped.
    if gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    selected = config.getoption("-m") or ""
    for item in items:
        if "gpu" in item.keywords and "gpu" not in selected.replace("not gpu", ""):
            item.add_marker(skip)


HARNESS_SEED = 20261019  # seed of the uniform-fed parity tests


def seeded_field(n_sites: int, seed: int = HARNESS_SEED, lo: int = -128, hi: int = 128) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi, size=n_sites, dtype=np.int32)
