"""Host side of liblqft without a GPU: the library loads, exports every symbol include/lqft.h declares,
and its decision spec (RNG keying, threshold table, geometry, Wilson mask, slab plan) agrees with the
oracle's independent restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lattice_qft_b200 as L
from lattice_qft_b200 import _lib
from lattice_qft_b200 import engine as E
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lqft.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # comments mention names too
    declared = set(re.findall(r"\b(lqft_[a-z0-9_]+)\s*\(", header))
    declared -= {"lqft_handle"}
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert len(declared) >= 30


def test_version_and_error_string():
    assert _lib.lib().lqft_version() == (0 << 16) | 2
    assert isinstance(_lib.lib().lqft_last_error(), bytes)


def test_no_silent_fallback_without_device():
    if E.device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(L.LqftError) as ei:
        L.Engine(16, 16, 16)
    assert ei.value.status == _lib.LQFT_ERR_CUDA and "no CPU fallback" in ei.value.message


def test_philox_matches_oracle():
    rng = np.random.default_rng(1)
    for _ in range(200):
        seed = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        ctr = [int(v) for v in rng.integers(0, 2**32, size=4)]
        c = (C.c_uint32 * 4)(*ctr)
        o = (C.c_uint32 * 4)()
        _lib.lib().lqft_host_philox4x32(seed, c, o)
        assert list(o) == O.philox4x32([seed & 0xFFFFFFFF, seed >> 32], ctr)


def test_site_word_pick_hot_match_oracle():
    rng = np.random.default_rng(2)
    for _ in range(300):
        seed = int(rng.integers(0, 2**63))
        idx = int(rng.integers(0, 2**34))
        sweep = int(rng.integers(0, 2**40))
        colour = int(rng.integers(0, 2))
        assert E.site_word(seed, idx, colour, sweep) == O.site_word(seed, idx, colour, sweep)
        n = int(rng.integers(1, 2**31))
        assert E.pick_site(seed, 2, sweep, 3, n) == O.pick_site(seed, 2, sweep, 3, n) < n
        assert _lib.lib().lqft_host_hot_value(seed, idx) == O.hot_value(seed, idx)


def test_hot_start_is_i8_uniform():
    v = np.array([O.hot_value(5, i) for i in range(1 << 14)])
    assert v.min() >= -128 and v.max() <= 127 and abs(v.mean()) < 3 and v.std() == pytest.approx(73.9, rel=0.03)


@pytest.mark.parametrize("size", [(4, 5, 3), (16, 16, 16), (6, 2, 10)])
def test_geometry_matches_oracle(size):
    lat, olat = L.lattice.Lattice(size), O.Lattice(size)
    for index in range(0, olat.n_sites, 3):
        assert lat.values[index] == olat.neighbours(index)
        c = lat.calc_coords_from_index(index)
        assert c.into_array() == olat.coords_from_index(index)
        assert lat.calc_index_from_coords(c) == index


def test_reference_lattice_kats_through_the_abi():  # lattice.rs:192-220
    lat = L.lattice.Lattice([4, 5, 3])
    assert lat.calc_coords_from_index(29).into_array() == [1, 2, 1]
    assert lat.values[19] == [16, 3, 39, 18, 15, 59]


def test_generic_dimension_involution():  # lattice.rs:225-239 (5-d)
    lat = L.lattice.Lattice([9, 26, 5, 3, 3])
    for index in range(0, lat.n_sites, 101):
        for direction, nb in enumerate(lat.values[index]):
            assert lat.values[nb][(direction + 5) % 10] == index


@pytest.mark.parametrize("size,w,h", [((2, 2, 2), 1, 1), ((8, 4, 6), 3, 2), ((8, 2, 4), 8, 4), ((6, 6, 6), 6, 6)])
def test_wilson_offset_matches_link_table(size, w, h):
    """The coordinate predicate equals the reference's LinksField lookup (fields.rs:282-291, 821-837)."""
    olat = O.Lattice(size)
    wil = O.Wilson(olat, w, h)
    for index in range(olat.n_sites):
        x, y, t = olat.coords_from_index(index)
        want = (1 if wil.is_active(index, 1) else 0) - (1 if wil.is_active(index, 4) else 0)
        for d in (0, 2, 3, 5):
            assert not wil.is_active(index, d)
        assert _lib.lib().lqft_host_wilson_offset(size[1], x, y, t, w, h) == want


@pytest.mark.parametrize("beta", [0.2, 0.25, 0.29, 0.5, 0.9, 0.05, 0.011])
def test_threshold_table_equals_literal_rule(beta):
    """u <= thr[clamp(m+3)]  <=>  u*2^-31 <= exp((old-new)*beta), old-new = -(6+2m) (algorithm.rs:151-153)."""
    thr = E.threshold_table(beta)
    assert thr[0] == 2**31 - 1 and thr[-1] == 0 and (np.diff(thr.astype(np.int64)) <= 0).all()
    rng = np.random.default_rng(3)
    K = len(thr)
    for m in list(range(-40, K + 40)):
        k = min(max(m + 3, 0), K - 1)
        prob = np.exp(float(-(6 + 2 * m)) * beta)
        t = int(thr[k])
        us = {0, 1, 2**31 - 1, t, max(t - 1, 0), min(t + 1, 2**31 - 1)} | {int(u) for u in rng.integers(0, 2**31, 8)}
        for u in us:
            assert (u <= t) == (u * 2.0**-31 <= prob), (beta, m, u)


def test_threshold_table_rejects_bad_beta():
    for beta in (0.0, -0.1, 1e-4, float("nan")):
        with pytest.raises(ValueError):
            E.threshold_table(beta)


def test_slab_plan():
    assert E.slab(1024, 8, 3) == (384, 128)
    assert E.slab(16, 1, 0) == (0, 16)
    with pytest.raises(L.LqftError):
        E.slab(12, 4, 0)  # 3 planes per slab: odd
    with pytest.raises(L.LqftError):
        E.slab(16, 2, 2)


def test_run_measurements():
    def brute(first, n, every):
        return sum(1 for s in range(first, first + n) if every and s % every == 0)
    for first in (0, 1, 9, 10, 11, 37):
        for n in (0, 1, 5, 10, 25, 100):
            for every in (0, 1, 3, 10):
                assert E.run_measurements(first, n, every) == brute(first, n, every)


def test_rust_shim_declares_every_symbol():
    """bindings/rust/src/ffi.rs (the reference-side binding of INTEGRATION.md) names exactly the header's symbols."""
    header = open(os.path.join(ROOT, "include", "lqft.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lqft_[a-z0-9_]+)\s*\(", header))
    rust = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (lqft_[a-z0-9_]+)\s*\(", rust))
    assert bound == declared, bound ^ declared
