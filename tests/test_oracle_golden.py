"""Pins the CPU oracle (oracle/lqft_oracle.c) against every known-answer fact the reference's own
unit tests hold for this path, against the Random123 Philox vectors, and against the committed
golden vectors.  No GPU."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from conftest import seeded_field

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- src/lattice.rs:191-245 ------------------------------------------------------------------
def test_index_coordinates_conv():  # lattice.rs:192-200
    lat = O.Lattice([4, 5, 3])
    assert lat.coords_from_index(29) == [1, 2, 1]
    assert lat.index_from_coords([1, 2, 1]) == 29


def test_neighbour_index_array_example():  # lattice.rs:214-220
    lat = O.Lattice([4, 5, 3])
    assert lat.neighbours(19) == [16, 3, 39, 18, 15, 59]


def test_big_dim_neighbour_list():  # lattice.rs:225-239, on a smaller 5-d lattice of the same shape family
    size = [9, 26, 5, 3, 3]
    lat = O.Lattice(size)
    d = len(size)
    for index in range(0, lat.n_sites, 7):
        for direction, neighbour in enumerate(lat.neighbours(index)):
            assert lat.neighbours(neighbour)[(direction + d) % (2 * d)] == index


# ---- src/fields.rs:851-896 ---------------------------------------------------------------------
def test_field_shift():  # fields.rs:851-861: 3^4 lattice of zeros shifted by 3 -> values[80] == -3
    v = np.zeros(81, dtype=np.int32)
    O.shift_values(v, 3)
    assert v[80] == -3 and (v == -3).all()


def test_wilson_field_action():  # fields.rs:866-896 (intent: see SURVEY.md §4)
    lat = O.Lattice([2, 2, 2])
    wil = O.Wilson(lat, 1, 1)
    for index in range(8):
        for direction in range(6):
            assert wil.is_active(index, direction) == ((index, direction) in [(0, 1), (2, 4)])
    v = seeded_field(8, seed=3)
    for index in range(8):
        for direction in range(6):
            nb = lat.neighbours(index)[direction]
            assert wil.bond_action(v, index, direction) == wil.bond_action(v, nb, (direction + 3) % 6)


# ---- src/kahan.rs:57-75, 140-168 -----------------------------------------------------------------
def test_kahan_new():
    k = O.Kahan()
    assert (k.value, k.kahan) == (0.0, 0.0) and O.lib().orc_kahan_mean(C.byref(k)) == 0.0
    O.lib().orc_kahan_add(C.byref(k), 1.0)
    assert k.value == 1.0 and O.lib().orc_kahan_mean(C.byref(k)) == 1.0 and k.kahan == 0.0


def test_kahan_f32():  # kahan.rs:69-75
    value, kahan = C.c_float(0.0), C.c_float(0.0)
    O.lib().orc_kahan_add_f32(C.byref(value), C.byref(kahan), 3.1)
    O.lib().orc_kahan_add_f32(C.byref(value), C.byref(kahan), -3.1)
    assert value.value == 0.0


def _rust_sum(xs):
    """Iterator::sum::<f64>() adds left to right (Python's sum() compensates since 3.12)."""
    acc = 0.0
    for x in xs:
        acc += x
    return acc


def test_welfords_mean_and_sd():  # kahan.rs:140-168
    values = [1.0, 0.0, 1.0, 2.0]
    w = O.Welford()
    for v in values:
        O.lib().orc_welford_update(C.byref(w), v)
    assert O.lib().orc_welford_mean(C.byref(w)) == _rust_sum(values) / len(values)
    values = [1.0, 0.0, 1.0, 2.0, 25.0]
    mean = _rust_sum(values) / len(values)
    var = _rust_sum((x - mean) * (x - mean) for x in values) / len(values)
    w = O.Welford()
    for v in values:
        O.lib().orc_welford_update(C.byref(w), v)
    assert O.lib().orc_welford_var(C.byref(w)) == var
    assert np.sqrt(var) == np.sqrt(O.lib().orc_welford_var(C.byref(w)))  # get_sd
    # get_standard_error is var_unbiased / sqrt(n) (sic, kahan.rs:135-137)
    assert O.lib().orc_welford_standard_error(C.byref(w)) == pytest.approx(
        sum((x - mean) ** 2 for x in values) / 4 / np.sqrt(5), rel=1e-15)


# ---- Philox4x32-10 known-answer vectors (Random123 kat_vectors) -------------------------------------
@pytest.mark.parametrize("ctr,key,out", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat(ctr, key, out):
    assert O.philox4x32(key, ctr) == out


# ---- the accept/reject rule, literal (algorithm.rs:137-157) -----------------------------------------
def test_metropolis_single_rule():
    lat = O.Lattice([4, 4, 4])
    v = np.zeros(64, dtype=np.int32)
    # flat field: proposing +-1 raises the local action by 6 -> prob = exp(-6 beta)
    beta = 0.2
    prob = np.exp(-6 * beta)
    assert O.metropolis_single(lat, v, 5, beta, True, np.nextafter(prob, 0)) == 1 and v[5] == 1
    v[:] = 0
    assert O.metropolis_single(lat, v, 5, beta, True, prob) == 1          # draw <= prob is inclusive
    v[:] = 0
    assert O.metropolis_single(lat, v, 5, beta, False, np.nextafter(prob, 1)) == 0 and v[5] == 0
    # a site sitting 1 above flat neighbours: moving down lowers the action -> always accepted
    v[:] = 0
    v[5] = 1
    assert O.metropolis_single(lat, v, 5, beta, False, 1.0) == 1 and v[5] == 0


def test_local_action_matches_definition():
    lat = O.Lattice([4, 6, 8])
    v = seeded_field(lat.n_sites)
    for index in [0, 17, 100, lat.n_sites - 1]:
        want = sum((7 - int(v[n])) ** 2 for n in lat.neighbours(index))
        assert O.lib().orc_local_action(lat._h, v.ctypes.data_as(C.c_void_p), index, 7) == want


def test_energy_action_definitions():
    lat = O.Lattice([4, 6, 8])
    v = seeded_field(lat.n_sites)
    X, Y, T = lat.size
    f = v.reshape(T, Y, X).astype(np.int64)
    dx = (f - np.roll(f, -1, axis=2)) ** 2
    dy = (f - np.roll(f, -1, axis=1)) ** 2
    dt = (f - np.roll(f, -1, axis=0)) ** 2
    e_int, e_wrapped = O.integer_energy(lat, v)
    assert e_int == int((dx + dy - dt).sum()) == e_wrapped
    assert O.float_action(lat, v) == float((dx + dy + dt).sum())
    assert O.energy_observable(lat, v, 0.3) == 0.3 * float(e_int)


def test_energy_i32_wrap_quirk():  # Q5: the reference accumulates in i32
    lat = O.Lattice([8, 8, 8])
    v = (np.arange(512, dtype=np.int64) % 2 * 40000).astype(np.int32)
    e_int, e_wrapped = O.integer_energy(lat, v)
    assert e_int != e_wrapped and (e_int - e_wrapped) % (1 << 32) == 0


def test_correlator_identity():
    """outputdata.rs:313-344 equals A h_x^2 - 2 h_x S1 + S2 per slice."""
    lat = O.Lattice([4, 6, 8])
    v = seeded_field(lat.n_sites)
    X, Y, T = lat.size
    sites = [3, 77, 150, 191, 77]
    got = O.correlator(lat, v, sites)
    f = v.reshape(T, Y * X).astype(np.int64)
    s1, s2 = f.sum(axis=1), (f * f).sum(axis=1)
    want = np.zeros(T)
    for s in sites:
        tx, h = s // (X * Y), int(v[s])
        for tau in range(T):
            t = (tx + tau) % T
            want[tau] += X * Y * h * h - 2 * h * s1[t] + s2[t]
    assert (got == want).all()


def test_orders_share_uniforms_and_distribution_property():
    """Lexicographic and red-black sweeps consume the same per-site words; on a lattice where no
    two sites interact twice they differ only through the visiting order."""
    lat = O.Lattice([4, 4, 4])
    a, b = seeded_field(64), seeded_field(64)
    O.sweep(lat, a, 0.2, 1, 0, order=0)
    O.sweep(lat, b, 0.2, 1, 0, order=1)
    assert a.shape == b.shape and not (a == seeded_field(64)).all()


# ---- committed golden vectors ---------------------------------------------------------------------
def _golden():
    with open(os.path.join(GOLDEN, "sweep_golden.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", sorted(_golden()))
def test_oracle_reproduces_golden(name):
    g = _golden()[name]
    lat = O.Lattice(g["size"])
    w, h = g["wilson"]
    wil = O.Wilson(lat, w, h) if w and h else None
    v = seeded_field(lat.n_sites)
    acc = [O.sweep(lat, v, g["beta"], g["seed"], s, order=1, wilson=wil) for s in range(g["sweeps"])]
    assert acc == g["accepted"]
    assert hashlib.sha256(v.tobytes()).hexdigest() == g["field_sha256"]
    assert O.integer_energy(lat, v)[0] == g["e_int"]
    assert O.float_action(lat, v) == g["s_float"]
    assert O.correlator(lat, v, g["corr_sites"]).tolist() == g["corr"]


def test_oracle_energy_near_equipartition():
    """E -> N/6 for small beta (SURVEY.md §4 v): 16^3, beta = 0.2 -> 682.67; published 682.650 +- 0.130
    (tests/golden/reference_energies.json, index 6060).  Short CPU run, loose 5-sigma window."""
    lat = O.Lattice([16, 16, 16])
    v = O.init_hot(lat, 42)
    O.run(lat, v, 0.2, 42, 0, 0, 1000, 0, 10, order=1)
    e, _ = O.run(lat, v, 0.2, 42, 1000, 0, 6000, 10, 10, order=1)
    # per-measurement sd ~ 26, autocorrelation ~ 1 measurement at frequency 10 -> sem ~ 1.1
    assert abs(e.mean() - 682.65) < 6.0
