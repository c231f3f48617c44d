"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact post-sweep fields on
the same seeded inputs and the same per-site uniforms, exact integer observables, committed golden
vectors, and size-independent properties at BASELINE.json's full sizes."""
import hashlib
import json
import os

import numpy as np
import pytest

from lattice_qft_b200 import Engine, _lib
from lattice_qft_b200 import engine as E
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden():
    with open(os.path.join(GOLDEN, "sweep_golden.json")) as fh:
        return json.load(fh)


def oracle_sweeps(size, v, beta, seed, first, n, wilson=(0, 0)):
    lat = O.Lattice(size)
    wil = O.Wilson(lat, *wilson) if wilson[0] and wilson[1] else None
    acc = 0
    for s in range(first, first + n):
        acc += O.sweep(lat, v, beta, seed, s, order=1, wilson=wil)
    return acc


@pytest.mark.parametrize("flags", [0, _lib.FLAG_FORCE_GENERIC, _lib.FLAG_I32_VECTOR])
@pytest.mark.parametrize("name", sorted(_golden()))
def test_sweep_matches_golden_vectors(name, flags):
    g = _golden()[name]
    X, Y, T = g["size"]
    with Engine(X, Y, T, flags=flags) as eng:
        eng.set_chain(0, g["beta"], g["seed"], *g["wilson"])
        eng.upload(0, seeded_field(X * Y * T))
        acc = [eng.sweep(s, 1, want_accepted=True)[0] for s in range(g["sweeps"])]
        v = eng.download(0)
        assert acc == g["accepted"]
        assert hashlib.sha256(v.tobytes()).hexdigest() == g["field_sha256"]
        assert int(eng.energy()[0][0]) == g["e_int"]
        assert float(eng.action()[0][0]) == g["s_float"]
        assert eng.correlator(0, ref_sites=g["corr_sites"]).tolist() == g["corr"]
        assert eng.correlator(0, reps=5, sweep_index=7).tolist() == g["corr"]  # device-side site picks


@pytest.mark.parametrize("size,beta,wilson", [
    ((16, 16, 16), 0.2, (0, 0)), ((8, 8, 8), 0.29, (0, 0)), ((4, 6, 8), 0.25, (0, 0)),
    ((32, 16, 8), 0.22, (0, 0)), ((10, 12, 6), 0.4, (0, 0)), ((54, 2, 4), 0.24, (0, 0)),
    ((16, 16, 16), 0.26, (8, 16)), ((24, 6, 4), 0.21, (13, 3)), ((8, 2, 2), 0.3, (8, 2)),
    ((64, 4, 4), 0.25, (21, 4)), ((2, 2, 2), 0.2, (1, 1)),
])
def test_sweep_bit_exact_vs_oracle(size, beta, wilson):
    X, Y, T = size
    seed = 0xC0FFEE + X
    v0 = seeded_field(X * Y * T, seed=X * 1000 + Y)
    want = v0.copy()
    want_acc = oracle_sweeps(size, want, beta, seed, 3, 4, wilson)
    for flags in (0, _lib.FLAG_FORCE_GENERIC):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed, *wilson)
            eng.upload(0, v0)
            got_acc = eng.sweep(3, 4, want_accepted=True)[0]
            got = eng.download(0)
        assert (got == want).all(), f"{int((got != want).sum())} sites differ (flags={flags})"
        assert got_acc == want_acc


@pytest.mark.parametrize("size,beta,wilson", [
    ((256, 8, 4), 0.25, (0, 0)), ((256, 4, 2), 0.2, (100, 2)), ((256, 10, 6), 0.3, (0, 0)),
    ((512, 4, 4), 0.25, (0, 0)), ((512, 6, 2), 0.22, (300, 1)), ((1024, 4, 2), 0.25, (0, 0)),
    ((1024, 2, 4), 0.28, (1000, 3)), ((256, 36, 2), 0.25, (0, 0)),
])
def test_wide_row_kernels_bit_exact_vs_oracle(size, beta, wilson):
    """Wide rows: the 16-bit kernel, the i32 vector kernel and the scalar kernel all equal the oracle."""
    X, Y, T = size
    seed = 0xBEEF + X + Y
    v0 = seeded_field(X * Y * T, seed=X + 7 * Y)
    want = v0.copy()
    want_acc = oracle_sweeps(size, want, beta, seed, 10, 3, wilson)
    for flags in (0, _lib.FLAG_I32_VECTOR, _lib.FLAG_FORCE_GENERIC):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed, *wilson)
            eng.upload(0, v0)
            got_acc = eng.sweep(10, 3, want_accepted=True)[0]
            eng.sweep(13, 2)              # the non-counting instantiation
            got2 = eng.download(0)
        want2 = want.copy()
        oracle_sweeps(size, want2, beta, seed, 13, 2, wilson)
        assert got_acc == want_acc, flags
        assert (got2 == want2).all(), f"{int((got2 != want2).sum())} sites differ (flags={flags})"


@pytest.mark.parametrize("size,beta,wilson", [
    ((256, 8, 4), 0.25, (0, 0)), ((256, 4, 2), 0.2, (100, 2)), ((256, 10, 6), 0.3, (0, 0)), ((256, 36, 2), 0.25, (0, 0)),
    ((256, 2, 34), 0.22, (255, 20)), ((256, 6, 48), 0.27, (0, 0)), ((256, 2, 2), 0.05, (3, 1)),
    ((512, 4, 4), 0.25, (0, 0)), ((512, 6, 2), 0.22, (300, 1)), ((512, 2, 20), 0.24, (77, 9)), ((512, 2, 36), 0.3, (511, 36)),
    ((1024, 4, 2), 0.25, (0, 0)), ((1024, 2, 4), 0.28, (1000, 3)), ((1024, 2, 18), 0.26, (513, 7)),
])
def test_narrow_working_copy_bit_exact_vs_oracle(size, beta, wilson):
    """X = 256 / 512 / 1024: lqft_sweep / lqft_run of >= 4 sweeps work on the 16-bit copy (kernels_narrow.cuh: packed
    sums, s16x2 clamp, interleaved table; X = 256: SHFL-shared t-neighbours, wider rows: one plane per warp, x-edge
    across row segments by a predicated load).  Counting and non-counting instantiations,
    graph and eager launches, energies and normalisation inside the run — all equal to the oracle and to the
    i32 kernels bit for bit."""
    X, Y, T = size
    seed = 0xA11CE + Y * T
    lat = O.Lattice([X, Y, T])
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    v0 = seeded_field(X * Y * T, seed=Y + 31 * T)
    want = v0.copy()
    wacc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(10, 15))
    for s in range(15, 21):
        O.sweep(lat, want, beta, seed, s, order=1, wilson=wil)
    we, _ = O.run(lat, want, beta, seed, 21, 0, 24, 4, 10, order=1, wilson=wil)
    nores = _lib.FLAG_NO_RESIDENT  # lattices this small would otherwise run from shared memory
    for flags in (nores, nores | _lib.FLAG_NO_GRAPH, nores | _lib.FLAG_NO_NARROW):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed, *wilson)
            eng.upload(0, v0)
            acc = eng.sweep(10, 5, want_accepted=True)[0]
            eng.sweep(15, 6)
            e_obs, _, _ = eng.run(24, sweep_base=21, energy_every=4, normalize_every=10)
            got = eng.download(0)
            segs, plan = eng.narrow_segments(), eng.plan_info()
        assert acc == wacc, flags
        assert (e_obs[0] == we).all(), flags
        assert (got == want).all(), f"{int((got != want).sum())} sites differ (flags={flags})"
        # the 16-bit arrays are the storage from the upload on: one conversion, no i32 arrays at all
        if flags & _lib.FLAG_NO_NARROW:
            assert segs == 0 and plan["storage"] == "i32", (flags, segs, plan)
        else:
            assert segs == 1 and plan["storage"] == "u16" and plan["i32_allocated"] == 0, (flags, segs, plan)


def test_narrow_budget_recentring_and_guard_band():
    """The 16-bit arrays store the field between calls.  A conversion that measured `worst` is good for 2047 - worst
    sweeps; then the heights are measured against the chain's current offset and re-centred in place.  A field outside
    the guard band lives in the i32 arrays.  Same trajectory either way."""
    X, Y, T = 256, 2, 2
    beta, seed = 0.3, 77
    lat = O.Lattice([X, Y, T])
    v0 = seeded_field(X * Y * T, seed=5)
    want = v0.copy()
    we, _ = O.run(lat, want, beta, seed, 0, 0, 2500, 100, 10, order=1)
    with Engine(X, Y, T, flags=_lib.FLAG_NO_RESIDENT) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, v0)
        budget0 = eng.plan_info()["budget"]
        assert 1900 <= budget0 <= 2047
        e_obs, _, _ = eng.run(2500, energy_every=100, normalize_every=10)
        got = eng.download(0)
        plan = eng.plan_info()
        assert eng.narrow_segments() == 2 and plan["storage"] == "u16" and plan["i32_allocated"] == 0, plan
        assert plan["budget"] > 1024
    assert (e_obs[0] == we).all() and (got == want).all()
    far = v0.copy()
    far[::7] += 5000
    want = far.copy()
    for s in range(6):
        O.sweep(lat, want, beta, seed, s, order=1)
    with Engine(X, Y, T, flags=_lib.FLAG_NO_RESIDENT) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, far)
        eng.sweep(0, 6)
        got = eng.download(0)
        plan = eng.plan_info()
        assert eng.narrow_segments() == 0 and plan["storage"] == "i32" and plan["i32_allocated"] == 1, plan
    assert (got == want).all()


def test_leaving_the_band_mid_call_falls_back_to_i32_bit_exactly():
    """The band is measured around the chain's normalisation offset.  A shift of 5000 moves the offset away from the
    heights; the running sweep budget stays valid (nothing was rewritten), but when it runs out in the middle of the
    next call the re-measurement fails and the call finishes on the i32 arrays — same results as the oracle.  After
    the shift is taken back, the retry interval (1000 sweeps) brings the field into the 16-bit arrays again."""
    X, Y, T = 256, 4, 8
    beta, seed = 0.3, 4242
    lat = O.Lattice([X, Y, T])
    v0 = seeded_field(X * Y * T, seed=9, lo=-2, hi=3)
    want = v0.copy()
    with Engine(X, Y, T, flags=_lib.FLAG_NO_RESIDENT) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, v0)
        budget = eng.plan_info()["budget"]
        assert eng.plan_info()["storage"] == "u16" and 2040 <= budget <= 2047
        eng.shift_values(0, 5000)
        O.shift_values(want, 5000)
        e_obs, _, _ = eng.run(2200, energy_every=100, normalize_every=0)   # a normalize would pull the offset back
        we, _ = O.run(lat, want, beta, seed, 0, 0, 2200, 100, 0, order=1)
        got = eng.download(0)
        plan = eng.plan_info()
        assert plan["storage"] == "i32" and plan["i32_allocated"] == 1, plan
        assert (e_obs[0] == we).all() and (got == want).all()
        eng.shift_values(0, -5000)
        O.shift_values(want, -5000)
        eng.sweep(2200, 1100)
        for s in range(2200, 3300):
            O.sweep(lat, want, beta, seed, s, order=1)
        assert eng.plan_info()["storage"] == "u16"
        assert (eng.download(0) == want).all()


def test_narrow_many_chains_match_single_chain_runs():
    X, Y, T = 256, 4, 4
    betas = [0.2, 0.25, 0.31]
    v0 = [seeded_field(X * Y * T, seed=40 + c) for c in range(3)]
    single = []
    for c in range(3):
        with Engine(X, Y, T, flags=_lib.FLAG_NO_NARROW | _lib.FLAG_NO_RESIDENT) as eng:
            eng.set_chain(0, betas[c], 900 + c, 17 * c, c)
            eng.upload(0, v0[c])
            e, _, _ = eng.run(20, energy_every=5, normalize_every=10)
            single.append((e[0].copy(), eng.download(0)))
    with Engine(X, Y, T, n_chains=3, flags=_lib.FLAG_NO_RESIDENT) as eng:
        for c in range(3):
            eng.set_chain(c, betas[c], 900 + c, 17 * c, c)
            eng.upload(c, v0[c])
        e, _, _ = eng.run(20, energy_every=5, normalize_every=10)
        assert eng.narrow_segments() == 3 and eng.plan_info()["storage"] == "u16"   # one conversion per uploaded chain
        for c in range(3):
            assert (e[c] == single[c][0]).all()
            assert (eng.download(c) == single[c][1]).all()


@pytest.mark.parametrize("size,wilson", [((256, 8, 8), (0, 0)), ((256, 6, 4), (77, 3)), ((512, 4, 6), (0, 0)),
                                         ((16, 16, 8), (5, 8)), ((10, 6, 4), (0, 0)), ((256, 4, 40), (200, 33)),
                                         ((256, 2, 2), (0, 0)), ((512, 2, 40), (400, 30)), ((1024, 2, 24), (0, 0))])
def test_self_halo_slab_path_on_one_gpu(size, wilson):
    """LQFT_FLAG_SELF_HALO: ghost planes + the multi-GPU code path (16-bit storage: ghost sets, pushes, flips and
    ghost tasks of k_sweep_n16; other extents: boundary planes first + exchange) with this rank as its own
    neighbour.  Must equal the oracle bit for bit: sweeps, fused schedule, observables, normalize."""
    X, Y, T = size
    beta, seed = 0.24, 4321
    lat = O.Lattice([X, Y, T])
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    v0 = seeded_field(X * Y * T, seed=3 * X + T)
    want = v0.copy()
    wacc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(4))
    we, _ = O.run(lat, want, beta, seed, 4, 0, 12, 4, 10, order=1, wilson=wil)
    with Engine(X, Y, T, flags=_lib.FLAG_SELF_HALO) as eng:
        eng.set_chain(0, beta, seed, *wilson)
        eng.upload(0, v0)
        acc = eng.sweep(0, 4, want_accepted=True)[0]
        e_obs, _, _ = eng.run(12, sweep_base=4, energy_every=4, normalize_every=10)
        got = eng.download(0)
        s1, s2 = eng.slice_moments(0)
        e_int = int(eng.energy()[0][0])
        s_int = int(eng.action()[0][0])
        eng.synchronize()
        # slabs: the upload goes to the i32 arrays, the first sweep call takes the field into the 16-bit ones for good
        assert eng.narrow_segments() == (1 if X in (16, 32, 64, 128, 256, 512, 1024) else 0)
        assert eng.plan_info()["storage"] == ("u16" if X in (16, 32, 64, 128, 256, 512, 1024) else "i32")
        # a schedule short-cycled enough for graph replay (on slabs: narrow copy only, exchanges inside the graph)
        want3 = want.copy()
        we3, _ = O.run(lat, want3, beta, seed, 16, 0, 40, 2, 4, order=1, wilson=wil)
        e3, _, _ = eng.run(40, sweep_base=16, energy_every=2, normalize_every=4)
        got3 = eng.download(0)
        e_int3 = int(eng.energy()[0][0])
    assert (e3[0] == we3).all()
    assert (got3 == want3).all(), f"{int((got3 != want3).sum())} sites differ after the graph-replayed run"
    assert e_int3 == O.integer_energy(lat, want3)[0]
    assert acc == wacc
    assert (e_obs[0] == we).all()
    assert (got == want).all(), f"{int((got != want).sum())} sites differ"
    f = want.reshape(T, Y * X).astype(np.int64)
    assert (s1 == f.sum(axis=1)).all() and (s2 == (f * f).sum(axis=1)).all()
    assert e_int == O.integer_energy(lat, want)[0] and float(s_int) == O.float_action(lat, want)


@pytest.mark.parametrize("depth", [2, 4, 6])
@pytest.mark.parametrize("size", [(256, 4, 40), (512, 2, 24), (1024, 2, 20)])
def test_self_halo_other_ghost_depths(size, depth, monkeypatch):
    """Large planes get a shallower deep halo (lqft.cu picks 8 / 4 / 2 by plane size); the test lattices are small,
    so the depth is forced through the tuning knob.  Interior-under-exchange + wedges at every depth."""
    X, Y, T = size
    monkeypatch.setenv("LQFT_N16_DEPTH", str(depth))
    beta, seed, wilson = 0.26, 99 + depth, (X // 3, T - 3)
    lat = O.Lattice([X, Y, T])
    wil = O.Wilson(lat, *wilson)
    v0 = seeded_field(X * Y * T, seed=depth + T)
    want = v0.copy()
    wacc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(7))
    we, _ = O.run(lat, want, beta, seed, 7, 0, 30, 3, 10, order=1, wilson=wil)
    with Engine(X, Y, T, flags=_lib.FLAG_SELF_HALO) as eng:
        eng.set_chain(0, beta, seed, *wilson)
        eng.upload(0, v0)
        acc = eng.sweep(0, 7, want_accepted=True)[0]
        e_obs, _, _ = eng.run(30, sweep_base=7, energy_every=3, normalize_every=10)
        got = eng.download(0)
        assert eng.narrow_segments() == 1 and eng.plan_info()["storage"] == "u16"
        eng.synchronize()
    assert acc == wacc and (e_obs[0] == we).all()
    assert (got == want).all(), f"{int((got != want).sum())} sites differ (depth {depth})"


def test_self_halo_many_chains_deep_halo():
    """Three chains with different couplings and Wilson sheets on a self-neighbouring slab (narrow copy, deep halo,
    exchange per chain) equal three single-chain runs on the plain periodic lattice."""
    X, Y, T = 256, 4, 32
    betas = [0.21, 0.25, 0.3]
    v0 = [seeded_field(X * Y * T, seed=70 + c) for c in range(3)]
    single = []
    for c in range(3):
        with Engine(X, Y, T, flags=_lib.FLAG_NO_NARROW | _lib.FLAG_NO_RESIDENT) as eng:
            eng.set_chain(0, betas[c], 300 + c, 40 * c, 5 * c)
            eng.upload(0, v0[c])
            acc = eng.sweep(0, 6, want_accepted=True)[0]
            e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
            single.append((acc, e[0].copy(), eng.download(0)))
    with Engine(X, Y, T, n_chains=3, flags=_lib.FLAG_SELF_HALO) as eng:
        for c in range(3):
            eng.set_chain(c, betas[c], 300 + c, 40 * c, 5 * c)
            eng.upload(c, v0[c])
        acc = eng.sweep(0, 6, want_accepted=True)
        e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
        assert eng.narrow_segments() == 1 and eng.plan_info()["storage"] == "u16"
        for c in range(3):
            assert acc[c] == single[c][0]
            assert (e[c] == single[c][1]).all()
            assert (eng.download(c) == single[c][2]).all()
        eng.synchronize()


def test_hot_start_matches_oracle():
    with Engine(16, 8, 4) as eng:
        eng.set_chain(0, 0.2, 99)
        eng.init_hot()
        got = eng.download(0)
    want = O.init_hot(O.Lattice([16, 8, 4]), 99)
    assert (got == want).all() and got.min() >= -128 and got.max() <= 127


def test_upload_download_roundtrip_and_edge_values():
    v = seeded_field(8 * 6 * 4, lo=-2**20, hi=2**20)
    with Engine(8, 6, 4) as eng:
        eng.set_chain(0, 0.2, 1)
        eng.upload(0, v)
        assert (eng.download(0) == v).all()


def test_batched_chains_are_independent():
    """64 chains in one handle (config 2 layout) == 64 single-chain oracles."""
    X = Y = T = 8
    n_chains = 6
    betas = [0.20 + 0.005 * c for c in range(n_chains)]
    seeds = [0xC0FFEE + c for c in range(n_chains)]
    with Engine(X, Y, T, n_chains=n_chains) as eng:
        for c in range(n_chains):
            eng.set_chain(c, betas[c], seeds[c], 3 if c == 4 else 0, 4 if c == 4 else 0)
        eng.init_hot()
        acc = eng.sweep(0, 3, want_accepted=True)
        e_int, e_obs = eng.energy()
        s_int, _ = eng.action()
        for c in range(n_chains):
            lat = O.Lattice([X, Y, T])
            want = O.init_hot(lat, seeds[c])
            wacc = oracle_sweeps((X, Y, T), want, betas[c], seeds[c], 0, 3, (3, 4) if c == 4 else (0, 0))
            assert (eng.download(c) == want).all(), c
            assert acc[c] == wacc
            assert int(e_int[c]) == O.integer_energy(lat, want)[0]
            assert e_obs[c] == O.energy_observable(lat, want, betas[c])
            assert float(s_int[c]) == O.float_action(lat, want)


def test_observables_exact():
    X, Y, T = 12, 10, 8
    lat = O.Lattice([X, Y, T])
    v = seeded_field(lat.n_sites, seed=5)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.27, 1)
        eng.upload(0, v)
        e_int, e_obs = eng.energy()
        s_int, s_obs = eng.action()
        s1, s2 = eng.slice_moments(0)
    assert int(e_int[0]) == O.integer_energy(lat, v)[0]
    assert e_obs[0] == O.energy_observable(lat, v, 0.27)
    assert float(s_int[0]) == O.float_action(lat, v) and s_obs[0] == O.action_observable(lat, v, 0.27)
    f = v.reshape(T, Y * X).astype(np.int64)
    assert (s1 == f.sum(axis=1)).all() and (s2 == (f * f).sum(axis=1)).all()


def test_energy_int64_superset_of_i32_quirk():
    """Q5: the reference's i32 accumulator holds the low 32 bits of the exact sum."""
    X = Y = T = 8
    v = (np.arange(512, dtype=np.int64) % 2 * 40000).astype(np.int32)
    lat = O.Lattice([X, Y, T])
    exact, wrapped = O.integer_energy(lat, v)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.2, 1)
        eng.upload(0, v)
        e = int(eng.energy()[0][0])
    assert e == exact and np.int64(e).astype(np.int32) == wrapped != exact


def test_correlator_matches_reference_loop():
    X, Y, T = 8, 6, 10
    lat = O.Lattice([X, Y, T])
    v = seeded_field(lat.n_sites, seed=8)
    sites = [int(s) for s in np.random.default_rng(4).integers(0, lat.n_sites, 100)]
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.25, 321)
        eng.upload(0, v)
        got = eng.correlator(0, ref_sites=sites)
        got_auto = eng.correlator(0, reps=100, sweep_index=40)
    assert (got == O.correlator(lat, v, sites)).all()
    auto_sites = [O.pick_site(321, 4, 40, r, lat.n_sites) for r in range(100)]
    assert (got_auto == O.correlator(lat, v, auto_sites)).all()


def test_normalize_and_shift():
    X, Y, T = 8, 8, 4
    lat = O.Lattice([X, Y, T])
    v = seeded_field(lat.n_sites, seed=6)
    with Engine(X, Y, T, n_chains=2) as eng:
        eng.set_chain(0, 0.2, 10)
        eng.set_chain(1, 0.2, 11)
        eng.upload(0, v)
        eng.upload(1, v)
        eng.normalize(17)
        for c, seed in ((0, 10), (1, 11)):
            site = O.pick_site(seed, 2, 17, 0, lat.n_sites)
            want = v.copy()
            O.shift_values(want, int(v[site]))
            got = eng.download(c)
            assert (got == want).all() and got[site] == 0
        eng.upload(0, v)
        eng.normalize_site(0, 5)
        assert (eng.download(0) == v - v[5]).all() and (eng.download(1) == v - v[O.pick_site(11, 2, 17, 0, lat.n_sites)]).all()
        eng.upload(0, v)
        eng.shift_values(0, 3)  # fields.rs:851-861
        assert (eng.download(0) == v - 3).all()


@pytest.mark.parametrize("flags", [0, _lib.FLAG_NO_RESIDENT, _lib.FLAG_NO_RESIDENT | _lib.FLAG_NO_GRAPH])
@pytest.mark.parametrize("first_step,n_steps,energy_every", [(0, 50, 10), (3, 44, 10), (0, 25, 1), (7, 30, 4), (0, 12, 0)])
def test_run_schedule_matches_oracle(first_step, n_steps, energy_every, flags):
    """lqft_run == loop body of Simulation::run (computation.rs:249-270), graph replay included."""
    X, Y, T = 8, 8, 8
    lat = O.Lattice([X, Y, T])
    beta, seed, base = 0.23, 555, 1000
    want = O.init_hot(lat, seed)
    we, wacc = O.run(lat, want, beta, seed, base, first_step, n_steps, energy_every, 10, order=1)
    with Engine(X, Y, T, flags=flags) as eng:
        eng.set_chain(0, beta, seed)
        eng.init_hot()
        e_obs, e_int, acc = eng.run(n_steps, first_step=first_step, sweep_base=base, energy_every=energy_every,
                                    normalize_every=10, want_accepted=True)
        got = eng.download(0)
    assert (got == want).all()
    assert e_obs.shape == (1, len(we)) and (e_obs[0] == we).all()
    assert acc[0] == wacc


@pytest.mark.parametrize("size,wilson", [((16, 16, 16), (5, 16)), ((24, 24, 24), (0, 0)), ((36, 36, 36), (0, 0)),
                                         ((6, 10, 4), (3, 2)), ((12, 6, 10), (0, 0))])
def test_resident_kernel_whole_run_in_one_launch(size, wilson):
    """Lattices that fit one SM's shared memory (the reference's 16^3 ... 36^3) run their whole schedule in ONE
    launch; batch of chains, Wilson chain included, identical to the multi-launch path and to the oracle."""
    X, Y, T = size
    n_chains = 3
    betas, seeds = [0.2, 0.25, 0.29], [11, 12, 13]
    steps = 12 if X * Y * T > 20000 else 40
    res = {}
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        with Engine(X, Y, T, n_chains=n_chains, flags=flags) as eng:
            for c in range(n_chains):
                eng.set_chain(c, betas[c], seeds[c], *(wilson if c == 1 else (0, 0)))
            eng.init_hot()
            before = eng.kernel_launches()
            e_obs, e_int, acc = eng.run(steps, first_step=0, sweep_base=5, energy_every=4, normalize_every=10,
                                        want_accepted=True)
            launches = eng.kernel_launches() - before
            res[flags] = (e_obs, acc, [eng.download(c) for c in range(n_chains)], launches, eng.energy()[0])
    assert res[0][3] == 1 and res[_lib.FLAG_NO_RESIDENT][3] > steps
    assert (res[0][0] == res[_lib.FLAG_NO_RESIDENT][0]).all() and res[0][1] == res[_lib.FLAG_NO_RESIDENT][1]
    assert (res[0][4] == res[_lib.FLAG_NO_RESIDENT][4]).all()
    for c in range(n_chains):
        assert (res[0][2][c] == res[_lib.FLAG_NO_RESIDENT][2][c]).all()
    # and the oracle for one of the chains
    lat = O.Lattice([X, Y, T])
    want = O.init_hot(lat, seeds[1])
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    we, wacc = O.run(lat, want, betas[1], seeds[1], 5, 0, steps, 4, 10, order=1, wilson=wil)
    assert (res[0][0][1] == we).all() and res[0][1][1] == wacc and (res[0][2][1] == want).all()


def test_run_wilson_chain_measures_plain_field():
    """WilsonSim: Wilson-shifted sampling, energy of the plain field every step (quirks Q3, Q4)."""
    X, Y, T = 8, 8, 8
    lat = O.Lattice([X, Y, T])
    wil = O.Wilson(lat, 4, 8)
    want = O.init_hot(lat, 9)
    we, _ = O.run(lat, want, 0.26, 9, 0, 0, 30, 1, 10, order=1, wilson=wil)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.26, 9, 4, 8)
        eng.init_hot()
        e_obs, _, _ = eng.run(30, energy_every=1)
        assert (eng.download(0) == want).all()
    assert (e_obs[0] == we).all()


def test_difference_plot_matches_welford():
    X, Y, T = 6, 4, 4
    lat = O.Lattice([X, Y, T])
    v = seeded_field(lat.n_sites, seed=12)
    wel = O.new_welford_cells(3 * lat.n_sites)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.2, 4)
        eng.upload(0, v)
        for s in range(7):
            eng.sweep(s, 1)
            cur = eng.download(0)
            eng.difference_update(0)
            O.difference_update(lat, cur, wel)
        mean, count = eng.difference_read(0)
    assert count == 7
    assert (mean.reshape(-1) == O.difference_means(lat, wel)).all()


def test_generic_and_vector_kernels_agree_at_64cubed():
    X = Y = T = 64
    with Engine(X, Y, T) as a, Engine(X, Y, T, flags=_lib.FLAG_FORCE_GENERIC) as b, \
            Engine(X, Y, T, flags=_lib.FLAG_I32_VECTOR) as c:
        for e in (a, b, c):
            e.set_chain(0, 0.25, 2024, 21, 32)
            e.init_hot()
            e.sweep(0, 6)
        va, vb, vc = a.download(0), b.download(0), c.download(0)
        assert (va == vb).all() and (va == vc).all()
        assert a.energy()[0][0] == b.energy()[0][0]


def test_oracle_parity_at_64cubed_one_sweep():
    X = Y = T = 64
    lat = O.Lattice([X, Y, T])
    want = O.init_hot(lat, 31337)
    acc_w = O.sweep(lat, want, 0.25, 31337, 0, order=1)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, 0.25, 31337)
        eng.init_hot()
        acc = eng.sweep(0, 1, want_accepted=True)[0]
        assert (eng.download(0) == want).all() and acc == acc_w


# ---- full-size, size-independent properties -------------------------------------------------------
def test_256cubed_properties():
    X = Y = T = 256
    n = X * Y * T
    with Engine(X, Y, T) as eng, Engine(X, Y, T, flags=_lib.FLAG_FORCE_GENERIC) as ref:
        for e in (eng, ref):
            e.set_chain(0, 0.25, 7)
            e.init_hot()
        v0 = eng.download(0)
        assert v0.min() == -128 and v0.max() == 127
        # round trip of the layout conversion at full size
        eng.upload(0, v0)
        assert (eng.download(0) == v0).all()
        acc = eng.sweep(0, 5, want_accepted=True)[0]
        assert eng.narrow_segments() > 0              # >= 4 sweeps per call: the 16-bit kernel bench.py times
        ref.sweep(0, 5)
        v = eng.download(0)
        assert (v == ref.download(0)).all()           # two independent kernels agree bit for bit
        assert 0 < acc <= 5 * n
        assert np.abs(v.astype(np.int64) - v0).max() <= 5  # +-1 per sweep
        # shift invariance of every observable (the action only sees differences)
        e0, s0 = int(eng.energy()[0][0]), int(eng.action()[0][0])
        s1, s2 = eng.slice_moments(0)
        assert s1.sum() == int(v.astype(np.int64).sum()) and s2.sum() == int((v.astype(np.int64) ** 2).sum())
        eng.normalize(0)
        assert int(eng.energy()[0][0]) == e0 and int(eng.action()[0][0]) == s0
        # RNG keyed by (seed, sweep, site): same sweep index twice from the same state == same result
        eng.upload(0, v0)
        eng.sweep(0, 5)
        assert (eng.download(0) == v).all()


@pytest.mark.parametrize("size,wilson", [((8, 6, 4), (0, 0)), ((16, 16, 16), (5, 16)), ((6, 10, 4), (3, 2))])
def test_f64_fields_within_1e12(size, wilson):
    """Field<f64> (fields.rs:573-599): post-sweep field and observables vs the f64 oracle.  The north-star
    tolerance is 1e-12 relative; the field itself comes out identical (heights stay on start + integers)."""
    X, Y, T = size
    beta, seed = 0.23, 808
    lat = O.Lattice([X, Y, T])
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    rng = np.random.default_rng(17)
    v0 = rng.normal(0.0, 3.0, size=X * Y * T)          # genuinely real-valued heights
    want = v0.copy()
    wacc = sum(O.sweep_f64(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(5))
    sites = [int(v) for v in rng.integers(0, X * Y * T, 20)]
    with Engine(X, Y, T, dtype="f64") as eng:
        eng.set_chain(0, beta, seed, *wilson)
        eng.upload(0, v0)
        assert (eng.download(0) == v0).all()
        acc = eng.sweep(0, 5, want_accepted=True)[0]
        got = eng.download(0)
        _, e_obs = eng.energy()
        _, s_obs = eng.action()
        corr = eng.correlator(0, ref_sites=sites)
        e_run, _, _ = eng.run(10, sweep_base=5, energy_every=5, normalize_every=10)
        got2 = eng.download(0)
    assert acc == wacc
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    assert e_obs[0] == pytest.approx(beta * O.energy_f64(lat, want), rel=1e-12)
    assert s_obs[0] == pytest.approx(beta * O.action_f64(lat, want), rel=1e-12)
    np.testing.assert_allclose(corr, O.correlator_f64(lat, want, sites), rtol=1e-11)
    # schedule: sweeps, energy at steps 0 and 5, normalize at step 0 (shift by the picked site's height)
    energies = []
    for step in range(10):
        O.sweep_f64(lat, want, beta, seed, 5 + step, order=1, wilson=wil)
        if step % 5 == 0:
            energies.append(beta * O.energy_f64(lat, want))
        if step % 10 == 0:
            want -= want[O.pick_site(seed, 2, 5 + step, 0, X * Y * T)]
    np.testing.assert_allclose(e_run[0], energies, rtol=1e-11)
    np.testing.assert_allclose(got2, want, rtol=1e-12, atol=1e-12)


def test_error_behaviour():
    from lattice_qft_b200 import LqftError
    with pytest.raises(LqftError) as ei:
        Engine(15, 16, 16)
    assert ei.value.status == _lib.LQFT_ERR_INVALID
    with Engine(4, 4, 4) as eng:
        with pytest.raises(LqftError) as ei:
            eng.sweep(0, 1)  # chain not configured
        assert ei.value.status == _lib.LQFT_ERR_STATE
        with pytest.raises(LqftError):
            eng.set_chain(0, -1.0, 0)
        with pytest.raises(LqftError):
            eng.set_chain(1, 0.2, 0)
        eng.set_chain(0, 0.2, 0)
        with pytest.raises(LqftError):
            eng.normalize_site(0, 64)
