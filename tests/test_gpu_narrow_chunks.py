"""Parity of k_sweep_n16 (kernels_narrow.cuh) on y-chunks LONGER than two rows: the third leg of the rotating
register window, the loop-back, odd chunk starts and uneven chunk lengths — the shapes the throughput numbers are
measured on (256^3: 37 chunks of 6-7 rows; 512^3 and 1024^2 x 128: 56-57 rows).  The reference logic these check is
metropolis_single (algorithm.rs:127-158) through oracle/lqft_oracle.c (lqft_oracle_sweep)."""
import numpy as np
import pytest

from lattice_qft_b200 import Engine, _lib
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu
NORES = _lib.FLAG_NO_RESIDENT


def _oracle_trajectory(size, v0, beta, seed, wilson):
    """5 counted sweeps, 4 plain ones, then a 12-step run (energy every 3rd, normalize every 10th)."""
    lat = O.Lattice(list(size))
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    want = v0.copy()
    acc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(5))
    for s in range(5, 9):
        O.sweep(lat, want, beta, seed, s, order=1, wilson=wil)
    e, _ = O.run(lat, want, beta, seed, 9, 0, 12, 3, 10, order=1, wilson=wil)
    return want, acc, e


def _engine_trajectory(eng, v0, beta, seed, wilson):
    eng.set_chain(0, beta, seed, *wilson)
    eng.upload(0, v0)
    acc = eng.sweep(0, 5, want_accepted=True)[0]
    eng.sweep(5, 4)
    e, _, _ = eng.run(12, sweep_base=9, energy_every=3, normalize_every=10)
    return eng.download(0), acc, e[0]


@pytest.mark.parametrize("chunks", [1, 2, 3])
@pytest.mark.parametrize("size,wilson", [
    ((256, 14, 4), (0, 0)), ((256, 10, 6), (100, 5)), ((256, 22, 2), (0, 0)), ((256, 10, 34), (255, 20)),
    ((512, 14, 4), (300, 3)), ((512, 10, 2), (0, 0)), ((1024, 10, 2), (0, 0)), ((1024, 14, 4), (513, 3)),
])
def test_narrow_rows_per_chunk_ge_3_vs_oracle(size, wilson, chunks, monkeypatch):
    """LQFT_N16_CHUNKS forces 1, 2 or 3 y-chunks: rows per chunk 3 ... 22, chunk starts at odd y (3 chunks of 14 rows
    start at 0, 4, 9), lengths that differ between chunks.  Counting / Wilson / plain instantiations, eager and
    graph-replayed launches; all bit-equal to the oracle."""
    monkeypatch.setenv("LQFT_N16_CHUNKS", str(chunks))
    X, Y, T = size
    beta, seed = 0.23 + 0.01 * chunks, 0xD1CE + Y
    v0 = seeded_field(X * Y * T, seed=11 * Y + T)
    want, wacc, we = _oracle_trajectory(size, v0, beta, seed, wilson)
    for flags in (NORES, NORES | _lib.FLAG_NO_GRAPH):
        with Engine(X, Y, T, flags=flags) as eng:
            got, acc, e = _engine_trajectory(eng, v0, beta, seed, wilson)
            plan = eng.plan_info()
            assert plan["path"] == "narrow" and plan["chunks"] == chunks and plan["rows_per_chunk"] >= 3, plan
            assert eng.narrow_segments() == 1 and plan["storage"] == "u16" and plan["i32_allocated"] == 0, plan
        assert acc == wacc
        assert (e == we).all()
        assert (got == want).all(), f"{int((got != want).sum())} sites differ (chunks={chunks}, flags={flags})"


@pytest.mark.parametrize("size,wilson,self_halo", [
    ((256, 10, 6), (100, 5), False), ((128, 16, 40), (0, 0), False), ((64, 12, 64), (21, 32), False), ((1024, 6, 4), (700, 2), False),
    ((16, 6, 1100), (5, 16), False), ((256, 10, 36), (255, 20), True), ((512, 6, 8), (0, 0), True),
])
def test_narrow_one_row_per_chunk_vs_oracle(size, wilson, self_halo):
    """The plan of a small lattice: as many y-chunks as rows (128^3 runs 128 chunks of one row).  The register window is
    loaded, used once and dropped; first rows of both parities; on slabs the ghost tasks of a one-row chunk."""
    X, Y, T = size
    beta, seed = 0.26, 0x1205 + X
    v0 = seeded_field(X * Y * T, seed=5 * Y + T)
    want, wacc, we = _oracle_trajectory(size, v0, beta, seed, wilson)
    for flags in (NORES, NORES | _lib.FLAG_NO_GRAPH):
        with Engine(X, Y, T, flags=flags | (_lib.FLAG_SELF_HALO if self_halo else 0)) as eng:
            got, acc, e = _engine_trajectory(eng, v0, beta, seed, wilson)
            plan = eng.plan_info()
            assert plan["path"] == "narrow" and plan["chunks"] == Y and plan["rows_per_chunk"] == 1, plan
        assert acc == wacc
        assert (e == we).all()
        assert (got == want).all(), f"{int((got != want).sum())} sites differ (flags={flags})"


@pytest.mark.parametrize("chunks", [1, 3])
@pytest.mark.parametrize("size,wilson", [((256, 10, 40), (77, 33)), ((512, 14, 24), (0, 0)), ((1024, 10, 20), (600, 9))])
def test_self_halo_rows_per_chunk_ge_3_vs_oracle(size, wilson, chunks, monkeypatch):
    """The slab path (deep halo, interior under the exchange, two-range wedge launches) with long chunks."""
    monkeypatch.setenv("LQFT_N16_CHUNKS", str(chunks))
    X, Y, T = size
    beta, seed = 0.25, 0x5E1F + chunks
    v0 = seeded_field(X * Y * T, seed=Y + T)
    want, wacc, we = _oracle_trajectory(size, v0, beta, seed, wilson)
    with Engine(X, Y, T, flags=_lib.FLAG_SELF_HALO) as eng:
        got, acc, e = _engine_trajectory(eng, v0, beta, seed, wilson)
        plan = eng.plan_info()
        assert plan["path"] == "narrow" and plan["rows_per_chunk"] >= 3 and plan["ghost_depth"] >= 2, plan
        eng.synchronize()
    assert acc == wacc and (e == we).all()
    assert (got == want).all(), f"{int((got != want).sum())} sites differ"


@pytest.mark.parametrize("size,min_rows", [((256, 256, 128), 3), ((256, 256, 256), 6), ((512, 256, 64), 4), ((1024, 256, 32), 4),
                                           ((256, 128, 256), 3)])
def test_narrow_natural_chunking_rows_per_chunk_ge_3_vs_scalar_kernel(size, min_rows):
    """The cost model's own chunking at bench-sized lattices (256^3: 37 chunks of 6-7 rows), against the scalar i32
    kernel (LQFT_FLAG_FORCE_GENERIC): post-sweep fields, acceptance counts and in-run energies bit for bit."""
    X, Y, T = size
    beta, seed = 0.25, 0xB16 + T
    v0 = seeded_field(X * Y * T, seed=T, lo=-3, hi=4)
    out = []
    for flags in (0, _lib.FLAG_FORCE_GENERIC):
        with Engine(X, Y, T, flags=flags) as eng:
            out.append(_engine_trajectory(eng, v0, beta, seed, (X // 3, T // 2)))
            plan = eng.plan_info()
            if flags == 0:
                assert plan["path"] == "narrow" and plan["rows_per_chunk"] >= min_rows, plan
                assert eng.narrow_segments() == 1 and plan["storage"] == "u16" and plan["i32_allocated"] == 0, plan
            else:
                assert plan["path"] == "generic", plan
    assert out[0][1] == out[1][1]
    assert (out[0][2] == out[1][2]).all()
    assert (out[0][0] == out[1][0]).all(), f"{int((out[0][0] != out[1][0]).sum())} sites differ"


def test_256cubed_narrow_rows_per_chunk_ge_3_vs_oracle():
    """The headline shape itself, 256^3, through the kernel and chunking bench.py times, against the oracle."""
    size = (256, 256, 256)
    X, Y, T = size
    beta, seed = 0.25, 20261019
    v0 = seeded_field(X * Y * T, seed=256, lo=-2, hi=3)
    lat = O.Lattice(list(size))
    want = v0.copy()
    wacc = sum(O.sweep(lat, want, beta, seed, s, order=1) for s in range(4))
    we = O.energy_observable(lat, want, beta)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, v0)
        acc = eng.sweep(0, 4, want_accepted=True)[0]
        plan = eng.plan_info()
        assert plan["path"] == "narrow" and plan["rows_per_chunk"] >= 6, plan
        assert eng.narrow_segments() == 1
        got = eng.download(0)
        e = eng.energy()[1][0]
    assert acc == wacc and e == we
    assert (got == want).all(), f"{int((got != want).sum())} sites differ"


@pytest.mark.parametrize("chunks", [0, 1, 3])
@pytest.mark.parametrize("size,wilson", [
    ((16, 16, 16), (5, 16)), ((16, 10, 70), (0, 0)), ((32, 12, 8), (0, 0)), ((32, 10, 36), (21, 30)),
    ((64, 64, 64), (21, 32)), ((64, 14, 18), (0, 0)), ((64, 10, 2), (64, 2)), ((128, 16, 12), (0, 0)),
    ((128, 10, 34), (100, 20)), ((128, 128, 128), (0, 0)),
])
def test_narrow_rows_shorter_than_a_warp_vs_oracle(size, wilson, chunks, monkeypatch):
    """X = 16 ... 128: a warp carries 32 / (X / 16) planes of one parity (kernels_narrow.cuh, PPW); the x-edge
    neighbour comes from a shuffle within the lanes of one row (X = 16: the thread's own other-colour quad).
    Plane counts that do not fill a warp or a block, forced long chunks, Wilson sheets, counts, graphs."""
    if chunks:
        monkeypatch.setenv("LQFT_N16_CHUNKS", str(chunks))
    X, Y, T = size
    beta, seed = 0.26, 0xFACE + X + T
    v0 = seeded_field(X * Y * T, seed=X + Y + T)
    want, wacc, we = _oracle_trajectory(size, v0, beta, seed, wilson)
    for flags in (NORES, NORES | _lib.FLAG_NO_GRAPH):
        with Engine(X, Y, T, flags=flags) as eng:
            got, acc, e = _engine_trajectory(eng, v0, beta, seed, wilson)
            plan = eng.plan_info()
            assert plan["path"] == "narrow", plan
            assert eng.narrow_segments() == 1 and plan["storage"] == "u16" and plan["i32_allocated"] == 0, plan
        assert acc == wacc
        assert (e == we).all()
        assert (got == want).all(), f"{int((got != want).sum())} sites differ (chunks={chunks}, flags={flags})"


def test_narrow_64cubed_many_chains_match_single_chain_runs():
    """Config 2's shape (64^3 chains batched along blockIdx.z) on the 16-bit kernel: five chains with their own
    couplings, seeds and Wilson sheets equal five single-chain runs on the scalar kernel."""
    X = Y = T = 64
    betas = [0.2, 0.23, 0.25, 0.28, 0.31]
    v0 = [seeded_field(X * Y * T, seed=640 + c) for c in range(5)]
    single = []
    for c in range(5):
        with Engine(X, Y, T, flags=_lib.FLAG_FORCE_GENERIC) as eng:
            eng.set_chain(0, betas[c], 64000 + c, 10 * c, 13 * c)
            eng.upload(0, v0[c])
            acc = eng.sweep(0, 6, want_accepted=True)[0]
            e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
            single.append((acc, e[0].copy(), eng.download(0)))
    with Engine(X, Y, T, n_chains=5, flags=NORES) as eng:   # NO_RESIDENT: the launch-per-half-sweep kernel, not the cluster one
        for c in range(5):
            eng.set_chain(c, betas[c], 64000 + c, 10 * c, 13 * c)
            eng.upload(c, v0[c])
        acc = eng.sweep(0, 6, want_accepted=True)
        e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
        assert eng.plan_info()["path"] == "narrow" and eng.narrow_segments() == 5 and eng.plan_info()["storage"] == "u16"
        for c in range(5):
            assert acc[c] == single[c][0]
            assert (e[c] == single[c][1]).all()
            assert (eng.download(c) == single[c][2]).all()


def test_one_sweep_per_call_and_standalone_observables_stay_on_the_16_bit_arrays():
    """The Algorithm::field_sweep drop-in: one sweep per lqft_sweep call, with every standalone operation in between
    (energy, action, normalize at a site, slice moments, correlator, shift, download) served from the 16-bit arrays —
    no i32 arrays are ever allocated, and everything equals the oracle."""
    X, Y, T = 64, 12, 6
    beta, seed = 0.27, 0x51E
    lat = O.Lattice([X, Y, T])
    v0 = seeded_field(X * Y * T, seed=99)
    want = v0.copy()
    sites = [int(O.pick_site(seed, 4, 3, r, lat.n_sites)) for r in range(6)]
    with Engine(X, Y, T, flags=NORES) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, v0)
        for s in range(7):
            acc = eng.sweep(s, 1, want_accepted=True)[0]
            assert acc == O.sweep(lat, want, beta, seed, s, order=1)
            assert int(eng.energy()[0][0]) == O.integer_energy(lat, want)[0]
            assert float(eng.action()[0][0]) == O.float_action(lat, want)
            if s == 2:
                eng.normalize_site(0, 1234)
                O.shift_values(want, int(want[1234]))
            if s == 4:
                eng.shift_values(0, -7)
                O.shift_values(want, -7)
            s1, s2 = eng.slice_moments(0)
            f = want.reshape(T, Y * X).astype(np.int64)
            assert (s1 == f.sum(axis=1)).all() and (s2 == (f * f).sum(axis=1)).all()
            assert (eng.correlator(0, ref_sites=sites) == O.correlator(lat, want, sites)).all()
            assert (eng.download(0) == want).all(), s
        plan = eng.plan_info()
        assert plan["path"] == "narrow" and plan["storage"] == "u16" and plan["i32_allocated"] == 0, plan


@pytest.mark.parametrize("size,wilson", [((256, 24, 6), (0, 0)), ((512, 10, 4), (300, 3)), ((1024, 6, 4), (0, 0)), ((64, 20, 10), (21, 7)),
                                         ((16, 8, 12), (0, 0))])
def test_energy_inside_the_colour_1_half_sweep_vs_observe_kernel_and_oracle(size, wilson):
    """A step whose energy is due accumulates it inside its colour-1 half-sweep (n16_obs_pair: an exact integer
    identity over the six bonds of every colour-1 site).  Same numbers as the separate observe kernel
    (LQFT_FLAG_NO_FUSED_ENERGY) and as the oracle's literal loop (fields.rs:728-737), eager and graph-replayed, with
    acceptance counts alongside."""
    X, Y, T = size
    beta, seed = 0.26, 0xE0 + X
    lat = O.Lattice(list(size))
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    v0 = seeded_field(X * Y * T, seed=X + Y)
    want = v0.copy()
    we, wacc = O.run(lat, want, beta, seed, 0, 0, 40, 1, 10, order=1, wilson=wil)
    for flags in (NORES, NORES | _lib.FLAG_NO_GRAPH, NORES | _lib.FLAG_NO_FUSED_ENERGY):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed, *wilson)
            eng.upload(0, v0)
            launches = eng.kernel_launches()
            e, e_int, acc = eng.run(40, energy_every=1, normalize_every=10, want_accepted=True)
            per_step = (eng.kernel_launches() - launches) / 40
            assert (e[0] == we).all(), flags
            assert acc[0] == wacc
            assert (eng.download(0) == want).all()
            assert int(e_int[0][-1]) == O.integer_energy(lat, want)[0]
