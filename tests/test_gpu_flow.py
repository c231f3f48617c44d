"""Parity of k_run_flow (kernels_flow.cuh): a whole run in one persistent launch on the 16-bit arrays in global memory,
half-sweeps ordered by per-tile progress counters (release / acquire at gpu scope) instead of kernel boundaries.  A
missed dependency — a tile reading a neighbour's row before it was written, or overwriting one before it was read —
changes decisions at once, so every case compares fields, acceptance counts and logged observables bit for bit with the
oracle (metropolis_single, algorithm.rs:127-158, through oracle/lqft_oracle.c) or with the launch-per-half-sweep path."""
import numpy as np
import pytest

from lattice_qft_b200 import Engine, _lib
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu


def _oracle(size, v0, beta, seed, wilson, n_sweeps=5, n_run=24):
    lat = O.Lattice(list(size))
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    want = v0.copy()
    acc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(n_sweeps))
    e, _ = O.run(lat, want, beta, seed, n_sweeps, 0, n_run, 3, 10, order=1, wilson=wil)
    return lat, wil, want, acc, e


@pytest.fixture(autouse=True)
def _flow_first(monkeypatch):
    monkeypatch.setenv("LQFT_FLOW_FIRST", "1")


# chunks: 0 = the plan's own choice (as many y chunks as block slots allow, two rows each on these small lattices);
# 1, 3, 5 force long and uneven chunks, odd first rows, the loop-back of the three-row register window
@pytest.mark.parametrize("chunks", [0, 1, 3, 5])
@pytest.mark.parametrize("size,wilson", [
    ((256, 24, 40), (0, 0)),      # two plane groups and a partly filled one (40 planes, 32 per group)
    ((256, 38, 8), (100, 5)),     # surplus plane slots in every block
    ((512, 20, 12), (0, 0)),      # one plane per warp
    ((1024, 10, 6), (700, 4)),    # two row segments per plane: the x neighbour across the segment edge
    ((128, 40, 64), (0, 0)),      # four planes per warp
    ((64, 64, 64), (21, 32)),     # eight planes per warp
    ((16, 8, 1200), (5, 16)),     # a warp holds 32 planes; three groups, the last one partly filled
    ((1024, 64, 2), (0, 0)),      # two planes: every t neighbour is the partner block's, periodic wrap on itself
])
def test_flow_run_vs_oracle(size, wilson, chunks, monkeypatch):
    X, Y, T = size
    if chunks:
        monkeypatch.setenv("LQFT_N16_CHUNKS", str(chunks))
    beta, seed = 0.23 + 0.01 * chunks, 0xF10 + X + T
    v0 = seeded_field(X * Y * T, seed=X + 3 * Y + T)
    lat, wil, want, wacc, we = _oracle(size, v0, beta, seed, wilson)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, beta, seed, *wilson)
        eng.upload(0, v0)
        plan = eng.plan_info()
        assert plan["path"] == "flow", plan
        if chunks:
            assert plan["chunks"] == min(chunks, max(Y // 2, 1)), plan
        acc = eng.sweep(0, 5, want_accepted=True)[0]
        launches = eng.kernel_launches()
        e, _, _ = eng.run(24, sweep_base=5, energy_every=3, normalize_every=10)
        assert eng.kernel_launches() == launches + 1       # the whole run is one launch
        got = eng.download(0)
        e_int = int(eng.energy()[0][0])
        s_int = int(eng.action()[0][0])
        assert eng.plan_info()["storage"] == "u16"
    assert acc == wacc
    assert (e == we).all(), (e, we)
    assert (got == want).all(), f"{int((got != want).sum())} sites differ"
    assert e_int == O.integer_energy(lat, want)[0] and float(s_int) == O.float_action(lat, want)


def test_flow_rows_per_chunk_at_least_3_on_wide_lattice():
    """256 x 256 x 128: the plan's own chunks are 3 to 4 rows long (74 chunks), every block slot of the device is
    taken, and the tiles run through 40 half-sweeps on nothing but their neighbours' counters.  Against the launch path
    and the scalar i32 kernel."""
    X, Y, T = 256, 256, 128
    beta, seed = 0.25, 0xF256
    v0 = seeded_field(X * Y * T, seed=77)
    res = []
    for flags in (0, _lib.FLAG_NO_RESIDENT, _lib.FLAG_FORCE_GENERIC | _lib.FLAG_NO_NARROW):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed)
            eng.upload(0, v0)
            plan = eng.plan_info()
            acc = eng.sweep(0, 8, want_accepted=True)[0]
            e, _, _ = eng.run(12, sweep_base=8, energy_every=4, normalize_every=10)
            res.append((plan, acc, e, eng.download(0)))
    assert res[0][0]["path"] == "flow" and res[0][0]["rows_per_chunk"] >= 3, res[0][0]
    assert res[1][0]["path"] == "narrow" and res[2][0]["path"] == "generic"
    for other in res[1:]:
        assert res[0][1] == other[1]
        assert (res[0][2] == other[2]).all()
        assert (res[0][3] == other[3]).all()


def test_flow_observables_logs_match_launch_path():
    """Energy and action on different frequencies from the colour-1 half-sweeps, first_step offsets, several calls in a
    row (slots continue), correlator and difference plot measured between persistent launches."""
    X, Y, T = 128, 32, 32
    beta, seed = 0.27, 778
    lat = O.Lattice([X, Y, T])
    v0 = seeded_field(X * Y * T, seed=22)
    res = []
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed)
            eng.upload(0, v0)
            a = eng.run_observables(30, first_step=4, sweep_base=0, energy_every=2, action_every=5, normalize_every=7)
            b = eng.run_observables(41, first_step=34, sweep_base=0, energy_every=3, action_every=1, normalize_every=10,
                                    correlator_every=10, correlator_reps=20, difference_every=4)
            mean, count = eng.difference_read(0)
            res.append((a, b, eng.download(0), eng.plan_info()["path"], mean, count))
    assert res[0][3] == "flow" and res[1][3] == "narrow"
    for k in ("e_int", "s_int", "e_obs", "s_obs", "corr"):
        assert (res[0][0][k] == res[1][0][k]).all() and (res[0][1][k] == res[1][1][k]).all(), k
    assert res[0][1]["corr"].shape[1] == 4 and res[0][1]["corr"].any()
    assert (res[0][2] == res[1][2]).all()
    assert res[0][5] == res[1][5] == 10 and (res[0][4] == res[1][4]).all()
    assert float(res[0][1]["s_int"][0][-1]) == O.float_action(lat, res[0][2])


def test_flow_many_chains_match_single_chain_runs():
    """Chains share a launch (blockIdx.z) but not a counter: each with its own coupling, seed and Wilson sheet, equal to
    single-chain runs on the scalar kernel."""
    X, Y, T = 64, 32, 48
    n = 24
    betas = [0.2 + 0.004 * c for c in range(n)]
    v0 = [seeded_field(X * Y * T, seed=300 + c) for c in range(n)]
    with Engine(X, Y, T, n_chains=n) as eng:
        for c in range(n):
            eng.set_chain(c, betas[c], 7000 + c, (7 * c) % 64, (11 * c) % 48)
            eng.upload(c, v0[c])
        assert eng.plan_info()["path"] == "flow"
        acc = eng.sweep(0, 6, want_accepted=True)
        e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
        got = [eng.download(c) for c in range(n)]
    for c in (0, 1, 13, 23):
        with Engine(X, Y, T, flags=_lib.FLAG_FORCE_GENERIC) as ref:
            ref.set_chain(0, betas[c], 7000 + c, (7 * c) % 64, (11 * c) % 48)
            ref.upload(0, v0[c])
            racc = ref.sweep(0, 6, want_accepted=True)[0]
            re, _, _ = ref.run(20, sweep_base=6, energy_every=5, normalize_every=10)
            assert acc[c] == racc and (e[c] == re[0]).all()
            assert (got[c] == ref.download(0)).all()


def test_flow_long_run_keeps_step_with_launch_path():
    """200 sweeps in one launch (400 half-sweeps of hand-offs between tiles) at the headline row width."""
    X, Y, T = 256, 64, 32
    v0 = seeded_field(X * Y * T, seed=5)
    out = []
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, 0.25, 99)
            eng.upload(0, v0)
            e, _, acc = eng.run(200, energy_every=10, normalize_every=10, want_accepted=True)
            out.append((e, acc, eng.download(0)))
    assert (out[0][0] == out[1][0]).all() and out[0][1] == out[1][1] and (out[0][2] == out[1][2]).all()
