"""Parity of k_run_cluster (kernels_cluster.cuh): a chain resident in the shared memory of a thread-block cluster for a
whole call — half-sweeps from shared memory, boundary planes pushed into the neighbour CTAs' halo planes through
distributed shared memory, energy and action accumulated inside the colour-1 half-sweep, normalize by one thread.
Everything must equal the oracle (metropolis_single, algorithm.rs:127-158, through oracle/lqft_oracle.c) and the
launch-per-half-sweep kernel bit for bit."""
import numpy as np
import pytest

from lattice_qft_b200 import Engine, _lib
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu


def _oracle(size, v0, beta, seed, wilson):
    lat = O.Lattice(list(size))
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    want = v0.copy()
    acc = sum(O.sweep(lat, want, beta, seed, s, order=1, wilson=wil) for s in range(5))
    e, _ = O.run(lat, want, beta, seed, 5, 0, 24, 3, 10, order=1, wilson=wil)
    return lat, wil, want, acc, e


@pytest.mark.parametrize("ctas", [1, 2, 4, 8])
@pytest.mark.parametrize("size,wilson", [
    ((64, 64, 64), (21, 32)), ((64, 10, 80), (0, 0)), ((16, 64, 64), (5, 16)), ((32, 48, 40), (0, 0)), ((128, 16, 32), (100, 9)),
    ((128, 64, 64), (0, 0)), ((64, 2, 512), (64, 3)), ((16, 4, 1024), (0, 0)),   # all beyond the one-CTA i32 kernel (36^3)
])
def test_cluster_resident_run_vs_oracle(size, wilson, ctas, monkeypatch):
    """Cluster sizes 1 ... 8 (forced: a lone chain is faster on the launch path and the plan leaves it there), one plane per CTA up to the whole chain in one CTA,
    Wilson sheets, acceptance counts, in-run energies, normalisation; the field afterwards serves every standalone
    operation from the 16-bit arrays."""
    X, Y, T = size
    if ctas:
        if T % ctas:
            pytest.skip("cluster size must divide T")
        monkeypatch.setenv("LQFT_CLUSTER_CTAS", str(ctas))
    beta, seed = 0.24 + 0.01 * ctas, 0xC1 + X + T
    v0 = seeded_field(X * Y * T, seed=X + 3 * Y + T)
    lat, wil, want, wacc, we = _oracle(size, v0, beta, seed, wilson)
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, beta, seed, *wilson)
        eng.upload(0, v0)
        plan = eng.plan_info()
        if plan["path"] != "cluster":          # a forced size that does not fit: nothing to test
            pytest.skip(f"{ctas} CTAs cannot hold {size}")
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


def test_cluster_run_observables_energy_and_action_logs(monkeypatch):
    """Energy and action on different frequencies from the colour-1 half-sweeps, first_step offsets, several calls
    in a row (slots continue), against the launch path (LQFT_FLAG_NO_RESIDENT) and the oracle."""
    monkeypatch.setenv("LQFT_CLUSTER_CTAS", "4")
    X, Y, T = 64, 32, 32
    beta, seed = 0.27, 777
    lat = O.Lattice([X, Y, T])
    v0 = seeded_field(X * Y * T, seed=21)
    res = []
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        with Engine(X, Y, T, flags=flags) as eng:
            eng.set_chain(0, beta, seed)
            eng.upload(0, v0)
            a = eng.run_observables(30, first_step=4, sweep_base=0, energy_every=2, action_every=5, normalize_every=7)
            b = eng.run_observables(11, first_step=34, sweep_base=0, energy_every=3, action_every=1, normalize_every=10)
            res.append((a, b, eng.download(0), eng.plan_info()["path"]))
    assert res[0][3] == "cluster" and res[1][3] == "narrow"
    for k in ("e_int", "s_int", "e_obs", "s_obs"):
        assert (res[0][0][k] == res[1][0][k]).all() and (res[0][1][k] == res[1][1][k]).all(), k
    assert (res[0][2] == res[1][2]).all()
    # the last logged action against the oracle on the final field is not available (the run goes on after it), but
    # an action measured after every step is: the last entry of b is the action of the downloaded field
    assert float(res[0][1]["s_int"][0][-1]) == O.float_action(lat, res[0][2])


def test_cluster_many_chains_match_single_chain_runs():
    """64^3 chains, each in a cluster of its own, with their own couplings, seeds and Wilson sheets — as many as the plan
    takes to the cluster path by itself (clusters resident together and on at least 3/8 of the SMs: 8 to 15 chains of
    64^3); equal to single-chain runs on the scalar kernel.  Fewer or more chains stay on the launch path."""
    X = Y = T = 64
    n = 12
    betas = [0.2 + 0.003 * c for c in range(n)]
    v0 = [seeded_field(X * Y * T, seed=900 + c) for c in range(n)]
    with Engine(X, Y, T, n_chains=n) as eng:
        for c in range(n):
            eng.set_chain(c, betas[c], 5000 + c, (7 * c) % 64, (11 * c) % 64)
            eng.upload(c, v0[c])
        assert eng.plan_info()["path"] == "cluster"
        acc = eng.sweep(0, 6, want_accepted=True)
        e, _, _ = eng.run(20, sweep_base=6, energy_every=5, normalize_every=10)
        got = [eng.download(c) for c in range(n)]
    for other in (4, 16):
        with Engine(X, Y, T, n_chains=other) as eng:
            for c in range(other):
                eng.set_chain(c, 0.25, c)
            assert eng.plan_info()["path"] == "narrow"
    for c in (0, 1, 7, 11):
        with Engine(X, Y, T, flags=_lib.FLAG_FORCE_GENERIC) as ref:
            ref.set_chain(0, betas[c], 5000 + c, (7 * c) % 64, (11 * c) % 64)
            ref.upload(0, v0[c])
            racc = ref.sweep(0, 6, want_accepted=True)[0]
            re, _, _ = ref.run(20, sweep_base=6, energy_every=5, normalize_every=10)
            assert acc[c] == racc and (e[c] == re[0]).all()
            assert (got[c] == ref.download(0)).all()
