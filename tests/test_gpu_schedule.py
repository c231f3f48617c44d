"""The device-side measurement schedule (lqft_run_observables): every OutputData kind of the reference's loop
(computation.rs:247-270) — energy, action, correlator, difference plot — on its own frequency inside ONE call,
against the oracle's literal loops (outputdata.rs:168-178, 202-212, 241-256, 313-344) and against the host-loop
path (one API call per due measurement)."""
import numpy as np
import pytest

from lattice_qft_b200 import Engine, _lib
from lattice_qft_b200 import engine as E
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu


def _oracle_loop(size, v, beta, seed, sweep_base, first, n, ee, ae, ce, reps, de, ne, wilson=(0, 0)):
    """computation.rs:247-270 with the decision spec's RNG: sweep, observables that are due, normalize."""
    lat = O.Lattice(list(size))
    wil = O.Wilson(lat, *wilson) if wilson[0] else None
    n_sites = lat.n_sites
    e, a, c = [], [], []
    wel = O.new_welford_cells(3 * n_sites)
    dn = 0
    for step in range(first, first + n):
        O.sweep(lat, v, beta, seed, sweep_base + step, order=1, wilson=wil)
        if ee and step % ee == 0:
            e.append(O.energy_observable(lat, v, beta))
        if ae and step % ae == 0:
            a.append(beta * O.float_action(lat, v))
        if ce and step % ce == 0:
            sites = [E.pick_site(seed, _lib.STREAM_CORRELATOR, sweep_base + step, r, n_sites) for r in range(reps)]
            c.append(O.correlator(lat, v, sites))
        if de and step % de == 0:
            O.difference_update(lat, v, wel)
            dn += 1
        if ne and step % ne == 0:
            site = E.pick_site(seed, _lib.STREAM_NORMALIZE, sweep_base + step, 0, n_sites)
            v -= v[site]
    return np.array(e), np.array(a), np.array(c), O.difference_means(lat, wel), dn


@pytest.mark.parametrize("flags", [0, _lib.FLAG_NO_RESIDENT, _lib.FLAG_NO_RESIDENT | _lib.FLAG_NO_GRAPH, _lib.FLAG_FORCE_GENERIC,
                                   _lib.FLAG_SELF_HALO])
@pytest.mark.parametrize("size,sched", [
    ((16, 16, 16), dict(first=0, n=40, ee=10, ae=4, ce=10, reps=7, de=20)),
    ((32, 8, 12), dict(first=3, n=37, ee=0, ae=1, ce=5, reps=100, de=0)),
    ((10, 6, 4), dict(first=0, n=24, ee=3, ae=0, ce=4, reps=3, de=6)),
    ((256, 4, 6), dict(first=0, n=30, ee=5, ae=10, ce=10, reps=300, de=15)),
    ((64, 16, 8), dict(first=10, n=30, ee=10, ae=10, ce=10, reps=100, de=0)),
])
def test_run_observables_matches_oracle_loop(size, sched, flags):
    X, Y, T = size
    if flags & _lib.FLAG_SELF_HALO and T < 4:
        pytest.skip("slab too thin")
    beta, seed, sweep_base = 0.24, 0x0B5 + X, 100
    v0 = seeded_field(X * Y * T, seed=X + T, lo=-20, hi=21)
    want = v0.copy()
    we, wa, wc, wd, dn = _oracle_loop(size, want, beta, seed, sweep_base, sched["first"], sched["n"], sched["ee"], sched["ae"],
                                      sched["ce"], sched["reps"], sched["de"], 10)
    with Engine(X, Y, T, flags=flags) as eng:
        eng.set_chain(0, beta, seed)
        eng.upload(0, v0)
        res = eng.run_observables(sched["n"], first_step=sched["first"], sweep_base=sweep_base, energy_every=sched["ee"],
                                  action_every=sched["ae"], correlator_every=sched["ce"], correlator_reps=sched["reps"],
                                  difference_every=sched["de"], normalize_every=10, want_accepted=True)
        got = eng.download(0)
        mean, count = eng.difference_read(0)
    assert (got == want).all()
    assert res["e_obs"].shape == (1, len(we)) and (res["e_obs"][0] == we).all()
    assert res["s_obs"].shape == (1, len(wa)) and (res["s_obs"][0] == wa).all()
    assert res["corr"].shape == (1, len(wc), T)
    if len(wc):
        assert (res["corr"][0] == wc).all()
    assert count == dn
    if dn:
        assert (mean.reshape(-1) == wd).all()


def test_run_observables_equals_host_loop_many_chains():
    """Three chains, Wilson sheets included (WilsonSim measures the plain field, quirk Q3): the fused call equals
    sweeping step by step and calling lqft_energy / lqft_action / lqft_correlator when due."""
    X, Y, T = 64, 8, 16
    betas, wil = [0.21, 0.25, 0.3], [(0, 0), (20, 8), (64, 16)]
    v0 = [seeded_field(X * Y * T, seed=900 + c, lo=-5, hi=6) for c in range(3)]

    def make():
        eng = Engine(X, Y, T, n_chains=3)
        for c in range(3):
            eng.set_chain(c, betas[c], 7000 + c, *wil[c])
            eng.upload(c, v0[c])
        return eng

    with make() as eng:
        res = eng.run_observables(40, sweep_base=5, energy_every=2, action_every=5, correlator_every=10, correlator_reps=100,
                                  normalize_every=10)
        final = [eng.download(c) for c in range(3)]
    e, a, cr = [[], [], []], [[], [], []], [[], [], []]
    with make() as eng:
        for step in range(40):
            eng.sweep(5 + step, 1)
            if step % 2 == 0:
                eo = eng.energy()[1]
                for c in range(3):
                    e[c].append(eo[c])
            if step % 5 == 0:
                so = eng.action()[1]
                for c in range(3):
                    a[c].append(so[c])
            if step % 10 == 0:
                for c in range(3):
                    cr[c].append(eng.correlator(c, reps=100, sweep_index=5 + step))
                eng.normalize(5 + step)
        for c in range(3):
            assert (eng.download(c) == final[c]).all()
    for c in range(3):
        assert (res["e_obs"][c] == np.array(e[c])).all()
        assert (res["s_obs"][c] == np.array(a[c])).all()
        assert (res["corr"][c] == np.array(cr[c])).all()


def test_sim_rs_workload_is_one_call_at_128cubed():
    """src/bin/sim.rs:78-86: 128^3, new_correlation_plot(&lattice, 100).set_frequency(10) — the whole measurement
    phase is ONE lqft_run_observables call; rows equal lqft_correlator's and the oracle's literal double loop."""
    X = Y = T = 128
    beta, seed = 0.25, 0xABCD
    with Engine(X, Y, T) as eng:
        eng.set_chain(0, beta, seed)
        eng.init_hot()
        eng.run(20, energy_every=0)
        start = eng.download(0)
        launches0 = eng.kernel_launches()
        res = eng.run_observables(40, sweep_base=20, correlator_every=10, correlator_reps=100, normalize_every=10)
        assert res["corr"].shape == (1, 4, T)
        final = eng.download(0)
        last = eng.correlator(0, reps=100, sweep_index=20 + 30)   # configuration after step 39 != after step 30
    lat = O.Lattice([X, Y, T])
    v = start.copy()
    for step in range(40):
        O.sweep(lat, v, beta, seed, 20 + step, order=1)
        if step % 10 == 0:
            sites = [E.pick_site(seed, _lib.STREAM_CORRELATOR, 20 + step, r, lat.n_sites) for r in range(100)]
            assert (res["corr"][0][step // 10] == O.correlator(lat, v, sites)).all()   # the literal 100 x N loop
            v -= v[E.pick_site(seed, _lib.STREAM_NORMALIZE, 20 + step, 0, lat.n_sites)]
    assert (final == v).all()
    assert last.shape == (T,)
