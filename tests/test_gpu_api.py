"""The reference-facing interface (lattice / fields / algorithm / outputdata / computation), exercised
the way the reference's own code uses it (src/bin/sim.rs, src/computation.rs)."""
import numpy as np
import pytest

from lattice_qft_b200.algorithm import AlgorithmType, Metropolis, SweepRng
from lattice_qft_b200.computation import Computation, parse_simulation_results
from lattice_qft_b200.fields import Field, WilsonField
from lattice_qft_b200.lattice import Lattice, Lattice3d
from lattice_qft_b200.outputdata import OutputData
from oracle import oracle as O
from conftest import seeded_field

pytestmark = pytest.mark.gpu


def test_field_sweep_like_the_reference():
    lattice = Lattice3d.new(8, 8, 8)
    field = Field.from_values(lattice, seeded_field(512))
    rng = SweepRng(seed=77)
    algo = AlgorithmType.new_metropolis()
    acc = [algo.field_sweep(field, 0.2, rng) for _ in range(3)]
    olat = O.Lattice([8, 8, 8])
    want = seeded_field(512)
    wacc = [O.sweep(olat, want, 0.2, 77, s, order=1) for s in range(3)]
    assert acc == wacc and (field.values == want).all()
    assert field.integer_energy() == O.integer_energy(olat, want)[0]
    assert field.energy_observable(0.2) == O.energy_observable(olat, want, 0.2)
    assert field.action_observable(0.2) == O.action_observable(olat, want, 0.2)
    assert field.local_action(17) == O.lib().orc_local_action(olat._h, want.ctypes.data, 17, int(want[17]))


def test_wilson_sweep_like_the_reference():
    lattice = Lattice3d.new(8, 4, 6)
    wilson = WilsonField.from_field(Field.from_values(lattice, seeded_field(192)), 3, 2)
    rng = SweepRng(seed=5)
    acc = Metropolis().wilson_sweep(wilson, 0.26, rng)
    olat = O.Lattice([8, 4, 6])
    want = seeded_field(192)
    assert acc == O.sweep(olat, want, 0.26, 5, 0, order=1, wilson=O.Wilson(olat, 3, 2))
    assert (wilson.field.values == want).all()
    # bond symmetry of fields.rs:884-895 on the Wilson-aware accessors
    for index in range(0, 192, 5):
        for direction in range(6):
            nb = lattice.values[index][direction]
            assert wilson.bond_action(index, direction) == wilson.bond_action(nb, (direction + 3) % 6)


def test_wilson_links_of_reference_test():  # fields.rs:866-881
    lattice = Lattice.new([2, 2, 2])
    wilson = WilsonField.from_field(Field.new(lattice), 1, 1)
    active = [(i, d) for i in range(8) for d in range(6) if wilson.is_active(i, d)]
    assert active == [(0, 1), (2, 4)]


def test_field_shift_and_normalize():  # fields.rs:851-861
    lattice = Lattice3d.new(4, 4, 4)
    field = Field.new(lattice)
    field.shift_values(3)
    assert field.values[63] == -3
    field = Field.from_values(lattice, seeded_field(64))
    before = field.values.copy()
    field.normalize_random(SweepRng(3))
    d = before - field.values
    assert (d == d[0]).all() and (field.values == 0).any()


def test_simulation_run_matches_oracle_schedule(tmp_path):
    lattice = Lattice3d.new(8, 8, 8)
    temp, burnin, iterations = 0.24, 20, 60
    obs = [OutputData.new_energy_observable(temp).set_frequency(10),
           OutputData.new_action_observable(temp).set_frequency(20),
           OutputData.new_correlation_plot(lattice, 5).set_frequency(30),
           OutputData.new_difference_plot(lattice).set_frequency(30)]
    comp = Computation.new_simulation(lattice, temp, AlgorithmType.new_metropolis(), burnin, iterations, obs,
                                      seed=4242).run()
    olat = O.Lattice([8, 8, 8])
    want = O.init_hot(olat, 4242)
    O.run(olat, want, temp, 4242, 0, 0, burnin, 0, 10, order=1)
    energies, actions = [], []
    for step in range(iterations):
        O.sweep(olat, want, temp, 4242, burnin + step, order=1)
        if step % 10 == 0:
            energies.append(O.energy_observable(olat, want, temp))
        if step % 20 == 0:
            actions.append(O.action_observable(olat, want, temp))
        if step % 10 == 0:
            O.shift_values(want, int(want[O.pick_site(4242, 2, burnin + step, 0, 512)]))
    assert (comp.field.values == want).all()
    assert obs[0].data.results() == energies and obs[1].data.results() == actions
    assert len(obs[2].data.plot()) == 2 and len(obs[2].data.plot()[0]) == 8
    assert str(comp) == "Metropolis Simulation"
    # CSV layout of parse_simulation_results (computation.rs:129-199)
    parse_simulation_results([comp], root=str(tmp_path))
    rows = (tmp_path / "data" / "results.csv").read_text().splitlines()
    assert rows[0] == ("index,d,size,x,y,t,temp,comptype,burnin,iterations,duration,algorithm_statistic,"
                       "action_data,energy_data,difference_data,correlation_data")
    assert rows[1].startswith("0,3,512,8,8,8,0.24,Metropolis Simulation,20,60,") and rows[1].endswith(",,true,true,true,true")
    e_rows = (tmp_path / "data" / "energy_data" / "energy_0.csv").read_text().split()
    assert [float(r) for r in e_rows] == energies
    assert (tmp_path / "data" / "plot_data" / "difference_0_2.csv").read_text().splitlines()[0] == "x,y,t,value"
    assert len((tmp_path / "data" / "correlation_data" / "correlation_0.csv").read_text().splitlines()) == 2


def test_wilson_sim_quirks():
    """WilsonSim ignores the frequency (Q4) and measures the plain field (Q3)."""
    lattice = Lattice3d.new(8, 8, 8)
    obs = [OutputData.new_energy_observable(0.26).set_frequency(10)]
    comp = Computation.new_wilson_sim(lattice, 0.26, AlgorithmType.new_metropolis(), 10, 25, 4, 8, obs, seed=9).run()
    assert len(obs[0].data.results()) == 25
    olat = O.Lattice([8, 8, 8])
    want = O.init_hot(olat, 9)
    wil = O.Wilson(olat, 4, 8)
    O.run(olat, want, 0.26, 9, 0, 0, 10, 0, 10, order=1, wilson=wil)
    we, _ = O.run(olat, want, 0.26, 9, 10, 0, 25, 1, 10, order=1, wilson=wil)
    assert obs[0].data.results() == we.tolist()
    assert str(comp) == "4x8 Wilson Metropolis Simulation"


def test_cluster_algorithms_are_out_of_scope():
    with pytest.raises(NotImplementedError):
        AlgorithmType.new_cluster()


def test_cpp_host_driver_matches_oracle(tmp_path):
    """lqft_sim = src/bin/sim.rs against the C++ host mirror (lattice_qft.hpp): same schedule, same CSVs."""
    import os
    import subprocess
    from lattice_qft_b200 import _build
    root = str(tmp_path)
    res = subprocess.run([_build.SIM, "8", "10", "30", "--temps", "0.2,0.26", "--energy", "10", "--corr", "4", "15",
                          "--root", root], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    rows = open(os.path.join(root, "data", "results.csv")).read().splitlines()
    assert rows[1].startswith("0,3,512,8,8,8,0.2,Metropolis Simulation,10,30,") and rows[1].endswith(",,false,true,false,true")
    assert rows[2].startswith("1,3,512,8,8,8,0.26,Metropolis Simulation,10,30,")
    olat = O.Lattice([8, 8, 8])
    for index, (temp, seed) in enumerate([(0.2, 0xC0FFEE), (0.26, 0xC0FFEE + 1)]):
        want = O.init_hot(olat, seed)
        O.run(olat, want, temp, seed, 0, 0, 10, 0, 10, order=1)
        we, _ = O.run(olat, want, temp, seed, 10, 0, 30, 10, 10, order=1)
        got = [float(v) for v in open(os.path.join(root, "data", "energy_data", f"energy_{index}.csv")).read().split()]
        assert got == we.tolist()
        corr = open(os.path.join(root, "data", "correlation_data", f"correlation_{index}.csv")).read().splitlines()
        assert len(corr) == 2 and len(corr[0].split(",")) == 8
