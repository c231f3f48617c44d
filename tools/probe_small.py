"""Small driver for ncu: a few steps of the secondary bench configs (64^3 x 64 chains, Wilson 16 x 64^3 with the energy
after every sweep, 128^3), eager launches so that every kernel shows up in the launch list."""
import sys
sys.path.insert(0, ".")
from lattice_qft_b200 import Engine, _lib

mode = sys.argv[1]
flags = _lib.FLAG_NO_GRAPH
if mode == "scan":
    eng = Engine(64, 64, 64, n_chains=64, flags=flags)
    for c in range(64):
        eng.set_chain(c, 0.20 + 0.005 * c, 0xC0FFEE + c)
    eng.init_hot()
    eng.run(20, energy_every=10)
elif mode == "wilson":
    shapes = [(w, h) for h in (16, 32, 64) for w in (10, 21, 32, 42, 53)] + [(0, 0)]
    eng = Engine(64, 64, 64, n_chains=len(shapes), flags=flags)
    for c, (w, h) in enumerate(shapes):
        eng.set_chain(c, 0.26, 0xBEEF + c, w, h)
    eng.init_hot()
    eng.run(12, energy_every=1)
elif mode == "c128":
    eng = Engine(128, 128, 128, flags=flags)
    eng.set_chain(0, 0.25, 0xABCD)
    eng.init_hot()
    eng.run(20, energy_every=0)
    eng.run_observables(20, sweep_base=20, correlator_every=10, correlator_reps=100)
eng.synchronize()
print(mode, eng.plan_info())
