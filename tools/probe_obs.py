"""Small driver for ncu: in-run energy measurements at 256^3 on the 16-bit copy, and a self-halo slab run."""
import sys
sys.path.insert(0, ".")
from lattice_qft_b200 import Engine, _lib

mode = sys.argv[1]
if mode == "energy":
    eng = Engine(256, 256, 256)
    eng.set_chain(0, 0.25, 1)
    eng.init_hot()
    eng.run(8, energy_every=1, normalize_every=10)
    eng.energy()
    eng.correlator(0, reps=100, sweep_index=3)
    eng.synchronize()
elif mode == "selfhalo":
    eng = Engine(256, 256, 256, flags=_lib.FLAG_SELF_HALO | _lib.FLAG_NO_GRAPH)
    eng.set_chain(0, 0.25, 1)
    eng.init_hot()
    eng.sweep(0, 12)
    eng.synchronize()
