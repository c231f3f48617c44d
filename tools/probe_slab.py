"""Small driver for ncu: a few sweeps of one 1024 x 1024 x 128 periodic slab (BASELINE config 5's per-GPU share)."""
import sys
sys.path.insert(0, ".")
from lattice_qft_b200 import Engine

eng = Engine(1024, 1024, 128)
eng.set_chain(0, 0.25, 20261019)
eng.init_hot()
eng.sweep(0, 12)
eng.synchronize()
print(eng.plan_info())
