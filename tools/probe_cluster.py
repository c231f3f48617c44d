"""Small driver for ncu: one k_run_cluster launch (Wilson sheets on 12 chains of 64^3, energy after every sweep)."""
import sys
sys.path.insert(0, ".")
from lattice_qft_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shapes = [(w, h) for h in (16, 32, 64) for w in (10, 21, 32, 42, 53)] + [(0, 0)]
eng = Engine(64, 64, 64, n_chains=n)
for c in range(n):
    w, h = shapes[c % len(shapes)]
    eng.set_chain(c, 0.26, 0xBEEF + c, w, h)
eng.init_hot()
eng.run(40, energy_every=1)
eng.synchronize()
print(eng.plan_info(), eng.last_device_ms())
