"""Times the secondary configs on the cluster-resident path and on the launch path (LQFT_FLAG_NO_RESIDENT) side by side."""
import sys, time
sys.path.insert(0, ".")
import torch
from lattice_qft_b200 import Engine, _lib


def timed(eng, fn, reps):
    stream = torch.cuda.ExternalStream(eng.stream())
    fn(); eng.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream); eng.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


shapes = [(w, h) for h in (16, 32, 64) for w in (10, 21, 32, 42, 53)] + [(0, 0)]
for name, n, every, wil in (("scan1", 1, 10, False), ("scan2", 2, 10, False), ("scan4", 4, 10, False), ("scan8", 8, 10, False), ("wilson8", 8, 1, True), ("scan12", 12, 10, False), ("wilson16", 16, 1, True), ("scan16", 16, 10, False), ("scan32", 32, 10, False), ("scan64", 64, 10, False)):
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        eng = Engine(64, 64, 64, n_chains=n, flags=flags)
        for c in range(n):
            w, h = shapes[c % 16] if wil else (0, 0)
            eng.set_chain(c, 0.26 if wil else 0.20 + 0.005 * c, 0xBEEF + c, w, h)
        eng.init_hot()
        eng.run(100, energy_every=0)
        sec = timed(eng, lambda: eng.run(100, sweep_base=100, energy_every=every), 3)
        print(f"{name} flags={flags} path={eng.plan_info()['path']} {n * 64**3 * 100 / sec:.4g} upd/s", flush=True)
        eng.close()
for size in ((128, 64, 64), (64, 64, 32), (32, 64, 64), (128, 128, 32)):
    for flags in (0, _lib.FLAG_NO_RESIDENT):
        eng = Engine(*size, n_chains=1, flags=flags)
        eng.set_chain(0, 0.25, 7)
        eng.init_hot()
        eng.run(100, energy_every=0)
        sec = timed(eng, lambda: eng.run(100, sweep_base=100, energy_every=10), 3)
        print(f"{size} flags={flags} path={eng.plan_info()['path']} {size[0] * size[1] * size[2] * 100 / sec:.4g} upd/s", flush=True)
        eng.close()
