"""A/B of library builds (LQFT_LIB=...) on the launch-bound configs: 128^3, one and 16 / 64 chains of 64^3, the Wilson scan."""
import sys
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
out = []
for name, size, n, every, wil in (("c128", (128, 128, 128), 1, 0, False), ("c128e10", (128, 128, 128), 1, 10, False),
                                  ("c64", (64, 64, 64), 1, 10, False), ("wilson16", (64, 64, 64), 16, 1, True),
                                  ("scan16", (64, 64, 64), 16, 10, False), ("scan64", (64, 64, 64), 64, 10, False),
                                  ("c256x128", (256, 256, 128), 1, 10, False), ("c512", (512, 512, 64), 1, 10, False)):
    eng = Engine(*size, n_chains=n, flags=_lib.FLAG_NO_RESIDENT)
    for c in range(n):
        w, h = shapes[c % 16] if wil else (0, 0)
        eng.set_chain(c, 0.26 if wil else 0.25, 0xBEEF + c, w, h)
    eng.init_hot()
    eng.run(100, energy_every=0)
    sec = min(timed(eng, lambda: eng.run(200, sweep_base=100, energy_every=every), 3) for _ in range(2))
    out.append(f"{name} {n * size[0] * size[1] * size[2] * 200 / sec:.4g}")
    eng.close()
print(" | ".join(out), flush=True)
