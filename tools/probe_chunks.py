"""Half-sweep launch geometry on small lattices: updates/s at 128^3 and 16 x 64^3 for forced y-chunk counts (LQFT_N16_CHUNKS)."""
import os, sys
sys.path.insert(0, ".")
import torch
from lattice_qft_b200 import Engine, _lib


def rate(size, n_chains, chunks):
    if chunks:
        os.environ["LQFT_N16_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("LQFT_N16_CHUNKS", None)
    eng = Engine(*size, n_chains=n_chains, flags=_lib.FLAG_NO_RESIDENT)
    for c in range(n_chains):
        eng.set_chain(c, 0.25, 7 + c)
    eng.init_hot()
    eng.run(100, energy_every=0)
    stream = torch.cuda.ExternalStream(eng.stream())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.synchronize()
    a.record(stream)
    for _ in range(3):
        eng.run(200, sweep_base=100, energy_every=0)
    b.record(stream)
    eng.synchronize()
    plan = eng.plan_info()
    eng.close()
    return n_chains * size[0] * size[1] * size[2] * 600 / (a.elapsed_time(b) * 1e-3), plan


for size, n in (((128, 128, 128), 1), ((64, 64, 64), 16), ((64, 64, 64), 64), ((256, 256, 64), 1)):
    for chunks in (0, 8, 16, 32, 64, 128):
        if chunks > size[1]:
            continue
        r, plan = rate(size, n, chunks)
        print(f"{size} x{n} chunks={chunks or 'plan'} ({plan['chunks']} x {plan['plane_blocks']} blocks, {plan['rows_per_chunk']} rows): {r:.4g} upd/s", flush=True)
