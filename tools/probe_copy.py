"""Host <-> device cost of lqft_upload / lqft_download at 256^3 (64 MiB of i32 each way) from pinned host memory."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from lattice_qft_b200 import Engine

eng = Engine(256, 256, 256)
eng.set_chain(0, 0.25, 1)
eng.init_hot()
host = torch.empty(256**3, dtype=torch.int32).pin_memory()
arr = host.numpy()
eng.download(0, arr); eng.synchronize()
for name, fn in (("upload", lambda: eng.upload(0, arr)), ("download", lambda: eng.download(0, arr))):
    fn(); eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    eng.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt * 1e3:.3f} ms = {arr.nbytes / dt / 1e9:.1f} GB/s")
d = torch.empty(256**3, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for name, fn in (("torch H2D", lambda: d.copy_(host, non_blocking=True)), ("torch D2H", lambda: host.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt * 1e3:.3f} ms = {arr.nbytes / dt / 1e9:.1f} GB/s")
