"""Host latency of a standalone observable call at 256^3: Engine.energy() (ctypes + numpy), the bare lqft_energy call,
lqft_synchronize on an idle stream, and the device time of the observe kernel inside a run."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
from lattice_qft_b200 import Engine
from lattice_qft_b200._lib import lib

eng = Engine(256, 256, 256)
eng.set_chain(0, 0.25, 1)
eng.init_hot()
eng.sweep(0, 50)
eng.synchronize()
e_int = (C.c_int64 * 1)(); e_obs = (C.c_double * 1)()
L = lib()
for name, fn in (("Engine.energy()", lambda: eng.energy()),
                 ("lqft_energy (bare ctypes)", lambda: L.lqft_energy(eng._h, e_int, e_obs)),
                 ("lqft_synchronize (idle)", lambda: L.lqft_synchronize(eng._h))):
    fn()
    t0 = time.perf_counter()
    n = 300
    for _ in range(n):
        fn()
    print(f"{name}: {(time.perf_counter() - t0) / n * 1e6:.1f} us per call", flush=True)
eng128 = Engine(128, 128, 128)
eng128.set_chain(0, 0.25, 1)
eng128.init_hot()
eng128.sweep(0, 50)
t0 = time.perf_counter()
for _ in range(200):
    eng128.correlator(0, reps=100, sweep_index=7)
print(f"Engine.correlator 128^3: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call")
