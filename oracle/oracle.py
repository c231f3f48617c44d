"""ctypes binding of oracle/liblqft_oracle.so (test infrastructure; see lqft_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblqft_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lqft_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, u32, i32, i64, f64, vp = C.c_uint64, C.c_uint32, C.c_int32, C.c_int64, C.c_double, C.c_void_p
        P = C.POINTER
        sig = {
            "orc_lattice_new": (vp, [u32, P(u64)]),
            "orc_lattice_free": (None, [vp]),
            "orc_n_sites": (u64, [vp]),
            "orc_neighbours": (P(u64), [vp, u64]),
            "orc_index_from_coords": (u64, [vp, P(u64)]),
            "orc_coords_from_index": (None, [vp, u64, P(u64)]),
            "orc_philox4x32": (None, [P(u32), P(u32), P(u32)]),
            "orc_site_word": (u32, [u64, u64, u32, u64]),
            "orc_pick_site": (u64, [u64, u32, u64, u32, u64]),
            "orc_hot_value": (i32, [u64, u64]),
            "orc_init_hot": (None, [vp, vp, u64]),
            "orc_local_action": (i32, [vp, vp, u64, i32]),
            "orc_wilson_new": (vp, [vp, u64, u64]),
            "orc_wilson_free": (None, [vp]),
            "orc_wilson_is_active": (C.c_int, [vp, vp, u64, u32]),
            "orc_wilson_bond_action": (i32, [vp, vp, vp, u64, u32]),
            "orc_metropolis_single": (C.c_int, [vp, vp, vp, u64, f64, C.c_int, f64]),
            "orc_sweep": (u64, [vp, vp, vp, f64, u64, u64, C.c_int]),
            "orc_integer_energy": (i64, [vp, vp, P(i32)]),
            "orc_energy_observable": (f64, [vp, vp, f64]),
            "orc_float_action": (f64, [vp, vp]),
            "orc_action_observable": (f64, [vp, vp, f64]),
            "orc_correlator": (None, [vp, vp, vp, u64, vp]),
            "orc_shift_values": (None, [vp, u64, i32]),
            "orc_kahan_add": (None, [vp, f64]),
            "orc_kahan_add_f32": (None, [P(C.c_float), P(C.c_float), C.c_float]),
            "orc_kahan_mean": (f64, [vp]),
            "orc_welford_update": (None, [vp, f64]),
            "orc_welford_mean": (f64, [vp]),
            "orc_welford_var": (f64, [vp]),
            "orc_welford_var_unbiased": (f64, [vp]),
            "orc_welford_standard_error": (f64, [vp]),
            "orc_difference_update": (None, [vp, vp, vp]),
            "orc_difference_means": (None, [vp, vp, vp]),
            "orc_sizeof_welford": (u64, []),
            "orc_run": (u64, [vp, vp, vp, f64, u64, u64, u64, u64, u32, u32, C.c_int, vp, P(u64)]),
            "orc_bench_chains": (f64, [vp, f64, u64, u64, u32, u32, P(f64)]),
            "orc_sweep_f64": (u64, [vp, vp, vp, f64, u64, u64, C.c_int]),
            "orc_energy_f64": (f64, [vp, vp]),
            "orc_action_f64": (f64, [vp, vp]),
            "orc_correlator_f64": (None, [vp, vp, vp, u64, vp]),
            "orc_half_sweep_slab": (u64, [u32, u32, u32, u32, u32, vp, C.c_int, f64, u64, u64, u32, u32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Kahan(C.Structure):
    _fields_ = [("value", C.c_double), ("kahan", C.c_double), ("entries", C.c_uint64)]


class Welford(C.Structure):
    _fields_ = [("sum", Kahan), ("diff_square_sum", Kahan), ("entries", C.c_uint64)]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Lattice:
    """Lattice<D,SIZE> of src/lattice.rs with its neighbour table."""

    def __init__(self, size):
        self.size = [int(s) for s in size]
        self.d = len(self.size)
        arr = (C.c_uint64 * self.d)(*self.size)
        self._h = lib().orc_lattice_new(self.d, arr)
        if not self._h:
            raise ValueError("bad lattice")
        self.n_sites = int(lib().orc_n_sites(self._h))

    def __del__(self):
        try:
            lib().orc_lattice_free(self._h)
        except Exception:
            pass

    def neighbours(self, index):
        p = lib().orc_neighbours(self._h, index)
        return [int(p[i]) for i in range(2 * self.d)]

    def index_from_coords(self, coords):
        arr = (C.c_uint64 * self.d)(*coords)
        return int(lib().orc_index_from_coords(self._h, arr))

    def coords_from_index(self, index):
        arr = (C.c_uint64 * self.d)()
        lib().orc_coords_from_index(self._h, index, arr)
        return [int(v) for v in arr]


class Wilson:
    def __init__(self, lattice: Lattice, width: int, height: int):
        self.lattice = lattice
        self._h = lib().orc_wilson_new(lattice._h, width, height)
        if not self._h:
            raise ValueError("wilson needs a 3-d lattice")

    def __del__(self):
        try:
            lib().orc_wilson_free(self._h)
        except Exception:
            pass

    def is_active(self, index, direction):
        return bool(lib().orc_wilson_is_active(self.lattice._h, self._h, index, direction))

    def bond_action(self, values, index, direction):
        return int(lib().orc_wilson_bond_action(self.lattice._h, _ptr(values), self._h, index, direction))


def philox4x32(key, ctr):
    k = (C.c_uint32 * 2)(*key)
    c = (C.c_uint32 * 4)(*ctr)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(k, c, o)
    return [int(v) for v in o]


def site_word(seed, idx, colour, sweep):
    return int(lib().orc_site_word(seed, idx, colour, sweep))


def pick_site(seed, stream, counter, rep, n):
    return int(lib().orc_pick_site(seed, stream, counter, rep, n))


def hot_value(seed, idx):
    return int(lib().orc_hot_value(seed, idx))


def init_hot(lat: Lattice, seed: int) -> np.ndarray:
    v = np.empty(lat.n_sites, dtype=np.int32)
    lib().orc_init_hot(lat._h, _ptr(v), seed)
    return v


def metropolis_single(lat, values, index, temp, coin, draw, wilson=None):
    return int(lib().orc_metropolis_single(lat._h, _ptr(values), wilson._h if wilson else None, index, temp,
                                           int(coin), float(draw)))


def sweep(lat, values, temp, seed, sweep_index, order=1, wilson=None) -> int:
    """One sweep in place; order 0 = lexicographic (reference), 1 = red-black."""
    assert values.dtype == np.int32 and values.flags.c_contiguous
    return int(lib().orc_sweep(lat._h, _ptr(values), wilson._h if wilson else None, temp, seed, sweep_index, order))


def integer_energy(lat, values):
    w = C.c_int32()
    e = lib().orc_integer_energy(lat._h, _ptr(values), C.byref(w))
    return int(e), int(w.value)


def energy_observable(lat, values, temp):
    return float(lib().orc_energy_observable(lat._h, _ptr(values), temp))


def float_action(lat, values):
    return float(lib().orc_float_action(lat._h, _ptr(values)))


def action_observable(lat, values, temp):
    return float(lib().orc_action_observable(lat._h, _ptr(values), temp))


def correlator(lat, values, ref_sites) -> np.ndarray:
    sites = np.ascontiguousarray(ref_sites, dtype=np.uint64)
    out = np.zeros(lat.size[2], dtype=np.float64)
    lib().orc_correlator(lat._h, _ptr(values), _ptr(sites), len(sites), _ptr(out))
    return out


def shift_values(values, shift):
    lib().orc_shift_values(_ptr(values), values.size, int(shift))


def run(lat, values, temp, seed, sweep_base, first_step, n_steps, energy_every, normalize_every, order=1,
        wilson=None):
    """Loop body of Simulation::run / WilsonSim::run; returns (energies, accepted)."""
    n_meas = 0
    if energy_every:
        last = first_step + n_steps - 1
        n_meas = max(0, last // energy_every - (first_step + energy_every - 1) // energy_every + 1) if n_steps else 0
    e = np.zeros(max(n_meas, 1), dtype=np.float64)
    acc = C.c_uint64()
    n = lib().orc_run(lat._h, _ptr(values), wilson._h if wilson else None, temp, seed, sweep_base, first_step,
                      n_steps, energy_every, normalize_every, order, _ptr(e), C.byref(acc))
    assert n == n_meas, (n, n_meas)
    return e[:n_meas].copy(), int(acc.value)


def difference_update(lat, values, wel: np.ndarray):
    lib().orc_difference_update(lat._h, _ptr(values), _ptr(wel))


def new_welford_cells(n: int) -> np.ndarray:
    assert lib().orc_sizeof_welford() == C.sizeof(Welford)
    return np.zeros(n * C.sizeof(Welford), dtype=np.uint8)


def difference_means(lat, wel: np.ndarray) -> np.ndarray:
    out = np.zeros(lat.d * lat.n_sites, dtype=np.float64)
    lib().orc_difference_means(lat._h, _ptr(wel), _ptr(out))
    return out


def bench_chains(lat, temp, seed, sweeps, energy_every, threads):
    """Timed CPU baseline: `threads` independent lexicographic chains; returns (seconds, last energy)."""
    e = C.c_double()
    secs = lib().orc_bench_chains(lat._h, temp, seed, sweeps, energy_every, threads, C.byref(e))
    return float(secs), float(e.value)


def half_sweep_slab(X, Y, T, t0, tl, values_with_ghosts, colour, temp, seed, sweep_index, wilson=(0, 0)) -> int:
    """One colour of one sweep on a t-slab stored with one ghost plane on each side (in place)."""
    v = values_with_ghosts
    assert v.dtype == np.int32 and v.flags.c_contiguous and v.size == X * Y * (tl + 2)
    return int(lib().orc_half_sweep_slab(X, Y, T, t0, tl, _ptr(v), colour, temp, seed, sweep_index, wilson[0], wilson[1]))


def sweep_f64(lat, values, temp, seed, sweep_index, order=1, wilson=None) -> int:
    assert values.dtype == np.float64 and values.flags.c_contiguous
    return int(lib().orc_sweep_f64(lat._h, _ptr(values), wilson._h if wilson else None, temp, seed, sweep_index, order))


def energy_f64(lat, values) -> float:
    return float(lib().orc_energy_f64(lat._h, _ptr(values)))


def action_f64(lat, values) -> float:
    return float(lib().orc_action_f64(lat._h, _ptr(values)))


def correlator_f64(lat, values, ref_sites) -> np.ndarray:
    sites = np.ascontiguousarray(ref_sites, dtype=np.uint64)
    out = np.zeros(lat.size[2], dtype=np.float64)
    lib().orc_correlator_f64(lat._h, _ptr(values), _ptr(sites), len(sites), _ptr(out))
    return out
