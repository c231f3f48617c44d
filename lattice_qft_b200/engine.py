"""Thin object wrapper over the liblqft C ABI (include/lqft.h): one `Engine` = one lqft_handle.

Host arrays are numpy; everything that touches the field runs in liblqft's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Config, RunOutputs, Schedule, check, lib


def device_count() -> int:
    return int(lib().lqft_device_count())


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    check(lib().lqft_comm_unique_id(buf))
    return bytes(buf)


def threshold_table(beta: float) -> np.ndarray:
    thr = np.zeros(_lib.MAX_THRESHOLDS, dtype=np.uint32)
    n = lib().lqft_host_threshold_table(beta, thr.ctypes.data_as(C.POINTER(C.c_uint32)), thr.size)
    if n == 0:
        raise ValueError(f"beta={beta} is not representable")
    return thr[:n].copy()


def site_word(seed: int, idx: int, colour: int, sweep: int) -> int:
    return int(lib().lqft_host_site_word(seed, idx, colour, sweep))


def pick_site(seed: int, stream: int, counter: int, rep: int, n_sites: int) -> int:
    return int(lib().lqft_host_pick_site(seed, stream, counter, rep, n_sites))


def slab(T: int, world_size: int, rank: int):
    t0, nt = C.c_uint32(), C.c_uint32()
    check(lib().lqft_host_slab(T, world_size, rank, C.byref(t0), C.byref(nt)))
    return int(t0.value), int(nt.value)


def run_measurements(first_step: int, n_steps: int, every: int) -> int:
    return int(lib().lqft_run_measurements(first_step, n_steps, every))


class Engine:
    """`n_chains` Markov chains of the height model on an X*Y*T periodic lattice, on one GPU
    (or one t-slab of it when world_size > 1)."""

    def __init__(self, x: int, y: int, t: int, n_chains: int = 1, device: int = 0, world_size: int = 1,
                 rank: int = 0, nccl_id: Optional[bytes] = None, flags: int = 0, dtype: str = "i32"):
        cfg = Config()
        cfg.x, cfg.y, cfg.t = x, y, t
        if dtype not in ("i32", "f64"):
            raise ValueError("dtype must be 'i32' or 'f64'")
        self.dtype = np.int32 if dtype == "i32" else np.float64
        cfg.dtype = _lib.LQFT_I32 if dtype == "i32" else _lib.LQFT_F64
        cfg.n_chains = n_chains
        cfg.device = device
        cfg.world_size, cfg.rank = world_size, rank
        cfg.flags = flags
        if world_size > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("world_size > 1 needs the 128-byte id of comm_unique_id()")
            C.memmove(cfg.nccl_id, nccl_id, 128)
        self._h = C.c_void_p()
        check(lib().lqft_create(C.byref(cfg), C.byref(self._h)))
        self.x, self.y, self.t, self.n_chains = x, y, t, n_chains
        self.world_size, self.rank = world_size, rank
        self.t0, self.nt = slab(t, world_size, rank)
        self.n_sites = x * y * t
        self.n_slab = x * y * self.nt
        self.beta = [None] * n_chains

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().lqft_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration ----------------------------------------------------------------------
    def set_chain(self, chain: int, beta: float, seed: int, wilson_width: int = 0, wilson_height: int = 0):
        check(lib().lqft_set_chain(self._h, chain, beta, seed & 0xFFFFFFFFFFFFFFFF, wilson_width, wilson_height))
        self.beta[chain] = beta

    def init_hot(self):
        check(lib().lqft_init_hot(self._h))

    def upload(self, chain: int, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=self.dtype).reshape(-1)
        if v.size != self.n_slab:
            raise ValueError(f"expected {self.n_slab} values, got {v.size}")
        check(lib().lqft_upload(self._h, chain, v.ctypes.data_as(C.c_void_p)))

    def download(self, chain: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_slab, dtype=self.dtype)
        assert out.dtype == self.dtype and out.size == self.n_slab and out.flags.c_contiguous
        check(lib().lqft_download(self._h, chain, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- update -----------------------------------------------------------------------------
    def sweep(self, first_sweep: int, n_sweeps: int = 1, want_accepted: bool = False):
        if want_accepted:
            acc = (C.c_uint64 * self.n_chains)()
            check(lib().lqft_sweep(self._h, first_sweep, n_sweeps, acc))
            return [int(a) for a in acc]
        check(lib().lqft_sweep(self._h, first_sweep, n_sweeps, None))
        return None

    def normalize(self, sweep_index: int):
        check(lib().lqft_normalize(self._h, sweep_index))

    def normalize_site(self, chain: int, site: int):
        check(lib().lqft_normalize_site(self._h, chain, site))

    def shift_values(self, chain: int, shift: int):
        check(lib().lqft_shift_values(self._h, chain, float(shift)))

    # -- observables ------------------------------------------------------------------------
    def energy(self):
        e_int = (C.c_int64 * self.n_chains)()
        e_obs = (C.c_double * self.n_chains)()
        check(lib().lqft_energy(self._h, e_int, e_obs))
        return np.array(e_int[:], dtype=np.int64), np.array(e_obs[:], dtype=np.float64)

    def action(self):
        s_int = (C.c_int64 * self.n_chains)()
        s_obs = (C.c_double * self.n_chains)()
        check(lib().lqft_action(self._h, s_int, s_obs))
        return np.array(s_int[:], dtype=np.int64), np.array(s_obs[:], dtype=np.float64)

    def slice_moments(self, chain: int):
        s1 = (C.c_int64 * self.t)()
        s2 = (C.c_int64 * self.t)()
        check(lib().lqft_slice_moments(self._h, chain, s1, s2))
        return np.array(s1[:], dtype=np.int64), np.array(s2[:], dtype=np.int64)

    def correlator(self, chain: int, reps: int = 0, sweep_index: int = 0,
                   ref_sites: Optional[Sequence[int]] = None) -> np.ndarray:
        out = (C.c_double * self.t)()
        if ref_sites is not None:
            sites = (C.c_uint64 * len(ref_sites))(*[int(s) for s in ref_sites])
            check(lib().lqft_correlator(self._h, chain, sites, len(ref_sites), sweep_index, out))
        else:
            check(lib().lqft_correlator(self._h, chain, None, reps, sweep_index, out))
        return np.array(out[:], dtype=np.float64)

    def difference_update(self, chain: int):
        check(lib().lqft_difference_update(self._h, chain))

    def difference_read(self, chain: int):
        mean = np.empty(3 * self.n_slab, dtype=np.float64)
        count = C.c_uint64()
        check(lib().lqft_difference_read(self._h, chain, mean.ctypes.data_as(C.POINTER(C.c_double)), C.byref(count)))
        return mean.reshape(3, self.n_slab), int(count.value)

    # -- fused schedule ---------------------------------------------------------------------
    def run(self, n_steps: int, first_step: int = 0, sweep_base: int = 0, energy_every: int = 0,
            normalize_every: int = 10, want_accepted: bool = False):
        """Loop body of Simulation::run / WilsonSim::run (computation.rs:229-270, 313-340).
        Returns (e_obs[n_chains, n_meas], e_int[n_chains, n_meas], accepted or None)."""
        s = Schedule(sweep_base, first_step, n_steps, energy_every, normalize_every, 0, 0, 0, 0)
        n_meas = run_measurements(first_step, n_steps, energy_every)
        e_obs = np.zeros((self.n_chains, n_meas), dtype=np.float64)
        e_int = np.zeros((self.n_chains, n_meas), dtype=np.int64)
        acc = (C.c_uint64 * self.n_chains)() if want_accepted else None
        check(lib().lqft_run(self._h, C.byref(s),
                             e_obs.ctypes.data_as(C.POINTER(C.c_double)) if n_meas else None,
                             e_int.ctypes.data_as(C.POINTER(C.c_int64)) if n_meas else None, acc))
        return e_obs, e_int, ([int(a) for a in acc] if want_accepted else None)

    def run_observables(self, n_steps: int, first_step: int = 0, sweep_base: int = 0, energy_every: int = 0,
                        action_every: int = 0, correlator_every: int = 0, correlator_reps: int = 0,
                        difference_every: int = 0, normalize_every: int = 10, want_accepted: bool = False) -> dict:
        """The whole measurement loop of Simulation::run (computation.rs:247-270) as ONE lqft_run_observables call:
        every OutputData kind on its own frequency, logged on the device.  Returns a dict with e_obs / e_int
        [n_chains, n_energy], s_obs / s_int [n_chains, n_action], corr [n_chains, n_corr, T], accepted."""
        s = Schedule(sweep_base, first_step, n_steps, energy_every, normalize_every, action_every, correlator_every,
                     correlator_reps, difference_every)
        n_e = run_measurements(first_step, n_steps, energy_every)
        n_a = run_measurements(first_step, n_steps, action_every)
        n_c = run_measurements(first_step, n_steps, correlator_every)
        res = {
            "e_obs": np.zeros((self.n_chains, n_e), dtype=np.float64), "e_int": np.zeros((self.n_chains, n_e), dtype=np.int64),
            "s_obs": np.zeros((self.n_chains, n_a), dtype=np.float64), "s_int": np.zeros((self.n_chains, n_a), dtype=np.int64),
            "corr": np.zeros((self.n_chains, n_c, self.t), dtype=np.float64), "accepted": None,
        }
        acc = (C.c_uint64 * self.n_chains)() if want_accepted else None
        o = RunOutputs()
        if n_e:
            o.e_obs = res["e_obs"].ctypes.data_as(C.POINTER(C.c_double))
            o.e_int = res["e_int"].ctypes.data_as(C.POINTER(C.c_int64))
        if n_a:
            o.s_obs = res["s_obs"].ctypes.data_as(C.POINTER(C.c_double))
            o.s_int = res["s_int"].ctypes.data_as(C.POINTER(C.c_int64))
        if n_c:
            o.corr = res["corr"].ctypes.data_as(C.POINTER(C.c_double))
        if acc is not None:
            o.accepted = C.cast(acc, C.POINTER(C.c_uint64))
        check(lib().lqft_run_observables(self._h, C.byref(s), C.byref(o)))
        if acc is not None:
            res["accepted"] = [int(a) for a in acc]
        return res

    # -- misc -------------------------------------------------------------------------------
    def synchronize(self):
        check(lib().lqft_synchronize(self._h))

    def last_device_ms(self) -> float:
        ms = C.c_float()
        check(lib().lqft_last_device_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        check(lib().lqft_kernel_launches(self._h, C.byref(n)))
        return int(n.value)

    def narrow_segments(self) -> int:
        n = C.c_uint64()
        check(lib().lqft_narrow_segments(self._h, C.byref(n)))
        return int(n.value)

    def plan_info(self) -> dict:
        """Launch plan of the update kernels as a dict (path, chunks, rows_per_chunk, ...)."""
        buf = C.create_string_buffer(256)
        check(lib().lqft_plan_info(self._h, buf, len(buf)))
        out = {}
        for kv in buf.value.decode().split():
            k, v = kv.split("=", 1)
            out[k] = int(v) if v.lstrip("-").isdigit() else v
        return out

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().lqft_stream(self._h, C.byref(s)))
        return int(s.value or 0)
