"""ctypes binding of liblqft.so (include/lqft.h).  There is no fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQFT_LIB") or os.path.join(PKG, "liblqft.so")  # LQFT_LIB: tuning builds

LQFT_OK, LQFT_ERR_INVALID, LQFT_ERR_CUDA, LQFT_ERR_NCCL, LQFT_ERR_STATE, LQFT_ERR_NOMEM = range(6)
LQFT_I32, LQFT_F64 = 0, 1
FLAG_NONE, FLAG_NO_GRAPH, FLAG_FORCE_GENERIC, FLAG_HALO_NCCL, FLAG_NO_PDL, FLAG_SELF_HALO, FLAG_NO_RESIDENT, FLAG_NO_NARROW = 0, 1, 2, 4, 32, 64, 128, 256
FLAG_I32_VECTOR = FLAG_NO_RESIDENT | FLAG_NO_NARROW  # the i32 vector kernel (X % 8 == 0) instead of the one-launch / 16-bit paths
FLAG_NO_FUSED_ENERGY = 512
STREAM_COLOUR0, STREAM_COLOUR1, STREAM_NORMALIZE, STREAM_HOTSTART, STREAM_CORRELATOR = range(5)
MAX_THRESHOLDS = 2048


class LqftError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"lqft status {status}: {message}")
        self.status = status
        self.message = message


class Config(C.Structure):
    _fields_ = [
        ("x", C.c_uint32), ("y", C.c_uint32), ("t", C.c_uint32), ("dtype", C.c_uint32),
        ("n_chains", C.c_uint32), ("device", C.c_int32), ("world_size", C.c_uint32), ("rank", C.c_uint32),
        ("flags", C.c_uint32), ("reserved", C.c_uint32), ("nccl_id", C.c_uint8 * 128),
    ]


class Schedule(C.Structure):
    _fields_ = [
        ("sweep_base", C.c_uint64), ("first_step", C.c_uint64), ("n_steps", C.c_uint64),
        ("energy_every", C.c_uint32), ("normalize_every", C.c_uint32),
        ("action_every", C.c_uint32), ("correlator_every", C.c_uint32), ("correlator_reps", C.c_uint32),
        ("difference_every", C.c_uint32),
    ]


class RunOutputs(C.Structure):
    _fields_ = [
        ("e_obs", C.POINTER(C.c_double)), ("e_int", C.POINTER(C.c_int64)),
        ("s_obs", C.POINTER(C.c_double)), ("s_int", C.POINTER(C.c_int64)),
        ("corr", C.POINTER(C.c_double)), ("accepted", C.POINTER(C.c_uint64)),
    ]


# every symbol include/lqft.h declares: name -> (restype, argtypes)
u8p, u32, u64, i32, i64, f64, vp = C.POINTER(C.c_uint8), C.c_uint32, C.c_uint64, C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER
SIGNATURES = {
    "lqft_version": (u32, []),
    "lqft_last_error": (C.c_char_p, []),
    "lqft_device_count": (i32, []),
    "lqft_comm_unique_id": (C.c_int, [u8p]),
    "lqft_create": (C.c_int, [P(Config), P(vp)]),
    "lqft_destroy": (C.c_int, [vp]),
    "lqft_set_chain": (C.c_int, [vp, u32, f64, u64, u32, u32]),
    "lqft_init_hot": (C.c_int, [vp]),
    "lqft_upload": (C.c_int, [vp, u32, vp]),
    "lqft_download": (C.c_int, [vp, u32, vp]),
    "lqft_sweep": (C.c_int, [vp, u64, u32, P(u64)]),
    "lqft_normalize": (C.c_int, [vp, u64]),
    "lqft_normalize_site": (C.c_int, [vp, u32, u64]),
    "lqft_shift_values": (C.c_int, [vp, u32, f64]),
    "lqft_energy": (C.c_int, [vp, P(i64), P(f64)]),
    "lqft_action": (C.c_int, [vp, P(i64), P(f64)]),
    "lqft_slice_moments": (C.c_int, [vp, u32, P(i64), P(i64)]),
    "lqft_correlator": (C.c_int, [vp, u32, P(u64), u32, u64, P(f64)]),
    "lqft_difference_update": (C.c_int, [vp, u32]),
    "lqft_difference_read": (C.c_int, [vp, u32, P(f64), P(u64)]),
    "lqft_run_measurements": (u64, [u64, u64, u32]),
    "lqft_run": (C.c_int, [vp, P(Schedule), P(f64), P(i64), P(u64)]),
    "lqft_run_observables": (C.c_int, [vp, P(Schedule), P(RunOutputs)]),
    "lqft_synchronize": (C.c_int, [vp]),
    "lqft_last_device_ms": (C.c_int, [vp, P(C.c_float)]),
    "lqft_kernel_launches": (C.c_int, [vp, P(u64)]),
    "lqft_narrow_segments": (C.c_int, [vp, P(u64)]),
    "lqft_plan_info": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
    "lqft_stream": (C.c_int, [vp, P(vp)]),
    "lqft_host_philox4x32": (None, [u64, P(u32), P(u32)]),
    "lqft_host_site_word": (u32, [u64, u64, u32, u64]),
    "lqft_host_threshold_table": (u32, [f64, P(u32), u32]),
    "lqft_host_pick_site": (u64, [u64, u32, u64, u32, u64]),
    "lqft_host_hot_value": (i32, [u64, u64]),
    "lqft_host_index": (u64, [u32, u32, u32, u32, u32, u32]),
    "lqft_host_coords": (None, [u32, u32, u32, u64, P(u32)]),
    "lqft_host_neighbours": (None, [u32, u32, u32, u64, P(u64)]),
    "lqft_host_wilson_offset": (i32, [u32, u32, u32, u32, u32, u32]),
    "lqft_host_slab": (C.c_int, [u32, u32, u32, P(u32), P(u32)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load liblqft.so; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  lattice_qft_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != LQFT_OK:
        raise LqftError(status, lib().lqft_last_error().decode("utf-8", "replace"))
