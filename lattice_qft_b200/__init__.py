"""lattice_qft_b200 — B200-native (sm_100a) Monte-Carlo engine for the 3-D discrete-Gaussian height
model of TimeTeleporter/lattice_qft, behind the reference's own interface:

    lattice.Lattice3d, fields.Field / WilsonField, algorithm.AlgorithmType / Metropolis,
    outputdata.OutputData, computation.Computation / parse_simulation_results, export

All field arithmetic runs in liblqft.so (CUDA, built by __graft_entry__.build()); importing the
package fails if the library has not been built — there is no CPU or PyTorch fallback.
"""
from ._lib import LqftError, lib  # noqa: F401

lib()  # fail loudly at import time if liblqft.so is missing or lacks a symbol

from . import algorithm, computation, engine, export, fields, lattice, outputdata  # noqa: E402,F401
from .engine import Engine  # noqa: E402,F401
