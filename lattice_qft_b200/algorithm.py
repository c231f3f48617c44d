"""Mirror of src/algorithm.rs: the update algorithms behind `Algorithm::field_sweep` and
`WilsonAlgorithm::wilson_sweep` (algorithm.rs:51-60, 80-88).

Only Metropolis is on the hot path; the three experimental cluster algorithms
(algorithm.rs:209-498) are never used by the reference's drivers and are out of scope.
"""
from __future__ import annotations

from . import _lib
from .engine import pick_site
from .fields import Field, WilsonField


class SweepRng:
    """Stands where the reference passes `&mut ThreadRng` (computation.rs:222).

    The reference's generator is a stateful, unseeded ChaCha12; the device uses counter-based
    Philox4x32-10, so the "state" is just (seed, sweep counter): every field_sweep consumes one
    sweep index, every auxiliary pick one auxiliary index."""

    def __init__(self, seed: int = 0, sweep: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.sweep = int(sweep)
        self.aux = 0

    @staticmethod
    def default() -> "SweepRng":
        return SweepRng(0, 0)

    def next_sweep(self) -> int:
        s = self.sweep
        self.sweep += 1
        return s

    def gen_site(self, n_sites: int) -> int:
        """`rng.gen_range(0..SIZE)` (lattice.rs:121)."""
        a = self.aux
        self.aux += 1
        return pick_site(self.seed, _lib.STREAM_NORMALIZE, a, 0xFFFF, n_sites)


class Metropolis:
    """algorithm.rs:105-207.  Per-site logic identical to metropolis_single (algorithm.rs:127-158);
    sites are visited in red-black order on the device instead of `for index in 0..SIZE`."""

    def field_sweep(self, field: Field, temp: float, rng: SweepRng) -> int:
        eng = field._to_device()
        field._set_params(temp, rng.seed, 0, 0)
        acc = eng.sweep(rng.next_sweep(), 1, want_accepted=True)[0]
        field._device_changed()
        return acc

    def wilson_sweep(self, field: WilsonField, temp: float, rng: SweepRng = None) -> int:
        """algorithm.rs:165-174 creates a fresh ThreadRng per sweep; pass the chain's SweepRng."""
        rng = rng if rng is not None else field.__dict__.setdefault("_rng", SweepRng.default())
        f = field.field
        eng = f._to_device()
        f._set_params(temp, rng.seed, field.width, field.height)
        acc = eng.sweep(rng.next_sweep(), 1, want_accepted=True)[0]
        f._device_changed()
        return acc

    def __str__(self):
        return "Metropolis"


class AlgorithmType:
    """enum AlgorithmType (algorithm.rs:11-47) with its dispatch impls (:63-76, :91-103)."""

    def __init__(self, algo):
        self.algo = algo

    @staticmethod
    def new_metropolis() -> "AlgorithmType":
        return AlgorithmType(Metropolis())

    @staticmethod
    def _out_of_scope(name):
        raise NotImplementedError(
            f"{name}: the cluster algorithms (algorithm.rs:209-498) are experimental, unused by the "
            "reference's drivers, and not part of the accelerated path")

    @staticmethod
    def new_cluster():
        AlgorithmType._out_of_scope("Cluster")

    @staticmethod
    def new_newcluster():
        AlgorithmType._out_of_scope("NewCluster")

    @staticmethod
    def new_newnewcluster():
        AlgorithmType._out_of_scope("NewNewCluster")

    def field_sweep(self, field: Field, temp: float, rng: SweepRng) -> int:
        return self.algo.field_sweep(field, temp, rng)

    def wilson_sweep(self, field: WilsonField, temp: float, rng: SweepRng = None) -> int:
        return self.algo.wilson_sweep(field, temp, rng)

    def __str__(self):
        return str(self.algo)
