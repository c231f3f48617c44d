"""Mirror of src/outputdata.rs: observables updated from the field during a run."""
from __future__ import annotations

from typing import List

import numpy as np

from . import _lib
from .engine import pick_site
from .fields import Field
from .lattice import Lattice


class EnergyObservable:
    """outputdata.rs:152-184."""

    def __init__(self, temp: float):
        self.values: List[float] = []
        self.temp = temp

    def update(self, field: Field, rng=None):
        self.values.append(field.energy_observable(self.temp))

    def results(self) -> List[float]:
        return self.values


class ActionObservableNew:
    """outputdata.rs:186-218."""

    def __init__(self, temp: float):
        self.values: List[float] = []
        self.temp = temp

    def update(self, field: Field, rng=None):
        self.values.append(field.action_observable(self.temp))

    def results(self) -> List[float]:
        return self.values


class DifferencePlotOutputData:
    """outputdata.rs:220-300: running per-bond mean of bond_difference (Welford over Kahan sums,
    kahan.rs:78-107), accumulated on the device."""

    def __init__(self, lattice: Lattice):
        self.lattice = lattice
        self._field = None

    def update(self, field: Field, rng=None):
        if self._field is None:
            self._field = field
        elif self._field is not field:
            raise ValueError("a DifferencePlot accumulates one chain")
        field._to_device().difference_update(0)

    def plot(self):
        """Vec<Vec<Plotdata3d>>: per direction, rows (x, y, t, value) in index order."""
        if self._field is None:
            mean = np.zeros((3, self.lattice.n_sites))
        else:
            mean, _ = self._field._to_device().difference_read(0)
        X, Y, _T = self.lattice.size
        idx = np.arange(self.lattice.n_sites)
        x, y, t = idx % X, (idx // X) % Y, idx // (X * Y)
        return [list(zip(x.tolist(), y.tolist(), t.tolist(), mean[d].tolist())) for d in range(3)]


class CorrelationPlotOutputData:
    """outputdata.rs:278-351: C(t) = sum_rep sum_y (h_x - h_y)^2 binned by time distance."""

    def __init__(self, lattice: Lattice, repetitions: int):
        self.lattice = lattice
        self.corr_fn: List[List[float]] = []
        self.repetitions = int(repetitions)

    def update(self, field: Field, rng=None):
        eng = field._to_device()
        if rng is not None:
            sites = [rng.gen_site(self.lattice.n_sites) for _ in range(self.repetitions)]
        else:
            n = len(self.corr_fn)
            sites = [pick_site(0, _lib.STREAM_CORRELATOR, n, r, self.lattice.n_sites)
                     for r in range(self.repetitions)]
        self.corr_fn.append(eng.correlator(0, ref_sites=sites).tolist())

    def plot(self) -> List[List[float]]:
        return self.corr_fn


class OutputData:
    """outputdata.rs:42-119."""

    def __init__(self, data, frequency: int = 1):
        self.data = data
        self.frequency = frequency

    @staticmethod
    def new_action_observable(temp: float) -> "OutputData":
        return OutputData(ActionObservableNew(temp))

    @staticmethod
    def new_difference_plot(lattice: Lattice) -> "OutputData":
        return OutputData(DifferencePlotOutputData(lattice))

    @staticmethod
    def new_correlation_plot(lattice: Lattice, repetitions: int) -> "OutputData":
        return OutputData(CorrelationPlotOutputData(lattice, repetitions))

    @staticmethod
    def new_energy_observable(temp: float) -> "OutputData":
        return OutputData(EnergyObservable(temp))

    def set_frequency(self, frequency: int) -> "OutputData":
        self.frequency = int(frequency)
        return self

    def is_step_for_next_update(self, step: int) -> bool:
        return step % self.frequency == 0

    def update(self, field: Field, rng=None):
        self.data.update(field, rng)

    def clone(self) -> "OutputData":
        import copy
        return copy.deepcopy(self)
