"""Mirror of src/computation.rs: the measurement driver (Computation / Compute::run /
ComputationSummary / parse_simulation_results).

`run()` keeps the reference's schedule (computation.rs:221-281, 304-350) — hot start, burn-in,
then per step: sweep, observables when due, normalize every 10th step — but owns a device handle
for the whole loop: the burn-in and every stretch between host-visible measurements is ONE call of
lqft_run, which replays the (sweep, energy, normalize) cycle as a CUDA graph on the device.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional

from . import export
from .algorithm import AlgorithmType, SweepRng
from .fields import Field, WilsonField
from .lattice import Lattice
from .outputdata import (ActionObservableNew, CorrelationPlotOutputData, DifferencePlotOutputData,
                         EnergyObservable, OutputData)

RESULTS_PATH = "./data/results.csv"                       # lib.rs:63
CORR_FN_PATH_INCOMPLETE = "./data/correlation_data/"      # lib.rs:66
DIFFERENCE_PLOT_PATH_INCOMPLETE = "./data/plot_data/difference_"   # lib.rs:74
ACTION_DATA_PATH_INCOMPLETE = "./data/action_data/action_"         # lib.rs:77
ENERGY_DATA_PATH_INCOMPLETE = "./data/energy_data/energy_"         # lib.rs:78

NORMALIZE_EVERY = 10  # computation.rs:242, 267, 316, 337


class Computation:
    """enum Computation (computation.rs:18-28): Simulation or WilsonSim.  The exhaustive
    `Test` / `WilsonTest` variants have no public constructor in the reference and are out of scope."""

    def __init__(self, lattice: Lattice, temp: float, algorithm: AlgorithmType, burnin: int, iterations: int,
                 output: List[OutputData], width: Optional[int] = None, height: Optional[int] = None,
                 seed: int = 0, device: int = 0):
        self.lattice, self.temp, self.algorithm = lattice, float(temp), algorithm
        self.burnin, self.iterations = int(burnin), int(iterations)
        self.output = output
        self.width, self.height = width, height
        self.seed, self.device = seed, device
        self.duration: Optional[int] = None
        self.accepted: Optional[int] = None
        self.field: Optional[Field] = None

    @staticmethod
    def new_simulation(lattice, temp, algorithm, burnin, iterations, output, seed: int = 0, device: int = 0):
        return Computation(lattice, temp, algorithm, burnin, iterations, output, None, None, seed, device)

    @staticmethod
    def new_wilson_sim(lattice, temp, algorithm, burnin, iterations, width, height, output, seed: int = 0,
                       device: int = 0):
        return Computation(lattice, temp, algorithm, burnin, iterations, output, int(width), int(height), seed,
                           device)

    @property
    def is_wilson(self) -> bool:
        return self.width is not None

    def __str__(self):  # computation.rs:32-52
        if self.is_wilson:
            return f"{self.width}x{self.height} Wilson {self.algorithm} Simulation"
        return f"{self.algorithm} Simulation"

    def into_computation(self) -> "Computation":
        return self

    def run(self) -> "Computation":
        """Simulation::run / WilsonSim::run."""
        t_start = time.time()
        field = Field.random(self.lattice, seed=self.seed, device=self.device)  # hot start, :225-226
        eng = field._to_device()
        w, h = (self.width, self.height) if self.is_wilson else (0, 0)
        field._set_params(self.temp, self.seed, w, h)
        accepted = 0
        # burn-in (:229-245 / :313-319): sweeps + normalize every 10th step, no measurements
        if self.burnin:
            _, _, acc = eng.run(self.burnin, first_step=0, sweep_base=0, energy_every=0,
                                normalize_every=NORMALIZE_EVERY, want_accepted=True)
            accepted += acc[0]
        # WilsonSim ignores `frequency` (quirk Q4, :334-336) and measures the plain field (Q3, :335)
        freq = (lambda o: 1) if self.is_wilson else (lambda o: o.frequency)
        # the first output of each kind is measured and logged on the device inside ONE lqft_run_observables call;
        # a second output of the same kind (no reference driver has one) is updated from the host between calls
        fused, others = {}, []
        for o in self.output:
            if type(o.data) not in fused:
                fused[type(o.data)] = o
            else:
                others.append(o)
        fe, fa = fused.get(EnergyObservable), fused.get(ActionObservableNew)
        fc, fd = fused.get(CorrelationPlotOutputData), fused.get(DifferencePlotOutputData)
        rng = SweepRng(self.seed)
        step, end = 0, self.iterations
        while step < end:
            nxt = end - 1
            for o in others:
                f = freq(o)
                nxt = min(nxt, ((step + f - 1) // f) * f)
            n = nxt - step + 1
            res = eng.run_observables(
                n, first_step=step, sweep_base=self.burnin,
                energy_every=freq(fe) if fe else 0, action_every=freq(fa) if fa else 0,
                correlator_every=freq(fc) if fc else 0, correlator_reps=fc.data.repetitions if fc else 0,
                difference_every=freq(fd) if fd else 0, normalize_every=NORMALIZE_EVERY, want_accepted=True)
            accepted += res["accepted"][0]
            field._device_changed()
            if fe is not None:
                fe.data.values.extend(res["e_obs"][0].tolist())
            if fa is not None:
                fa.data.values.extend(res["s_obs"][0].tolist())
            if fc is not None:
                fc.data.corr_fn.extend(res["corr"][0].tolist())
            if fd is not None:
                fd.data._field = field
            for o in others:
                if nxt % freq(o) == 0:
                    o.update(field, rng)
            step = nxt + 1
        field._device_changed()
        self.field = field
        self.accepted = accepted
        self.duration = int(time.time() - t_start)  # :272-278 (integer seconds)
        return self


def parse_simulation_results(data: List[Computation], root: str = "."):
    """computation.rs:129-199: append each run to results.csv and write its data files."""
    def p(rel: str) -> str:
        path = os.path.normpath(os.path.join(root, rel))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        return path

    for comp in data:
        rows = export.read_summaries(p(RESULTS_PATH))
        index = int(rows[-1]["index"]) + 1 if rows else 0
        x, y, t = comp.lattice.size
        summary = {
            "index": index, "d": 3, "size": comp.lattice.n_sites, "x": x, "y": y, "t": t, "temp": comp.temp,
            "comptype": str(comp), "burnin": comp.burnin, "iterations": comp.iterations,
            # WilsonSim summaries drop the duration (computation.rs:588-600)
            "duration": None if comp.is_wilson else comp.duration,
            "algorithm_statistic": None,
            "action_data": False, "energy_data": False, "difference_data": False, "correlation_data": False,
        }
        for out in comp.output:
            d = out.data
            if isinstance(d, ActionObservableNew):
                export.append_floats(p(ACTION_DATA_PATH_INCOMPLETE + f"{index}.csv"), d.results())
                summary["action_data"] = True
            elif isinstance(d, DifferencePlotOutputData):
                for direction, plot in enumerate(d.plot()):
                    export.append_plotdata(p(DIFFERENCE_PLOT_PATH_INCOMPLETE + f"{index}_{direction}.csv"), plot)
                summary["difference_data"] = True
            elif isinstance(d, CorrelationPlotOutputData):
                export.overwrite_float_rows(p(CORR_FN_PATH_INCOMPLETE + f"correlation_{index}.csv"), d.plot())
                summary["correlation_data"] = True
            elif isinstance(d, EnergyObservable):
                export.append_floats(p(ENERGY_DATA_PATH_INCOMPLETE + f"{index}.csv"), d.results())
                summary["energy_data"] = True
        export.append_summary(p(RESULTS_PATH), summary)
