"""Host-side statistics the validation harness needs: the second-moment correlation
length of src/lib.rs:106-167, the symmetrisation of src/bin/correlation_data_analysis.rs:541-548, and
chain-mean errors.  O(T) host arithmetic on measured observables — nothing here touches the field."""
from __future__ import annotations

import cmath
import math
from typing import Sequence, Tuple

import numpy as np


def symmetrize_single_correlation_function(corr_fn: Sequence[float]) -> np.ndarray:
    """correlation_data_analysis.rs:541-548: c[i] <- (c[i] + c[T-i]) / 2 for i >= 1 (test: [1..6] -> [1,4,4,4,4,4])."""
    c = np.array(corr_fn, dtype=np.float64)
    out = c.copy()
    out[1:] = (c[1:] + c[::-1][: len(c) - 1]) / 2.0
    return out


def discrete_fourier_transform(corr_fn: Sequence[float], momentum: float) -> complex:
    """lib.rs:161-167."""
    return sum(cmath.exp(1j * pos * momentum) * g for pos, g in enumerate(corr_fn))


def calculate_correlation_length(corr_fn: Sequence[float], p1: float, p2: float) -> Tuple[float, float]:
    """lib.rs:106-120: (Re m, Im m) with cosh(m) = (g1 cos p1 - g2 cos p2) / (g1 - g2); xi = 1 / Re m."""
    g1 = discrete_fourier_transform(corr_fn, p1)
    g2 = discrete_fourier_transform(corr_fn, p2)
    cosh = (g1 * math.cos(p1) - g2 * math.cos(p2)) / (g1 - g2)
    ma = cmath.acosh(cosh)
    return ma.real, ma.imag


def second_moment_xi(corr_fn: Sequence[float]) -> float:
    """xi_12 as the analysis binary computes it: p1 = 2 pi / T, p2 = 2 p1 (correlation_data_analysis.rs:431-441)."""
    t = len(corr_fn)
    p1 = 2.0 * math.pi / t
    m, _ = calculate_correlation_length(symmetrize_single_correlation_function(corr_fn), p1, 2.0 * p1)
    return 1.0 / m


def chain_mean_and_error(per_chain_means: Sequence[float]) -> Tuple[float, float]:
    """Mean and standard error over independent chains (free of autocorrelation by construction)."""
    a = np.asarray(per_chain_means, dtype=np.float64)
    return float(a.mean()), float(a.std(ddof=1) / math.sqrt(len(a)))


def jackknife(per_chain: np.ndarray, estimator) -> Tuple[float, float]:
    """Delete-one-chain jackknife of a non-linear estimator of the chain-averaged vector."""
    n = len(per_chain)
    full = estimator(per_chain.mean(axis=0))
    loo = np.array([estimator(np.delete(per_chain, i, axis=0).mean(axis=0)) for i in range(n)])
    err = math.sqrt((n - 1) / n * float(((loo - loo.mean()) ** 2).sum()))
    return float(full), err


# ---- binned bootstrap of src/bin/energy_data_analysis.rs:238-315 ---------------------------------------
BIN_SIZES = (1, 2, 4, 8, 16, 32)      # energy_data_analysis.rs:17
BOOTSTRAPPING_ENSEMBLES = 100         # energy_data_analysis.rs:18


def binned_bootstrap(runs: Sequence[Sequence[float]], bin_size: int, ensembles: int = BOOTSTRAPPING_ENSEMBLES,
                     seed0: int = 0) -> Tuple[float, float]:
    """(energy, energy_err) of one `EnergyData` row: every run is cut into bins of `bin_size` consecutive
    measurements (chunks_exact), each bootstrap ensemble redraws runs with replacement and, inside a run,
    bins with replacement; the ensemble value is the mean over all redrawn measurements; the row is the
    mean and the unbiased standard deviation over the ensembles.  The reference seeds ensemble i with
    StdRng::seed_from_u64(i); numpy's generator stands in (same recipe, different stream)."""
    binned = []
    for run in runs:
        a = np.asarray(run, dtype=np.float64)
        n = len(a) // bin_size
        binned.append(a[: n * bin_size].reshape(n, bin_size))
    means = np.empty(ensembles)
    for e in range(ensembles):
        rng = np.random.default_rng(seed0 + e)
        total, count = 0.0, 0
        for _ in binned:
            b = binned[rng.integers(len(binned))]
            pick = b[rng.integers(len(b), size=len(b))]
            total += float(pick.sum())
            count += pick.size
        means[e] = total / count
    return float(means.mean()), float(means.std(ddof=1))


def integrated_autocorrelation_time(series: Sequence[float], c: float = 5.0) -> float:
    """tau_int with Sokal's automatic window (smallest W with W >= c * tau_int(W)); 0.5 = uncorrelated."""
    a = np.asarray(series, dtype=np.float64)
    a = a - a.mean()
    n = len(a)
    var = float(a @ a) / n
    if var == 0.0:
        return 0.5
    tau = 0.5
    for w in range(1, n // 2):
        tau += float(a[:-w] @ a[w:]) / ((n - w) * var)
        if w >= c * tau:
            break
    return max(tau, 0.5)


def autocorrelation_corrected_error(series: Sequence[float]) -> Tuple[float, float, float]:
    """(mean, error, tau_int): naive standard error inflated by sqrt(2 tau_int)."""
    a = np.asarray(series, dtype=np.float64)
    tau = integrated_autocorrelation_time(a)
    return float(a.mean()), float(a.std(ddof=1) / math.sqrt(len(a)) * math.sqrt(2.0 * tau)), tau


def weighted_linear_fit(x: Sequence[float], y: Sequence[float], err: Sequence[float]) -> Tuple[float, float]:
    """Slope and offset of `regression` in energy_data_analysis.rs:125-160, weights diag(err)^-1 as there
    (sic: not err^-2): the slope of E(width) is the string tension times L."""
    x, y, w = np.asarray(x, float), np.asarray(y, float), 1.0 / np.asarray(err, float)
    a = np.stack([x, np.ones_like(x)], axis=1)
    sol = np.linalg.solve(a.T @ (w[:, None] * a), a.T @ (w * y))
    return float(sol[0]), float(sol[1])
