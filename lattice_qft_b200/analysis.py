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
