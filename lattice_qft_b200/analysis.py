"""Host-side statistics the validation harness needs: the second-moment correlation
length of src/lib.rs:106-167, the symmetrisation of src/bin/correlation_data_analysis.rs:541-548, and
chain-mean errors, the error-propagated correlation length (lib.rs:122-159), the compounded binned bootstrap of the
correlation functions (correlation_data_analysis.rs:371-473) and the cosh fit (:584-712).  O(T) host arithmetic on
measured observables — nothing here touches the field."""
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


# ---- correlation length with errors (src/lib.rs:122-159) -------------------------------------------------------
def calculate_correlation_length_errors(corr_fn: Sequence[float], corr_fn_err: Sequence[float], p1: float,
                                        p2: float) -> Tuple[Tuple[float, float], Tuple[float, float]]:
    """((Re m, Im m), (err Re m, 0)): linear error propagation of the per-point errors u_j of the correlation
    function through m = acosh((g1 cos p1 - g2 cos p2) / (g1 - g2)), g_k = sum_j exp(i j p_k) c_j:
        dm/dc_j = (g1 e^{i j p2} - e^{i j p1} g2)(cos p1 - cos p2) / ((g1 - g2)^2 sqrt(z - 1) sqrt(z + 1)),
    err = sqrt(sum_j |dm/dc_j u_j|^2) — lib.rs:143-156, literally."""
    g1 = discrete_fourier_transform(corr_fn, p1)
    g2 = discrete_fourier_transform(corr_fn, p2)
    cos1, cos2 = math.cos(p1), math.cos(p2)
    y = g1 - g2
    z = (g1 * cos1 - g2 * cos2) / y
    ma = cmath.acosh(z)
    total = 0.0
    for pos, u in enumerate(corr_fn_err):
        dg1, dg2 = cmath.exp(1j * pos * p1), cmath.exp(1j * pos * p2)
        df = ((g1 * dg2 - dg1 * g2) * (cos1 - cos2)) / (y * y * cmath.sqrt(z - 1.0) * cmath.sqrt(z + 1.0))
        total += abs(df * u) ** 2
    return (ma.real, ma.imag), (math.sqrt(total), 0.0)


class Welford:
    """WelfordsAlgorithm64 (kahan.rs:78-138) on plain floats, vectorised over a trailing axis; keeps the crate's
    quirk that get_standard_error returns var_unbiased / sqrt(n) (kahan.rs:135-137), not sqrt(var / n)."""

    def __init__(self, shape=()):
        self.sum = np.zeros(shape)
        self.dss = np.zeros(shape)
        self.n = 0

    def update(self, value):
        value = np.asarray(value, dtype=np.float64)
        old = self.sum / self.n if self.n else np.zeros_like(self.sum)
        self.sum = self.sum + value
        self.n += 1
        new = self.sum / self.n
        self.dss = self.dss + (value - old) * (value - new)

    def update_many(self, rows: np.ndarray):
        """Same result as update() row by row up to rounding: mean and centred sum of squares of all rows."""
        rows = np.asarray(rows, dtype=np.float64)
        if self.n == 0:
            self.n = rows.shape[0]
            self.sum = rows.sum(axis=0)
            self.dss = ((rows - rows.mean(axis=0)) ** 2).sum(axis=0)
        else:
            for r in rows:
                self.update(r)

    def get_mean(self):
        return self.sum / self.n if self.n else np.zeros_like(self.sum)

    def get_var_unbiased(self):
        return self.dss / (self.n - 1) if self.n > 1 else np.zeros_like(self.dss)

    def get_sd_unbiased(self):
        return np.sqrt(self.get_var_unbiased())

    def get_standard_error(self):
        return self.get_var_unbiased() / math.sqrt(self.n) if self.n else np.zeros_like(self.dss)


def compounded_binned_bootstrap(simulations: Sequence[np.ndarray], bin_size: int, ensembles: int = BOOTSTRAPPING_ENSEMBLES,
                                seed0: int = 0, symmetrize: bool = True) -> dict:
    """average_over_simulation_then_build_statistics (src/bin/correlation_data_analysis.rs:371-473).

    `simulations`: one array [n_measurements, T] of correlation functions per simulation of the same parameters.
    Each is symmetrised (:541-548) and cut into bins of `bin_size` consecutive measurements (chunks_exact).  Every
    bootstrap ensemble redraws the simulations with replacement and, inside a simulation, as many bins as it has,
    with replacement; over all redrawn measurements it forms per-tau mean and `get_standard_error` (the crate's
    var/sqrt(n), see Welford), and from those xi_12 and its propagated error (calculate_correlation_length_errors,
    p1 = 2 pi / T, p2 = 2 p1).  Returned: statistics over the ensembles — corr_fn mean / sd per tau, mass
    m_12 = 1/xi_12 mean and sd, and the mean propagated error.  The reference seeds ensemble i with
    StdRng::seed_from_u64(i); numpy's generator stands in (same recipe, different stream)."""
    binned = []
    for sim in simulations:
        a = np.asarray(sim, dtype=np.float64)
        if symmetrize:
            a = np.stack([symmetrize_single_correlation_function(r) for r in a])
        n = a.shape[0] // bin_size
        if n:
            binned.append(a[: n * bin_size].reshape(n, bin_size, a.shape[1]))
    if not binned:
        raise ValueError("no simulation holds a full bin")
    t = binned[0].shape[2]
    p1 = 2.0 * math.pi / t
    fn_stats, m_stats, err_stats = Welford((t,)), Welford(), Welford()
    for e in range(ensembles):
        rng = np.random.default_rng(seed0 + e)
        rows = []
        for _ in binned:
            b = binned[rng.integers(len(binned))]
            rows.append(b[rng.integers(len(b), size=len(b))].reshape(-1, t))
        w = Welford((t,))
        w.update_many(np.concatenate(rows))
        (m, _), (m_err, _) = calculate_correlation_length_errors(w.get_mean(), w.get_standard_error(), p1, 2.0 * p1)
        fn_stats.update(w.get_mean())
        m_stats.update(m)
        err_stats.update(m_err)
    return {"corr_fn": fn_stats.get_mean(), "corr_fn_sd": fn_stats.get_sd_unbiased(),
            "m12": float(m_stats.get_mean()), "m12_sd": float(m_stats.get_sd_unbiased()),
            "m12_propagated_err": float(err_stats.get_mean()), "xi12": 1.0 / float(m_stats.get_mean()),
            "ensembles": ensembles, "bin_size": bin_size}


def cosh_fit(corr_fn: Sequence[float], m0: float = 0.1) -> dict:
    """nonlin_regression (src/bin/correlation_data_analysis.rs:635-712): y_j = max(c) - c_j is fitted by
    c0 cosh(m (j - n/2)) + c1 with n = len(c); m by Levenberg-Marquardt from m0 = 0.1, the linear coefficients by
    variable projection (the `varpro` crate's separable model: basis {cosh(m (x - n/2)), 1}).  Returns m, n, c0,
    c1 — the columns of correlation_{i}_fits.csv."""
    from scipy.optimize import least_squares
    c = np.asarray(corr_fn, dtype=np.float64)
    n = float(len(c))
    y = c.max() - c
    x = np.arange(len(c), dtype=np.float64)

    def basis(m):
        return np.stack([np.cosh(m * (x - n / 2.0)), np.ones_like(x)], axis=1)

    def residual(p):
        phi = basis(p[0])
        coeff, *_ = np.linalg.lstsq(phi, y, rcond=None)
        return phi @ coeff - y

    sol = least_squares(residual, x0=[m0], method="lm", xtol=np.finfo(np.float64).eps)
    if not sol.success:
        raise RuntimeError("termination was not successful")
    m = float(sol.x[0])
    coeff, *_ = np.linalg.lstsq(basis(m), y, rcond=None)
    return {"m": m, "n": n, "c0": float(coeff[0]), "c1": float(coeff[1])}
