"""Host statistics mirrored from the reference's analysis binaries (no GPU)."""
import numpy as np
import pytest

from lattice_qft_b200 import analysis


def test_symmetrize_kat():  # correlation_data_analysis.rs:550-556
    assert analysis.symmetrize_single_correlation_function([1, 2, 3, 4, 5, 6]).tolist() == [1, 4, 4, 4, 4, 4]


def test_xi_estimator_on_pure_cosh():  # lib.rs:106-120
    T, m = 32, 0.2
    c = [np.cosh(m * (t - T / 2)) for t in range(T)]
    assert analysis.second_moment_xi(c) == pytest.approx(1 / m, rel=1e-9)


def test_binned_bootstrap_recovers_mean_and_error():
    rng = np.random.default_rng(0)
    runs = [rng.normal(10.0, 2.0, size=4096) for _ in range(4)]
    mean, err = analysis.binned_bootstrap(runs, 8)
    expected = 2.0 / np.sqrt(4 * 4096)
    assert mean == pytest.approx(10.0, abs=5 * expected)
    assert 0.4 * expected < err < 2.5 * expected


def test_autocorrelation_time_of_ar1():
    rng = np.random.default_rng(1)
    rho, n = 0.8, 200000
    x = np.empty(n)
    x[0] = 0.0
    eps = rng.normal(size=n)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + eps[i]
    tau = analysis.integrated_autocorrelation_time(x)
    assert tau == pytest.approx(0.5 * (1 + rho) / (1 - rho), rel=0.1)
    assert analysis.integrated_autocorrelation_time(rng.normal(size=50000)) == pytest.approx(0.5, abs=0.1)


def test_weighted_linear_fit():
    x = [2, 5, 8, 10, 13]
    y = [3.0 * v + 680.0 for v in x]
    slope, offset = analysis.weighted_linear_fit(x, y, [0.06, 0.07, 0.07, 0.08, 0.09])
    assert slope == pytest.approx(3.0) and offset == pytest.approx(680.0)
