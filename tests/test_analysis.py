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


def test_correlation_length_errors_matches_finite_differences():  # lib.rs:122-159
    T, m = 24, 0.3
    rng = np.random.default_rng(5)
    c = np.array([np.cosh(m * (t - T / 2)) for t in range(T)]) + rng.normal(0, 1e-3, T)
    err = rng.uniform(1e-3, 2e-3, T)
    p1 = 2 * np.pi / T
    (re, im), (re_err, im_err) = analysis.calculate_correlation_length_errors(c, err, p1, 2 * p1)
    assert (re, im) == analysis.calculate_correlation_length(c, p1, 2 * p1)
    assert im_err == 0.0
    # the formula sums |dm/dc_j * u_j|^2 with the COMPLEX derivative: check it against central differences
    total = 0.0
    for j in range(T):
        d = np.zeros(T)
        d[j] = 1e-6
        mp = complex(*analysis.calculate_correlation_length(c + d, p1, 2 * p1))
        mm = complex(*analysis.calculate_correlation_length(c - d, p1, 2 * p1))
        total += abs((mp - mm) / 2e-6 * err[j]) ** 2
    assert re_err == pytest.approx(np.sqrt(total), rel=1e-5)


def test_welford_mirror_kats():  # kahan.rs:140-168: mean, variances and the var/sqrt(n) "standard error" quirk
    w = analysis.Welford()
    for v in (2.0, 4.0, 4.0, 4.0, 5.0, 5.0, 7.0, 9.0):
        w.update(v)
    assert w.get_mean() == 5.0
    assert w.dss / w.n == 4.0                       # biased variance
    assert w.get_var_unbiased() == pytest.approx(32.0 / 7.0)
    assert w.get_standard_error() == pytest.approx(32.0 / 7.0 / np.sqrt(8.0))
    rows = np.random.default_rng(3).normal(size=(50, 4))
    a, b = analysis.Welford((4,)), analysis.Welford((4,))
    for r in rows:
        a.update(r)
    b.update_many(rows)
    assert np.allclose(a.get_mean(), b.get_mean()) and np.allclose(a.get_var_unbiased(), b.get_var_unbiased())


def test_compounded_bootstrap_recovers_mass():  # correlation_data_analysis.rs:371-473
    T, m = 16, 0.5
    rng = np.random.default_rng(11)
    clean = np.array([np.cosh(m * (t - T / 2)) for t in range(T)])
    sims = [clean + rng.normal(0, 0.02, size=(256, T)) for _ in range(3)]
    res = analysis.compounded_binned_bootstrap(sims, bin_size=4, ensembles=40)
    assert res["corr_fn"].shape == (T,) and res["ensembles"] == 40
    assert res["m12"] == pytest.approx(m, abs=5 * max(res["m12_sd"], 1e-4))
    assert res["xi12"] == pytest.approx(1 / res["m12"])
    assert res["m12_sd"] > 0 and res["m12_propagated_err"] > 0


def test_cosh_fit_recovers_parameters():  # correlation_data_analysis.rs:635-712
    T, m = 32, 0.21
    c = np.array([100.0 - 3.0 * np.cosh(m * (t - T / 2)) for t in range(T)])
    fit = analysis.cosh_fit(c)
    # y = max(c) - c = 3 cosh(m (x - n/2)) - 3 cosh(0) ... up to the constant: c0 = 3, c1 = -3 (min of cosh is 1)
    assert fit["m"] == pytest.approx(m, rel=1e-6) and fit["n"] == T
    assert fit["c0"] == pytest.approx(3.0, rel=1e-6) and fit["c1"] == pytest.approx(-3.0, rel=1e-5)
