"""Property tests of the CPU oracle (hypothesis): the algebra the GPU parity tests lean on, checked against
scikit-learn -- the library whose calls ARE the reference's arithmetic (stage_1_train_model.py:81-83, 98-107)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from sklearn.linear_model import LinearRegression, Ridge
from sklearn.metrics import max_error, mean_absolute_percentage_error, r2_score

from bodywork_mlops_demo_b200 import stage_1_train_model as s1
from oracle import ols_oracle as orc

SETTINGS = dict(max_examples=25, deadline=None)


@settings(**SETTINGS)
@given(n=st.integers(12, 400), d=st.integers(1, 9), seed=st.integers(0, 10_000), alpha=st.sampled_from([0.0, 0.5, 30.0]))
def test_fit_from_stats_equals_sklearn_for_random_problems(n, d, seed, alpha):
    if n < 3 * d + 3:
        n = 3 * d + 3
    X, y = orc.generate_dataset(n, d, seed=seed)
    fo = orc.fit_from_stats(orc.gram_stats(X, y), alpha=alpha)
    ref = (Ridge(alpha=alpha, solver="cholesky") if alpha > 0 else LinearRegression()).fit(X, y)
    assert np.max(np.abs(fo["coef"] - ref.coef_)) < 1e-8
    assert abs(fo["intercept"] - ref.intercept_) < 1e-6


@settings(**SETTINGS)
@given(n=st.integers(2, 300), seed=st.integers(0, 10_000), zeros=st.integers(0, 3))
def test_score_stats_reproduce_the_three_sklearn_metrics(n, seed, zeros):
    """model_metrics (stage_1_train_model.py:79-90) = MAPE, r2_score, max_error -- including labels equal to 0,
    where sklearn clamps the denominator to eps."""
    rng = np.random.RandomState(seed)
    y = rng.uniform(-50, 80, n)
    y[:min(zeros, n)] = 0.0
    p = y + rng.normal(0, 5, n)
    mape, r2, mx = s1.metrics_from_stats(orc.score_stats(y, p))
    assert mape == pytest.approx(mean_absolute_percentage_error(y, p), rel=1e-10)
    assert mx == pytest.approx(max_error(y, p), rel=1e-12)
    if np.ptp(y) > 0:
        assert r2 == pytest.approx(r2_score(y, p), rel=1e-9, abs=1e-9)


@settings(**SETTINGS)
@given(n=st.integers(4, 500), d=st.integers(1, 6), seed=st.integers(0, 10_000), cut=st.floats(0.1, 0.9))
def test_statistic_is_additive_and_a_mask_is_a_gather(n, d, seed, cut):
    """S(A u B) = S(A) + S(B); S(rows with mask == 1) + S(rows with mask == 0) = S(all rows): what row sharding
    (one all-reduce of S) and the train / hold-out row mask rely on."""
    X, y = orc.generate_dataset(n, d, seed=seed)
    k = int(n * cut)
    whole = orc.gram_stats(X, y)
    assert np.allclose(orc.gram_stats(X[:k], y[:k]) + orc.gram_stats(X[k:], y[k:]), whole, rtol=1e-12, atol=1e-9)
    mask = s1.split_mask(n)
    both = orc.gram_stats(X[mask == 1], y[mask == 1]) + orc.gram_stats(X[mask == 0], y[mask == 0])
    assert np.allclose(both, whole, rtol=1e-12, atol=1e-9)
    assert whole[d, d] == n and np.array_equal(whole, whole.T)


@settings(**SETTINGS)
@given(n=st.integers(5, 2000))
def test_split_mask_sizes_follow_train_test_split(n):
    """test = ceil(0.2 n) rows, train = the rest (sklearn _split.py via stage_1_train_model.py:98-103)."""
    mask = s1.split_mask(n)
    assert int((mask == 0).sum()) == int(np.ceil(0.2 * n))
    assert int((mask == 1).sum()) == n - int(np.ceil(0.2 * n))
