"""Round-2 GPU parity tests (through the C-ABI): the one-call fit (b2_fit: finalize + peer scatter + gather fused), the peer-memory exchange between
two contexts (incl. its failure path), the eigenvalue-only spectrum, b2_metrics on float64 vectors, the device tranche
generator (stage_3's y >= 0 filter and alpha(day)), NCCL entry points on a one-rank communicator, the staging-ring
write-after-read fix and estimator isolation.

Tolerances as in test_gpu_parity.py: tensor-core coefficients asserted at 2e-5 (contract 1e-4) against the fp64 oracle
of the same rows; statistics of the exact kernels 1e-12; eigenvalues 1e-9 relative to the largest.
"""
import threading

import numpy as np
import pytest

import bodywork_mlops_demo_b200 as b2
from bodywork_mlops_demo_b200 import sharding
from oracle import ols_oracle as orc

pytestmark = pytest.mark.gpu

COEF_TOL = 2e-5
INTERCEPT_TOL = 3e-2


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def _oracle_fit(X, y, mask=None, keep=1, alpha=0.0):
    if mask is not None:
        X, y = X[mask == keep], y[mask == keep]
    return orc.fit_from_stats(orc.gram_stats(X.astype(np.float64), y.astype(np.float64)), alpha=alpha)


# ------------------------------------------------------------------------------------------------
# b2_fit: fused path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,kind,masked", [
    (4096, 128, "f32", False), (100_003, 128, "f32", False), (65_536, 128, "f32", True), (300_000, 128, "bf16", False),
    (50_000, 32, "f32", False), (50_001, 32, "f32", True), (70_007, 24, "f32", False), (40_000, 40, "bf16", True),
    (33_000, 100, "f32", False), (20_001, 64, "f32", False), (2048, 20, "f32", False), (30_011, 48, "bf16", False)])
def test_fused_fit_matches_oracle_and_takes_four_launches(ctx, n, d, kind, masked):
    X, y = orc.generate_dataset(n, d, seed=n % 97 + d, dtype=np.float32)
    mask = (np.random.RandomState(d).rand(n) < 0.8).astype(np.uint8) if masked else None
    if kind == "bf16":
        Xb = b2.native.to_bf16_bits(X)
        X = b2.native.from_bf16_bits(Xb)
        Xd = ctx.to_device(Xb, "bf16")
    else:
        Xd = ctx.to_device(X)
    yd = ctx.to_device(y)
    md = ctx.to_device(mask) if masked else None
    ctx.set_kernel(b2.KERNEL_TCGEN05)
    try:
        before = ctx.stats()
        coef, b0 = ctx.fit(Xd, yd, md, 1)
        after = ctx.stats()
        S = ctx.gram_export()
    finally:
        ctx.set_kernel(b2.KERNEL_AUTO)
    assert after["fused_fits"] == before["fused_fits"] + 1
    if d > 64:                                               # no packing leftovers: shift sample, Gram, finalize, solve
        assert after["launches"] - before["launches"] == 4
    ref = _oracle_fit(X, y, mask)
    assert np.max(np.abs(coef - ref["coef"])) < COEF_TOL
    assert abs(b0 - ref["intercept"]) < INTERCEPT_TOL
    n_used = int(mask.sum()) if masked else n
    assert S[d, d] == n_used                                  # the row count is exact
    assert np.array_equal(S, S.T)
    full = orc.gram_stats(X[mask == 1] if masked else X, y[mask == 1] if masked else y)
    assert _rel(S, full) < 2e-6
    for a in (Xd, yd, md):
        if a is not None:
            a.free()


def test_fused_fit_is_bit_deterministic_and_agrees_with_the_four_call_sequence(ctx):
    n, d = 150_000, 128
    X, y = orc.generate_dataset(n, d, seed=5, dtype=np.float32)
    Xd, yd = ctx.to_device(X), ctx.to_device(y)
    ctx.set_kernel(b2.KERNEL_TCGEN05)
    try:
        c1, b1 = ctx.fit(Xd, yd); S1 = ctx.gram_export()
        c2, b2_ = ctx.fit(Xd, yd); S2 = ctx.gram_export()
        ctx.gram_reset(d); ctx.gram_accumulate(Xd, yd); c3, b3 = ctx.solve(); S3 = ctx.gram_export()
    finally:
        ctx.set_kernel(b2.KERNEL_AUTO)
    assert np.array_equal(S1, S2) and np.array_equal(c1, c2) and b1 == b2_
    # the same kernels in the same order: the one-call fit and the four-call sequence agree bit for bit
    assert np.array_equal(S1, S3) and np.array_equal(c1, c3) and b1 == b3
    Xd.free(); yd.free()


@pytest.mark.parametrize("i", range(32))
def test_randomized_one_call_fits(ctx, i):
    """b2_fit through the AUTO dispatch on random shapes / storage types / masks / ridge terms / intercept settings,
    device and host rows: coefficients of the oracle's fit of the same (kept) rows."""
    rng = np.random.RandomState(5000 + i)
    d = int(rng.choice([1, 3, 8, 16, 20, 24, 32, 36, 48, 64, 72, 100, 128]))
    n = int(rng.randint(max(2500, 40 * d), 120_000))
    kind = "bf16" if (d % 8 == 0 and rng.rand() < 0.3) else "f32"
    masked = rng.rand() < 0.4
    alpha = float(rng.choice([0.0, 0.0, 1.0, 100.0]))
    fit_intercept = bool(rng.rand() < 0.8)
    host = kind == "f32" and rng.rand() < 0.25
    X, y = orc.generate_dataset(n, d, seed=6000 + i, dtype=np.float32)
    mask = (rng.rand(n) < 0.75).astype(np.uint8) if masked else None
    if kind == "bf16":
        bits = b2.native.to_bf16_bits(X)
        X = b2.native.from_bf16_bits(bits)
    if host:
        coef, b0 = ctx.fit(X, y, mask, 1, alpha=alpha, fit_intercept=fit_intercept)
    else:
        Xd = ctx.to_device(bits, "bf16") if kind == "bf16" else ctx.to_device(X)
        yd = ctx.to_device(y)
        md = ctx.to_device(mask) if masked else None
        coef, b0 = ctx.fit(Xd, yd, md, 1, alpha=alpha, fit_intercept=fit_intercept)
        for a in (Xd, yd, md):
            if a is not None:
                a.free()
    sel = slice(None) if mask is None else (mask == 1)
    ref = orc.fit_from_stats(orc.gram_stats(X[sel].astype(np.float64), y[sel].astype(np.float64)), alpha=alpha,
                             fit_intercept=fit_intercept)
    n_used = int(mask.sum()) if masked else n
    tol = COEF_TOL * max(1.0, 3000.0 / n_used) ** 0.5 * (4 if d > 64 else 1)
    if not fit_intercept:
        tol = max(tol, 1e-3)   # the uncentred Gram of U(0,100) columns has condition ~ 1 + 3 D: outside the 1e-4 contract,
                               # which is stated for the reference's fit_intercept=True
    assert np.max(np.abs(coef - ref["coef"])) < tol, (n, d, kind, masked, alpha, fit_intercept, host)
    if fit_intercept:
        assert abs(b0 - ref["intercept"]) < INTERCEPT_TOL * max(1.0, d / 32)
    else:
        assert b0 == 0.0


def test_fit_entry_point_on_every_other_path_equals_the_sequence(ctx):
    """b2_fit outside the fused conditions (narrow rows, tiny tranche, host rows, forced SIMT) = the four calls."""
    for n, d, host in ((1440, 1, False), (50_000, 1, False), (30_000, 8, False), (5000, 37, False), (300_000, 32, True)):
        X, y = orc.generate_dataset(n, d, seed=n + d, dtype=np.float32)
        if host:
            coef, b0 = ctx.fit(X, y)
        else:
            Xd, yd = ctx.to_device(X), ctx.to_device(y)
            coef, b0 = ctx.fit(Xd, yd)
            Xd.free(); yd.free()
        ref = _oracle_fit(X, y)
        assert np.max(np.abs(coef - ref["coef"])) < COEF_TOL, (n, d)
        assert abs(b0 - ref["intercept"]) < INTERCEPT_TOL
    ctx.set_kernel(b2.KERNEL_SIMT)
    try:
        X, y = orc.generate_dataset(20_000, 128, seed=3, dtype=np.float32)
        Xd, yd = ctx.to_device(X), ctx.to_device(y)
        coef, b0 = ctx.fit(Xd, yd, alpha=10.0)
        ref = _oracle_fit(X, y, alpha=10.0)
        assert np.max(np.abs(coef - ref["coef"])) < 1e-9
        Xd.free(); yd.free()
    finally:
        ctx.set_kernel(b2.KERNEL_AUTO)


# ------------------------------------------------------------------------------------------------
# peer-memory exchange between two contexts of one process (runs on a single GPU)
# ------------------------------------------------------------------------------------------------
def _two_contexts():
    n_dev = b2.native.device_count()
    return b2.Context(0), b2.Context(1 if n_dev > 1 else 0)


def _run_both(fns):
    out, err = [None] * len(fns), [None] * len(fns)

    def work(i):
        try:
            out[i] = fns[i]()
        except Exception as exc:  # noqa: BLE001
            err[i] = exc
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    return out, err


def test_peer_exchange_fused_and_standalone_between_two_contexts():
    n, d = 200_003, 128
    X, y = orc.generate_dataset(n, d, seed=31, dtype=np.float32)
    full = orc.gram_stats(X, y)
    ref = orc.fit_from_stats(full)
    cs = _two_contexts()
    try:
        b2.Context.comm_p2p_attach_local(cs)
        assert cs[0].comm_info() == {"n_ranks": 2, "rank": 0, "exchange": "p2p"}
        assert cs[1].comm_info()["rank"] == 1
        shards = [sharding.shard_bounds(n, 2, r) for r in range(2)]
        dev = [(c.to_device(X[lo:hi]), c.to_device(y[lo:hi])) for c, (lo, hi) in zip(cs, shards)]
        for c in cs:
            c.set_kernel(b2.KERNEL_TCGEN05)
        for rep in range(3):                                     # several exchanges: epoch / parity handling
            out, err = _run_both([lambda c=c, a=a: c.fit(a[0], a[1]) for c, a in zip(cs, dev)])
            assert err == [None, None], err
            S = [c.gram_export() for c in cs]
            assert np.array_equal(S[0], S[1])                    # bit-identical on both ranks
            assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
            assert S[0][d, d] == n and _rel(S[0], full) < 2e-6
            assert np.max(np.abs(out[0][0] - ref["coef"])) < COEF_TOL
        assert cs[0].stats()["fused_fits"] == 3 and cs[0].stats()["peer_exchanges"] == 3
        # the stand-alone exchange (b2_gram_allreduce) with the exact kernel: statistic to 1e-12
        for c in cs:
            c.set_kernel(b2.KERNEL_SIMT)

        def seq(c, a):
            c.gram_reset(d); c.gram_accumulate(a[0], a[1]); c.gram_allreduce()
            return c.solve()
        out, err = _run_both([lambda c=c, a=a: seq(c, a) for c, a in zip(cs, dev)])
        assert err == [None, None], err
        S = [c.gram_export() for c in cs]
        assert np.array_equal(S[0], S[1]) and _rel(S[0], full) < 1e-12
        assert np.max(np.abs(out[0][0] - ref["coef"])) < 1e-8
    finally:
        for c in cs:
            c.close()


def test_peer_exchange_timeout_is_an_error_not_a_partial_fit():
    n, d = 20_000, 128
    X, y = orc.generate_dataset(n, d, seed=2, dtype=np.float32)
    cs = _two_contexts()
    try:
        b2.Context.comm_p2p_attach_local(cs)
        Xd, yd = cs[0].to_device(X), cs[0].to_device(y)
        cs[0].comm_set_timeout_ms(150)
        # rank 1 never takes part: both flavours of the exchange must fail on rank 0
        cs[0].set_kernel(b2.KERNEL_TCGEN05)
        with pytest.raises(RuntimeError, match="timed out"):
            cs[0].fit(Xd, yd)                                   # wait inside the fused solve kernel
        cs[0].gram_reset(d); cs[0].gram_accumulate(Xd, yd); cs[0].gram_allreduce()
        with pytest.raises(RuntimeError, match="timed out"):
            cs[0].solve()                                       # status word of the gather kernel
        # after a failed exchange the group re-attaches (exchange numbers restart on every rank) and works again
        for c in cs:
            c.comm_p2p_detach()
        b2.Context.comm_p2p_attach_local(cs)
        cs[0].comm_set_timeout_ms(10_000)
        X1d, y1d = cs[1].to_device(X), cs[1].to_device(y)
        cs[1].set_kernel(b2.KERNEL_TCGEN05)
        out, err = _run_both([lambda: cs[0].fit(Xd, yd), lambda: cs[1].fit(X1d, y1d)])
        assert err == [None, None], err
        S = cs[0].gram_export()
        assert S[d, d] == 2 * n
        ref = orc.fit_from_stats(2.0 * orc.gram_stats(X, y))
        assert np.max(np.abs(out[0][0] - ref["coef"])) < COEF_TOL
    finally:
        for c in cs:
            c.close()


def test_nccl_entry_points_on_a_one_rank_communicator(ctx):
    """b2_score_allreduce / b2_comm_barrier / the NCCL flavour of b2_gram_allreduce run their real ncclAllReduce calls
    on a communicator of one rank (a single-GPU box can exercise them); the N > 1 values are checked by bench.py."""
    c = b2.Context(0)
    try:
        c.comm_init(1, 0, b2.Context.comm_unique_id())
        assert c.comm_info()["exchange"] == "nccl"
        n, d = 30_000, 16
        X, y = orc.generate_dataset(n, d, seed=8, dtype=np.float32)
        Xd, yd = c.to_device(X), c.to_device(y)
        c.gram_reset(d); c.gram_accumulate(Xd, yd); c.gram_allreduce(); coef, b0 = c.solve()
        ref = _oracle_fit(X, y)
        assert np.max(np.abs(coef - ref["coef"])) < COEF_TOL
        _, stats = c.score(Xd, coef, b0, y=yd, want_yhat=False)
        red = c.score_allreduce(stats.copy())
        assert np.allclose(red, stats, rtol=0, atol=0)
        c.comm_barrier()
    finally:
        c.close()


# ------------------------------------------------------------------------------------------------
# spectrum without eigenvectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(2000, 1), (500, 2), (5000, 8), (3000, 33), (20_000, 128), (4000, 127), (300, 100)])
def test_eigvals_match_the_svd_of_the_centred_rows(ctx, n, d):
    rng = np.random.RandomState(n + d)
    X = (rng.rand(n, d) * 100).astype(np.float32)
    X[:, 0] *= 1e-2                                               # spread the spectrum
    if d > 4:
        X[:, 3] += 0.999 * X[:, 2]                                # a nearly dependent pair
    y = (X.sum(axis=1) + rng.randn(n)).astype(np.float32)
    ctx.set_kernel(b2.KERNEL_SIMT)
    try:
        Xd, yd = ctx.to_device(X), ctx.to_device(y)
        ctx.gram_reset(d); ctx.gram_accumulate(Xd, yd)
        sing, rank, rows = ctx.solve_eigvals(cond=1e-6)
        Xd.free(); yd.free()
    finally:
        ctx.set_kernel(b2.KERNEL_AUTO)
    Xc = X.astype(np.float64) - X.astype(np.float64).mean(axis=0)
    sv = np.linalg.svd(Xc, compute_uv=False)
    assert rows == n
    # eigenvalues of the Gram are exact to eps * lambda_max, so singular values to eps * s_max^2 / (2 s)
    lam, lam_ref = sing ** 2, sv[:d] ** 2
    assert np.max(np.abs(lam - lam_ref)) < 1e-9 * lam_ref[0]
    assert rank == int(np.sum(sv > 1e-6 * sv[0]))
    assert np.all(np.diff(sing) <= 0)


def test_estimator_attributes_match_sklearn_and_rank_deficiency_falls_back_to_min_norm(ctx):
    from sklearn.linear_model import LinearRegression
    rng = np.random.RandomState(4)
    X = (rng.rand(5000, 12) * 50).astype(np.float32)
    y = (X @ np.arange(1, 13) + 3 + rng.randn(5000)).astype(np.float32)
    est = b2.B200LinearRegression(ctx=ctx).fit(X, y)
    ref = LinearRegression().fit(X.astype(np.float64), y.astype(np.float64))
    assert est.rank_ == ref.rank_ == 12
    assert np.allclose(est.singular_, ref.singular_, rtol=1e-7)
    assert np.max(np.abs(est.coef_ - ref.coef_)) < 1e-6
    X2 = X.copy(); X2[:, 5] = X2[:, 4]                            # exactly dependent columns: Cholesky must refuse
    est2 = b2.B200LinearRegression(ctx=ctx).fit(X2, y)
    ref2 = LinearRegression().fit(X2.astype(np.float64), y.astype(np.float64))
    assert est2.rank_ == ref2.rank_ == 11
    assert np.max(np.abs(est2.coef_ - ref2.coef_)) < 1e-5         # the minimum-norm solution gelsd returns
    sk = est2.to_sklearn()
    assert sk.rank_ == 11 and sk.singular_.shape == (12,)


# ------------------------------------------------------------------------------------------------
# model_metrics on float64 vectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 1317, 300_001])
def test_metrics_on_float64_vectors_match_sklearn_to_rounding(ctx, n):
    from sklearn.metrics import max_error, mean_absolute_percentage_error, r2_score
    rng = np.random.RandomState(n)
    y = rng.normal(50, 30, n)
    p = y + rng.normal(0, 10, n)
    if n > 5:
        y[3] = 0.0                                                # the eps clamp of sklearn's MAPE
    stats = ctx.metrics(y, p)
    so = orc.score_stats(y, p)
    fin = np.isfinite(so)
    assert np.allclose(stats[fin], so[fin], rtol=1e-13, atol=0) and np.array_equal(np.isfinite(stats), fin)
    from bodywork_mlops_demo_b200 import stage_1_train_model as s1
    mape, r2, mx = s1.metrics_from_stats(stats)
    assert np.isclose(mape, mean_absolute_percentage_error(y, p), rtol=1e-13)
    assert np.isclose(mx, max_error(y, p), rtol=1e-15)
    if n > 1:
        assert np.isclose(r2, r2_score(y, p), rtol=1e-11)
    # device vectors, float32 flavour
    yd, pd_ = ctx.to_device(y.astype(np.float32)), ctx.to_device(p.astype(np.float32))
    s32 = ctx.metrics(yd, pd_)
    so32 = orc.score_stats(y.astype(np.float32).astype(np.float64), p.astype(np.float32).astype(np.float64))
    fin = np.isfinite(so32)
    assert np.allclose(s32[fin], so32[fin], rtol=1e-13)
    yd.free(); pd_.free()


# ------------------------------------------------------------------------------------------------
# the data-generating process on the device (stage_3_synthetic_data_generation.py:28-43)
# ------------------------------------------------------------------------------------------------
def test_synth_tranche_follows_the_reference_dgp(ctx):
    from scipy import stats as st
    n, day = 400_000, 46
    alpha = 1.0 + 0.5 * np.sin(2 * np.pi * 6 * (day - 1) / 364)   # stage_3...:31-33,38
    Xd, yd, kept = ctx.synth_tranche(n, day, seed=77)
    X, y = Xd.to_host()[:kept, 0].astype(np.float64), yd.to_host()[:kept].astype(np.float64)
    assert np.all(y >= 0.0) and 0 < kept < n                      # rows with y < 0 are dropped (:43)
    # P(y >= 0) = E_X Phi((alpha + 0.5 X) / 10), X ~ U(0, 100)
    xs = (np.arange(200_000) + 0.5) / 2000.0
    p_keep = float(np.mean(st.norm.cdf((alpha + 0.5 * xs) / 10.0)))
    assert abs(kept / n - p_keep) < 4 * np.sqrt(p_keep * (1 - p_keep) / n)
    # where the filter is inactive (X > 80: 4 sigma) the rows are the unfiltered DGP: X uniform, eps standard normal
    hi = X > 80.0
    assert st.kstest((X[hi] - 80.0) / 20.0, "uniform").pvalue > 1e-3
    eps = (y[hi] - alpha - 0.5 * X[hi]) / 10.0
    assert st.kstest(eps, "norm").pvalue > 1e-3
    assert abs(eps.mean()) < 5 / np.sqrt(hi.sum())                # alpha(day) is the intercept of that day
    # the compaction keeps order and is reproducible; a different day shifts the intercept
    X2d, y2d, kept2 = ctx.synth_tranche(n, day, seed=77)
    assert kept2 == kept and np.array_equal(X2d.to_host()[:kept, 0], Xd.to_host()[:kept, 0])
    X3d, y3d, kept3 = ctx.synth_tranche(n, 16, seed=77)           # alpha(16) = 1.5 (peak of the sinusoid)
    X3, y3 = X3d.to_host()[:kept3, 0].astype(np.float64), y3d.to_host()[:kept3].astype(np.float64)
    hi3 = X3 > 80.0
    assert abs(np.mean(y3[hi3] - 0.5 * X3[hi3]) - (1.0 + 0.5 * np.sin(2 * np.pi * 6 * 15 / 364))) < 5 * 10 / np.sqrt(hi3.sum())
    # nothing is filtered when the noise cannot reach zero, and the reference's 1 440-row day keeps ~92 %
    X4d, y4d, kept4 = ctx.synth_tranche(50_000, 1, seed=3, sigma=1e-3)
    assert kept4 == 50_000 and st.kstest(X4d.to_host()[:, 0].astype(np.float64) / 100.0, "uniform").pvalue > 1e-3
    X5d, y5d, kept5 = ctx.synth_tranche(1440, 1, seed=11)
    assert 1250 <= kept5 <= 1400                                  # notebooks/4-...ipynb: 1 317 - 1 346 of 1 440
    for a in (Xd, yd, X2d, y2d, X3d, y3d, X4d, y4d, X5d, y5d):
        a.free()


def test_synth_rows_have_the_reference_marginals(ctx):
    """b2_synth (D columns, no filter): X_ij ~ U(0, 100) i.i.d., eps ~ N(0, 1) -- goodness of fit, not just moments."""
    from scipy import stats as st
    n, d = 200_000, 8
    Xd, yd = ctx.synth(n, d, seed=99)
    X, y = Xd.to_host().astype(np.float64), yd.to_host().astype(np.float64)
    for j in (0, 3, 7):
        assert st.kstest(X[:, j] / 100.0, "uniform").pvalue > 1e-3
    eps = (y - 1.0 - 0.5 * X.sum(axis=1)) / 10.0
    assert st.kstest(eps, "norm").pvalue > 1e-3                   # fp32 rounding of y (~1e-5) is far below the KS resolution
    assert abs(np.corrcoef(X[:, 0], X[:, 1])[0, 1]) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(X[:-1, 2], X[1:, 2])[0, 1]) < 5 / np.sqrt(n)   # consecutive rows are independent draws
    Xd.free(); yd.free()


def test_replay_on_device_generated_reference_tranches(ctx):
    """BASELINE configs[4] with device-generated D = 1 tranches (alpha(day) drift + filter): the incremental refit
    equals the oracle refit on the same cumulative train rows every day."""
    from bodywork_mlops_demo_b200 import incremental
    from bodywork_mlops_demo_b200.stage_1_train_model import split_mask
    tranches = []
    for day in range(1, 11):
        Xd, yd, kept = ctx.synth_tranche(1440, day, seed=1000 + day)
        tranches.append((Xd.to_host()[:kept].copy(), yd.to_host()[:kept].copy()))
        Xd.free(); yd.free()
    res = incremental.replay(tranches, d=1, mode="incremental", ctx=ctx)
    Xs, ys = [], []
    for (X, y), r in zip(tranches, res):
        m = split_mask(len(y))
        Xs.append(X[m == 1]); ys.append(y[m == 1])
        ref = orc.fit_from_stats(orc.gram_stats(np.concatenate(Xs).astype(np.float64), np.concatenate(ys).astype(np.float64)))
        assert abs(r.coef[0] - ref["coef"][0]) < 1e-9 and abs(r.intercept - ref["intercept"]) < 1e-8
    assert 0.45 < res[-1].coef[0] < 0.55


# ------------------------------------------------------------------------------------------------
# host-side regressions called out by the round-1 review
# ------------------------------------------------------------------------------------------------
def test_back_to_back_host_streamed_accumulates_do_not_overwrite_a_block_in_use(ctx):
    rows = (1 << 18) + 12_345                                     # two staging blocks per call: the call ends on buffer 1,
    d = 32                                                        # the next one starts on buffer 0 while it may be in use
    X, y = orc.generate_dataset(rows, d, seed=13, dtype=np.float32)
    Xp, yp = ctx.pinned((rows, d), np.float32), ctx.pinned((rows,), np.float32)
    Xp.array[:] = X; yp.array[:] = y
    ctx.gram_reset(d)
    reps = 5
    for _ in range(reps):
        ctx.gram_accumulate(Xp.array, yp.array)                   # no sync in between
    S = ctx.gram_export()
    one = orc.gram_stats(X, y)
    assert S[d, d] == reps * rows and _rel(S, reps * one) < 2e-6
    Xp.free(); yp.free()


def test_pageable_host_rows_take_the_bounce_ring_and_equal_pinned_rows(ctx):
    """numpy / pandas rows are pageable: they go through the library's pinned bounce ring (host threads copy block k+1
    while block k is on the wire) -- same statistic, same predictions as page-locked rows."""
    rows, d = 2 * (1 << 18) + 4_321, 64                           # three staging blocks
    X, y = orc.generate_dataset(rows, d, seed=21, dtype=np.float32)
    mask = (np.arange(rows) % 5 != 0).astype(np.uint8)
    Xp, yp, mp = ctx.pinned((rows, d), np.float32), ctx.pinned((rows,), np.float32), ctx.pinned((rows,), np.uint8)
    Xp.array[:] = X; yp.array[:] = y; mp.array[:] = mask
    c_pin, b_pin = ctx.fit(Xp.array, yp.array, mp.array, 1); S_pin = ctx.gram_export()
    c_pag, b_pag = ctx.fit(X, y, mask, 1); S_pag = ctx.gram_export()
    assert np.array_equal(S_pin, S_pag) and np.array_equal(c_pin, c_pag) and b_pin == b_pag
    ref = _oracle_fit(X, y, mask)
    assert np.max(np.abs(c_pag - ref["coef"])) < COEF_TOL
    yh_pag, st_pag = ctx.score(X, c_pag, b_pag, y=y, row_mask=mask, mask_keep=0)
    yh_pin, st_pin = ctx.score(Xp.array, c_pag, b_pag, y=yp.array, row_mask=mp.array, mask_keep=0)
    assert np.array_equal(yh_pag, yh_pin) and np.array_equal(st_pag, st_pin)
    for a in (Xp, yp, mp):
        a.free()


def test_estimators_sharing_a_context_do_not_share_a_statistic(ctx):
    Xa, ya = orc.generate_dataset(6000, 8, seed=1, dtype=np.float32)
    Xb, yb = orc.generate_dataset(5000, 8, seed=2, dtype=np.float32)
    yb = (yb + 7.0).astype(np.float32)
    e1 = b2.B200LinearRegression(ctx=ctx).fit(Xa, ya, with_spectrum=False)
    e2 = b2.B200LinearRegression(ctx=ctx).partial_fit(Xb, yb)     # must NOT fold B into A's rows
    ref_b = _oracle_fit(Xb, yb)
    assert np.max(np.abs(e2.coef_ - ref_b["coef"])) < COEF_TOL and abs(e2.intercept_ - ref_b["intercept"]) < 1e-2
    with pytest.raises(RuntimeError, match="no longer resident"):
        e1.to_sklearn()                                           # A's deferred spectrum would come from B's rows
    e2.partial_fit(Xa, ya)                                        # e2 = B then A, from its own statistic
    e3 = b2.B200LinearRegression(ctx=ctx).fit(Xa, ya)             # someone else uses the context in between
    e2.partial_fit(Xb, yb)
    ref = _oracle_fit(np.concatenate([Xb, Xa, Xb]), np.concatenate([yb, ya, yb]))
    assert np.max(np.abs(e2.coef_ - ref["coef"])) < COEF_TOL
    assert e2.to_sklearn().rank_ == 8 and e3.rank_ == 8
