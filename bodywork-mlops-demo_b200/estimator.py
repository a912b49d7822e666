"""Estimator-protocol mirror of what stage_1 uses from scikit-learn.

``B200LinearRegression`` keeps the constructor / ``fit`` / ``predict`` / attribute contract of
``sklearn.linear_model.LinearRegression`` as the reference uses it
(stage_1_train_model.py:105-107: ``LinearRegression(fit_intercept=True)``, ``.fit(X_train, y_train)``,
``.predict(X_test)``), with the arithmetic on the B200 through libb2gram.so.  ``alpha`` adds the
ridge term (alpha = 0 == the reference).

``to_sklearn()`` returns a genuine ``sklearn.linear_model.LinearRegression`` carrying our
coefficients: stage_2_serve_model.py:65,78,79 un-pickles the model with only sklearn / numpy /
joblib importable, calls ``.predict`` and ``str(model)`` -- a custom class could not be loaded there.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native as native

_shared_ctx: Optional[native.Context] = None


def default_context() -> native.Context:
    global _shared_ctx
    if _shared_ctx is None or _shared_ctx._h is None:
        _shared_ctx = native.Context(0)
    return _shared_ctx


def _as_f32_matrix(X) -> np.ndarray:
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    if X.ndim != 2:
        raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
    if X.shape[1] > native.MAX_D:
        raise ValueError(f"at most {native.MAX_D} features are supported, got {X.shape[1]}")
    return np.ascontiguousarray(X, dtype=np.float32)


class B200LinearRegression:
    def __init__(self, *, fit_intercept: bool = True, alpha: float = 0.0, tol: float = 1e-6,
                 ctx: Optional[native.Context] = None):
        self.fit_intercept = fit_intercept
        self.alpha = float(alpha)
        self.tol = tol
        self._ctx = ctx

    @property
    def ctx(self) -> native.Context:
        return self._ctx if self._ctx is not None else default_context()

    # -- fit -------------------------------------------------------------------------------------
    def _finish_fit(self, d: int, with_spectrum: bool) -> "B200LinearRegression":
        ctx = self.ctx
        try:
            coef, b0 = ctx.solve(alpha=self.alpha, fit_intercept=self.fit_intercept)
            spectral = None
        except np.linalg.LinAlgError:
            # rank deficient and alpha == 0: the minimum-norm solution gelsd would return (device Jacobi)
            spectral = ctx.solve_spectral(cond=self.tol, fit_intercept=self.fit_intercept)
            coef, b0 = spectral[0], spectral[1]
        if not (np.all(np.isfinite(coef)) and np.isfinite(b0)):
            # sklearn's check_array refuses such input up front; here it shows up in the statistic
            raise ValueError("Input X or y contains NaN, infinity or a value too large for dtype('float32').")
        if with_spectrum and spectral is None:
            spectral = ctx.solve_spectral(cond=self.tol, fit_intercept=self.fit_intercept)
        self.coef_ = coef
        self.intercept_ = np.float64(b0 if self.fit_intercept else 0.0)
        self.n_features_in_ = int(d)
        if spectral is not None:
            n_rows = int(round(float(ctx.gram_export()[d, d])))
            self.singular_ = spectral[2][: min(n_rows, d)]
            self.rank_ = int(spectral[3])
        return self

    def fit(self, X, y, row_mask=None, mask_keep: int = 1, with_spectrum: bool = True) -> "B200LinearRegression":
        """X: (n, D) host array (any float dtype; staged as fp32) or a ``DeviceArray`` (f32 / bf16).
        ``row_mask`` (uint8 per row) restricts the fit to rows equal to ``mask_keep``."""
        ctx = self.ctx
        if isinstance(X, native.DeviceArray):
            d = X.shape[1]
        else:
            X = _as_f32_matrix(X)
            y = np.ascontiguousarray(np.asarray(y).ravel(), dtype=np.float32)
            if y.shape[0] != X.shape[0]:
                raise ValueError(f"Found input variables with inconsistent numbers of samples: "
                                 f"[{X.shape[0]}, {y.shape[0]}]")
            d = X.shape[1]
        ctx.gram_reset(d)
        ctx.gram_accumulate(X, y, row_mask, mask_keep)
        ctx.gram_allreduce()  # no-op without a communicator
        return self._finish_fit(d, with_spectrum)

    def partial_fit(self, X, y, with_spectrum: bool = False) -> "B200LinearRegression":
        """Fold one more tranche into the running statistic and re-solve (incremental daily refit)."""
        ctx = self.ctx
        Xh = X if isinstance(X, native.DeviceArray) else _as_f32_matrix(X)
        d = Xh.shape[1]
        if ctx.d != d:
            ctx.gram_reset(d)
        if not isinstance(Xh, native.DeviceArray):
            y = np.ascontiguousarray(np.asarray(y).ravel(), dtype=np.float32)
        ctx.gram_accumulate(Xh, y)
        return self._finish_fit(d, with_spectrum)

    # -- predict -------------------------------------------------------------------------------------
    def predict(self, X):
        ctx = self.ctx
        if isinstance(X, native.DeviceArray):
            yhat, _ = ctx.score(X, self.coef_, float(self.intercept_))
            return yhat
        Xh = _as_f32_matrix(X)
        if Xh.shape[1] != self.n_features_in_:
            raise ValueError(f"X has {Xh.shape[1]} features, but B200LinearRegression is expecting "
                             f"{self.n_features_in_} features as input.")
        yhat, _ = ctx.score(Xh, self.coef_, float(self.intercept_))
        return yhat.astype(np.float64)

    # -- artefact ----------------------------------------------------------------------------------------
    def to_sklearn(self):
        """A real sklearn LinearRegression with the attributes ``fit`` would have set
        (the joblib layout stage_1_train_model.py:113-114 dumps and stage_2_serve_model.py:65 loads)."""
        from sklearn.linear_model import LinearRegression
        if not hasattr(self, "rank_"):
            spectral = self.ctx.solve_spectral(cond=self.tol, fit_intercept=self.fit_intercept)
            n_rows = int(round(float(self.ctx.gram_export()[self.n_features_in_, self.n_features_in_])))
            self.singular_ = spectral[2][: min(n_rows, self.n_features_in_)]
            self.rank_ = int(spectral[3])
        reg = LinearRegression(fit_intercept=self.fit_intercept)
        reg.coef_ = np.asarray(self.coef_, dtype=np.float64).copy()
        reg.intercept_ = np.float64(self.intercept_)
        reg.rank_ = int(self.rank_)
        reg.singular_ = np.asarray(self.singular_, dtype=np.float64).copy()
        reg.n_features_in_ = int(self.n_features_in_)
        return reg

    def __repr__(self) -> str:
        return "B200LinearRegression()" if self.alpha == 0.0 else f"B200LinearRegression(alpha={self.alpha})"
