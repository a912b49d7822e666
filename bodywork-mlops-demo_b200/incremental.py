"""Incremental daily refit (SURVEY.md section 5 "checkpoint/resume", section 8f rank 2; BASELINE config 5).

The reference refits from scratch on all history every day (stage_1_train_model.py:62-71) and re-draws the
80/20 split over the concatenated set (:98-103), so which rows are "train" changes as history grows.  Two modes:

* ``exact``  -- reproduce the reference: every day recompute the global ``RandomState(42)`` split mask over the
  concatenated rows and accumulate from scratch (O(history) per day, bit-for-bit the reference's train rows);
* ``incremental`` -- split each tranche on its own (same ``split_mask`` rule applied per tranche), fold only the
  new tranche's train rows into the persisted statistic S and re-solve: O(tranche) per day.  A documented
  deviation: the train/test membership differs from the reference's global re-shuffle, the estimator does not.

The state is S (``(D+2)^2`` fp64) saved beside the model as ``regressor-<date>.gram.npy``.
A replay day follows the pipeline DAG (bodywork.yaml:5): score tranche t with model(t-1) (the service test of
stage_4), then fold tranche t in and refit (stage_1 of the next day).
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from datetime import date
from typing import List, Optional

import numpy as np

from . import _native as native
from .estimator import B200LinearRegression, default_context
from .stage_1_train_model import metrics_from_stats, split_mask


def save_state(S: np.ndarray, day: date, bucket_dir: str) -> str:
    os.makedirs(os.path.join(bucket_dir, "models"), exist_ok=True)
    path = os.path.join(bucket_dir, "models", f"regressor-{day}.gram.npy")
    np.save(path, np.asarray(S, dtype=np.float64))
    return path


def load_state(day: date, bucket_dir: str) -> np.ndarray:
    return np.load(os.path.join(bucket_dir, "models", f"regressor-{day}.gram.npy"))


@dataclass
class DayResult:
    day: int
    n_rows: int
    n_train_total: int
    coef: np.ndarray
    intercept: float
    test_mape: Optional[float]      # tranche t scored with model(t-1): stage_4 semantics
    test_r2: Optional[float]
    test_max_residual: Optional[float]
    seconds: float


@dataclass
class IncrementalTrainer:
    d: int
    ctx: Optional[native.Context] = None
    mode: str = "incremental"                       # or "exact"
    history_X: List[np.ndarray] = field(default_factory=list)
    history_y: List[np.ndarray] = field(default_factory=list)
    model: Optional[B200LinearRegression] = None
    S: Optional[np.ndarray] = None
    days: int = 0

    def __post_init__(self):
        self.ctx = self.ctx or default_context()
        if self.mode not in ("incremental", "exact"):
            raise ValueError("mode must be 'incremental' or 'exact'")

    def step(self, X: np.ndarray, y: np.ndarray) -> DayResult:
        """One pipeline day: score the new tranche with yesterday's model, then fold it in and refit."""
        t0 = time.perf_counter()
        ctx = self.ctx
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(len(y), self.d)
        y = np.ascontiguousarray(y, dtype=np.float32)
        Xd, yd = ctx.to_device(X), ctx.to_device(y)
        mape = r2 = mx = None
        try:
            if self.model is not None:
                _, stats = ctx.score(Xd, self.model.coef_, float(self.model.intercept_), y=yd, want_yhat=False)
                mape, r2, mx = metrics_from_stats(stats)
            if self.mode == "incremental":
                mask = split_mask(len(y))
                md = ctx.to_device(mask)
                if self.S is None:
                    ctx.gram_reset(self.d)
                else:
                    ctx.gram_import(self.S)
                ctx.gram_accumulate(Xd, yd, md, 1)
                md.free()
            else:
                self.history_X.append(X); self.history_y.append(y)
                allX = np.concatenate(self.history_X); ally = np.concatenate(self.history_y)
                mask = split_mask(len(ally))
                ctx.gram_reset(self.d)
                ctx.gram_accumulate(allX, ally, mask, 1)      # host rows: streamed
            self.S = ctx.gram_export()
            self.model = B200LinearRegression(ctx=ctx).solve_resident(self.d, self.S)
        finally:
            Xd.free(); yd.free()
        self.days += 1
        return DayResult(self.days, len(y), int(round(self.S[self.d, self.d])), self.model.coef_.copy(),
                         float(self.model.intercept_), mape, r2, mx, time.perf_counter() - t0)


def replay(tranches, d: int, mode: str = "incremental", ctx=None) -> List[DayResult]:
    """Run the concept-drift replay over an iterable of (X, y) daily tranches."""
    tr = IncrementalTrainer(d=d, ctx=ctx, mode=mode)
    return [tr.step(X, y) for X, y in tranches]
