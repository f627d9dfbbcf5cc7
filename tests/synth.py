"""Seeded synthetic inputs shared by the golden generator, the tests,
``__graft_entry__.smoke()`` and ``bench.py``.  Host-side numpy only."""

from __future__ import annotations

import numpy as np
import pandas as pd


def daily_index(n: int, start: str = '1981-01-01') -> pd.DatetimeIndex:
    return pd.date_range(start, periods=n, freq='D')


def temperature(T: int, C: int, seed: int = 0, dtype=np.float32, t0: int = 0):
    """X_train, y_train, X_pred  [T, C] — seasonal cycle + independent noise per cell."""
    rng = np.random.default_rng(seed)
    t = (np.arange(T) + t0)[:, None]
    s = np.sin(2 * np.pi * t / 365.25)
    Xtr = (15 + 10 * s + 3 * rng.standard_normal((T, C))).astype(dtype)
    ytr = (14 + 12 * s + 2 * rng.standard_normal((T, C))).astype(dtype)
    Xp = (16.5 + 10 * s + 3 * rng.standard_normal((T, C))).astype(dtype)
    return Xtr, ytr, Xp


def precipitation(T: int, C: int, seed: int = 1, dtype=np.float32):
    """Zero-inflated gamma; p_dry = .6 / .5 / .55 for X_train / y_train / X_pred."""
    rng = np.random.default_rng(seed)

    def one(p_dry):
        wet = rng.random((T, C)) >= p_dry
        return np.where(wet, rng.gamma(0.8, 6.0, (T, C)), 0.0).astype(dtype)

    return one(0.6), one(0.5), one(0.55)


def analog(T: int, Tq: int, C: int, p: int = 3, seed: int = 2, dtype=np.float32):
    """X_train [T,p,C], y_train [T,C], X_pred [Tq,p,C] — N(0,1) predictors, linear target + noise."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, p, C))
    w = np.array([1.0, 0.5, -0.3, 0.2, -0.1, 0.4, -0.25, 0.15][:p])[None, :, None]
    y = (X * w).sum(axis=1) + 0.3 * rng.standard_normal((T, C))
    Xq = rng.standard_normal((Tq, p, C))
    return X.astype(dtype), y.astype(dtype), Xq.astype(dtype)
