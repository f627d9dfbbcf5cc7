"""Time groupers of the BCSD models — same names and meaning as
skdownscale/pointwise_models/groupers.py:11-89 — plus the host-side builders that turn
them into the integer group tables the CUDA kernels consume (done ONCE per call here;
the reference re-derives them per grid cell with ``Index.map`` / 366 boolean filters).
"""

from __future__ import annotations

import warnings

import numpy as np
import pandas as pd


class SkdownscaleGroupGeneratorBase:
    pass


def MONTH_GROUPER(x):
    return x.month


def DAY_GROUPER(x):
    return x.day


def grouper_keys(grouper, index) -> np.ndarray:
    """Group key of every timestamp.  MONTH/DAY groupers take the vectorised route; any other
    callable is mapped over the index like ``df.groupby(callable)`` does (bcsd.py:48-49)."""
    index = pd.DatetimeIndex(index) if not isinstance(index, pd.Index) else index
    if grouper is MONTH_GROUPER:
        return np.asarray(index.month)
    if grouper is DAY_GROUPER:
        return np.asarray(index.day)
    if callable(grouper):
        return np.asarray(index.map(grouper))
    raise NotImplementedError(
        f'time grouper {grouper!r} is not supported on the B200 path (callables and "daily_nasa-nex" are)')


def groups_from_keys(keys: np.ndarray):
    """Sorted unique keys with their time-ordered row numbers (pandas groupby order)."""
    keys = np.asarray(keys)
    order = np.argsort(keys, kind='stable')
    sk = keys[order]
    cut = np.flatnonzero(np.concatenate(([True], sk[1:] != sk[:-1])))
    ends = np.concatenate((cut[1:], [len(sk)]))
    out = []
    for a, b in zip(cut, ends):
        k = sk[a]
        out.append((k.item() if hasattr(k, 'item') else k, order[a:b]))
    return out


def padded_doy_member_days(n: int, ring: int, offset: int = 15) -> np.ndarray:
    """Days-of-year that belong to padded group ``n`` on a ``ring``-day calendar (365 or 366).

    The reference slices a ``np.pad(..., mode='wrap')`` copy of 1..ring (groupers.py:36-64); the
    same sets in modular arithmetic: wrapped position p holds day ((p - offset) mod ring) + 1, the
    group takes positions [n-1, n-1+offset) and [n+offset, n+2*offset), clipped to the wrapped
    length ring + 2*offset, plus day n itself.  The clipping is what makes group 366 of a
    365-day calendar irregular (days 351..365 and 2..15).
    """
    length = ring + 2 * offset
    first = np.arange(n - 1, min(n - 1 + offset, length))
    second = np.arange(n + offset, min(n + 2 * offset, length))
    pos = np.concatenate((first, second))
    days = (pos - offset) % ring + 1
    return np.concatenate((days, [n]))


def padded_doy_groups(index, offset: int = 15):
    """All 366 padded day-of-year groups as (key, rows); rows of leap years first, then rows
    of non-leap years, each in time order (groupers.py:73-78)."""
    index = pd.DatetimeIndex(index)
    doy = np.asarray(index.dayofyear)
    leap = np.asarray(index.is_leap_year)
    rows_leap = np.flatnonzero(leap)
    rows_noleap = np.flatnonzero(~leap)
    doy_leap = doy[rows_leap]
    doy_noleap = doy[rows_noleap]
    out = []
    for n in range(1, 367):
        in_leap = np.isin(doy_leap, padded_doy_member_days(n, 366, offset))
        in_noleap = np.isin(doy_noleap, padded_doy_member_days(n, 365, offset))
        out.append((n, np.concatenate((rows_leap[in_leap], rows_noleap[in_noleap]))))
    return out


def rolling_neighbours(groups, n_rows: int, window: int = 9) -> np.ndarray:
    """For every row, the rows of the centred ``window``-sample window inside its group
    (``rolling(window, center=True, min_periods=1)``, bcsd.py:247-250); -1 = absent."""
    half = window // 2
    nbr = np.full((n_rows, window), -1, dtype=np.int32)
    for _, rows in groups:
        n = len(rows)
        pos = np.arange(n)
        for d in range(-half, half + 1):
            src = pos + d
            ok = (src >= 0) & (src < n)
            nbr[rows[ok], d + half] = rows[src[ok]]
    return nbr


class PaddedDOYGrouper(SkdownscaleGroupGeneratorBase):
    """Iterator over 366 day-of-year groups padded by +/- ``offset`` days — drop-in for
    skdownscale.pointwise_models.PaddedDOYGrouper (groupers.py:19-89)."""

    def __init__(self, df, offset=15):
        self.df = df
        self.offset = offset
        self.max = 366
        self.n = 1
        idx = pd.DatetimeIndex(df.index)
        self.leap = 'leap' if ((idx.month == 2) & (idx.day == 29)).any() else 'noleap'
        self._groups = padded_doy_groups(idx, offset)
        self.days_of_leap_year = np.arange(1, self.max + 1)

    def __iter__(self):
        self.n = 1
        return self

    def __next__(self):
        if self.n > self.max:
            raise StopIteration
        if self.leap == 'noleap' and len(set(padded_doy_member_days(self.n, 366, self.offset).tolist())) != 2 * self.offset + 1:
            warnings.warn('leap days not included, day groups in leap years missing leap days')
        key, rows = self._groups[self.n - 1]
        self.n += 1
        return key, self.df.iloc[rows]

    def mean(self):
        arr_means = np.full((self.max, 1), np.inf)
        for key, group in self:
            arr_means[key - 1] = group.mean().values[0]
        return pd.DataFrame(arr_means, index=self.days_of_leap_year)
