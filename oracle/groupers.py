"""Oracle: time groupers (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows skdownscale/pointwise_models/groupers.py:11-89 and the way
skdownscale/pointwise_models/bcsd.py:46-57 turns them into groups.
"""

from __future__ import annotations

import numpy as np
import pandas as pd


def month_keys(index: pd.DatetimeIndex) -> np.ndarray:
    """MONTH_GROUPER applied to every timestamp (groupers.py:11-12)."""
    return np.asarray(index.month, dtype=np.int64)


def day_keys(index: pd.DatetimeIndex) -> np.ndarray:
    """DAY_GROUPER = day of *month* (groupers.py:15-16)."""
    return np.asarray(index.day, dtype=np.int64)


def groups_from_keys(keys: np.ndarray) -> list[tuple[int, np.ndarray]]:
    """``df.groupby(callable)`` (bcsd.py:48-49): sorted unique keys, each with the
    time-ordered row numbers that carry it."""
    keys = np.asarray(keys)
    out = []
    for k in np.unique(keys):
        out.append((k.item() if hasattr(k, 'item') else k, np.flatnonzero(keys == k)))
    return out


def padded_doy_groups(index: pd.DatetimeIndex, offset: int = 15) -> list[tuple[int, np.ndarray]]:
    """PaddedDOYGrouper(df, offset) iteration (groupers.py:19-82).

    Returns 366 (key, rows) pairs, key = 1..366.  Rows of leap years come first,
    then rows of non-leap years, exactly like the ``pd.concat`` at
    groupers.py:73-78.  The slice arithmetic (groupers.py:54-64) is restated
    literally so the irregular DOY-366 group (non-leap rows: DOYs 351..365 and
    2..15) comes out the same.
    """
    index = pd.DatetimeIndex(index)
    n_max = 366
    doy = np.asarray(index.dayofyear)
    is_leap = np.asarray(index.is_leap_year)
    rows_leap = np.flatnonzero(is_leap)
    rows_noleap = np.flatnonzero(~is_leap)
    days_noleap = np.arange(1, n_max)          # groupers.py:34
    days_leap = np.arange(1, n_max + 1)        # groupers.py:35
    wrap_noleap = np.pad(days_noleap, offset, mode='wrap')   # groupers.py:36-38
    wrap_leap = np.pad(days_leap, offset, mode='wrap')       # groupers.py:39
    total_days = 2 * offset + 1
    out = []
    for n in range(1, n_max + 1):
        i = n - 1
        first_leap = wrap_leap[i:i + offset]
        first_noleap = wrap_noleap[i:i + offset]
        sec_leap = wrap_leap[n + offset:i + total_days]
        sec_noleap = wrap_noleap[n + offset:i + total_days]
        all_leap = np.concatenate((first_leap, np.array([n]), sec_leap))
        all_noleap = np.concatenate((first_noleap, np.array([n]), sec_noleap))
        if len(set(all_noleap.tolist())) != total_days and n != 366:   # groupers.py:69-70
            raise ValueError('no leap day groups do not contain the correct set of days')
        rows = np.concatenate((rows_leap[np.isin(doy[rows_leap], all_leap)],
                               rows_noleap[np.isin(doy[rows_noleap], all_noleap)]))
        out.append((n, rows))
    return out
