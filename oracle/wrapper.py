"""Oracle: the PointWiseDownscaler cell loop (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows skdownscale/pointwise_models/core.py:35-37 (NaN-cell mask on the first
timestep / first feature), :69-97 (_fit_wrapper: one model per unmasked cell),
:100-143 (_predict_wrapper: output NaN-filled in X.dtype, per-cell predict,
``.squeeze()`` then cast on assignment).
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from . import bcsd, gard, groupers, quantile


def _bcsd_groups(time_grouper, index_fit, index_pred):
    """Group structures of BcsdBase._pre_fit/_create_groups (bcsd.py:34-57)."""
    if time_grouper == 'daily_nasa-nex':
        fit_groups = groupers.padded_doy_groups(index_fit)
        roll_groups = groupers.groups_from_keys(groupers.month_keys(index_pred))
        qm_groups = groupers.groups_from_keys(groupers.day_keys(index_pred))
    elif time_grouper == 'month':
        fit_groups = groupers.groups_from_keys(groupers.month_keys(index_fit))
        roll_groups = groupers.groups_from_keys(groupers.month_keys(index_pred))
        qm_groups = roll_groups
    else:
        raise ValueError(time_grouper)
    return fit_groups, roll_groups, qm_groups


def pointwise_fit_predict(spec: dict, X_train, y_train, X_pred, index_fit=None, index_pred=None):
    """fit on (X_train, y_train), predict on X_pred, every cell independently.

    Shapes: single-feature models take ``[T, C]``; GARD models take ``[T, p, C]``.
    ``spec['name']`` ∈ {QuantileMapper, BcsdTemperature, BcsdPrecipitation,
    PureAnalog, AnalogRegression}.  QuantileMapper is fitted on ``y_train``
    (``X_train`` ignored) and transforms ``X_pred``.
    Returns ``[T_pred, C]`` (or ``[T_pred, 3, C]``) in ``X_pred.dtype``.
    """
    name = spec['name']
    X_pred = np.asarray(X_pred)
    C = X_pred.shape[-1]
    Tp = X_pred.shape[0]
    multi = name in ('PureAnalog', 'AnalogRegression', 'PureRegression')
    out = np.full((Tp, 3, C) if multi else (Tp, C), np.nan, dtype=X_pred.dtype)   # core.py:129-135
    first = X_train if name != 'QuantileMapper' else y_train
    first = np.asarray(first)
    valid = ~np.isnan(first[0, 0] if first.ndim == 3 else first[0])    # core.py:35-37
    if name in ('BcsdTemperature', 'BcsdPrecipitation'):
        index_fit = pd.DatetimeIndex(index_fit)
        index_pred = pd.DatetimeIndex(index_pred if index_pred is not None else index_fit)
        fit_groups, roll_groups, qm_groups = _bcsd_groups(spec.get('time_grouper', 'month'),
                                                          index_fit, index_pred)
        anoms = spec.get('return_anoms', True)
        how = 'frame' if spec.get('time_grouper', 'month') == 'daily_nasa-nex' else 'groupby'
    for c in range(C):
        if not valid[c]:
            continue
        if name == 'QuantileMapper':
            if spec.get('detrend'):
                st = quantile.quantile_mapper_fit_detrend(np.asarray(y_train)[:, c])
                res = quantile.quantile_mapper_transform_detrend(X_pred[:, c], st, **(spec.get('qt_kwargs') or {}))
            else:
                st = quantile.quantile_mapper_fit(np.asarray(y_train)[:, c])
                res = quantile.quantile_mapper_transform(X_pred[:, c], st, **(spec.get('qt_kwargs') or {}))
        elif name == 'BcsdTemperature':
            st = bcsd.bcsd_temperature_fit(np.asarray(X_train)[:, c], np.asarray(y_train)[:, c], fit_groups, how,
                                           detrend=bool(spec.get('detrend')))
            res = bcsd.bcsd_temperature_predict(st, X_pred[:, c], roll_groups, qm_groups, anoms,
                                                qt=spec.get('qt_kwargs'))
        elif name == 'BcsdPrecipitation':
            st = bcsd.bcsd_precipitation_fit(np.asarray(y_train)[:, c], fit_groups, anoms, how,
                                             detrend=bool(spec.get('detrend')))
            res = bcsd.bcsd_precipitation_predict(st, X_pred[:, c], qm_groups, anoms, qt=spec.get('qt_kwargs'))
        elif name == 'PureAnalog':
            res = gard.pure_analog_predict(np.asarray(X_train)[:, :, c], np.asarray(y_train)[:, c],
                                           X_pred[:, :, c], spec.get('n_analogs', 200),
                                           spec.get('kind', 'best_analog'), spec.get('thresh'),
                                           spec.get('rand_inds'))
        elif name == 'AnalogRegression':
            res = gard.analog_regression_predict(np.asarray(X_train)[:, :, c], np.asarray(y_train)[:, c],
                                                 X_pred[:, :, c], spec.get('n_analogs', 200),
                                                 spec.get('thresh'), logistic_C=spec.get('logistic_C', 1.0))
        elif name == 'PureRegression':
            res = gard.pure_regression_fit_predict(np.asarray(X_train)[:, :, c], np.asarray(y_train)[:, c],
                                                   X_pred[:, :, c], spec.get('thresh'), spec.get('logistic_C', 1.0))
        else:
            raise ValueError(name)
        if multi:
            out[:, :, c] = res
        else:
            out[:, c] = res                                              # core.py:141 (cast to X.dtype)
    return out
