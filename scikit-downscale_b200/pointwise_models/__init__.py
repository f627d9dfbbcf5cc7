"""Drop-in surface of ``skdownscale.pointwise_models`` for the B200 hot path
(skdownscale/pointwise_models/__init__.py:1-36).  Estimators outside the hot path
(GroupedRegressor, ...) are not provided; QuantileMappingReressor and
EquidistantCdfMatcher are the first "next" row of SURVEY.md §8(f), LinearTrendTransformer and
TrendAwareQuantileMappingRegressor the second, PureRegression and ZScoreRegressor the fourth.
"""

from .bcsd import BcsdPrecipitation, BcsdTemperature
from .core import PointWiseDownscaler
from .gard import AnalogRegression, PureAnalog, PureRegression
from .groupers import DAY_GROUPER, MONTH_GROUPER, PaddedDOYGrouper
from .quantile import (EquidistantCdfMatcher, LinearTrendTransformer, QuantileMapper, QuantileMappingReressor,
                       TrendAwareQuantileMappingRegressor)
from .zscore import ZScoreRegressor

__all__ = [
    'BcsdPrecipitation',
    'BcsdTemperature',
    'PointWiseDownscaler',
    'AnalogRegression',
    'PureAnalog',
    'PureRegression',
    'DAY_GROUPER',
    'MONTH_GROUPER',
    'PaddedDOYGrouper',
    'QuantileMapper',
    'QuantileMappingReressor',
    'EquidistantCdfMatcher',
    'LinearTrendTransformer',
    'TrendAwareQuantileMappingRegressor',
    'ZScoreRegressor',
]
