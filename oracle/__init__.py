"""CPU oracle for the pointwise-downscaling hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain-numpy restatement of the algorithms that
pangeo-data/scikit-downscale runs per grid cell (reference files cited per
function as ``skdownscale/pointwise_models/<file>:<lines>``).  It exists so the
CUDA path can be checked on machines where ``/root/reference`` is absent (the
GPU box).  It is NOT part of the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* the product (``scikit-downscale_b200``) never imports it and has no CPU
  fallback — it fails loudly when the CUDA library is missing.

Parity pin: the oracle is asserted against the LIVE reference estimators
(imported from ``/root/reference`` in the build container) through the golden
vectors committed under ``tests/golden/`` (generator:
``tests/golden/make_golden.py``) and against the reference's own known-answer
tests (``skdownscale/test/test_pointwise_models.py:81-90`` quantile mapper,
``:302-312`` padded DOY grouper, ``:323-344`` EquidistantCdfMatcher, ``:236-299`` ZScoreRegressor) — see
``tests/test_oracle_golden.py``.  One exception, stated in its header: the FIT statistics of
``oracle/zscore.py`` restate xarray code that cannot run here and are pinned only by the reference's known-answer
tests (predict is pinned to the live reference).
"""

from .groupers import (  # noqa: F401
    day_keys,
    groups_from_keys,
    month_keys,
    padded_doy_groups,
)
from .quantile import (  # noqa: F401
    cunnane_inverse,
    edcdf_predict,
    extended_cdf,
    plotting_positions,
    qm_regressor_fit,
    qm_regressor_predict,
    qmr_well_conditioned,
    linear_trend_fit,
    quantile_mapper_fit,
    quantile_mapper_fit_detrend,
    quantile_mapper_transform,
    quantile_mapper_transform_detrend,
    rank_max_ties,
    trend_aware_qm_fit_predict,
)
from .bcsd import (  # noqa: F401
    bcsd_precipitation_fit,
    bcsd_precipitation_predict,
    bcsd_temperature_fit,
    bcsd_temperature_predict,
)
from .gard import (  # noqa: F401
    analog_regression_predict,
    knn_bruteforce,
    pure_analog_predict,
    pure_regression_fit_predict,
)
from .wrapper import pointwise_fit_predict  # noqa: F401
from .zscore import (  # noqa: F401
    zscore_calc_stats,
    zscore_day_table,
    zscore_expand_index,
    zscore_fit,
    zscore_predict,
    zscore_window_columns,
)
