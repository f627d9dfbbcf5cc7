"""skdownscale_b200 — B200-native pointwise statistical downscaling.

Drop-in for the hot path of ``skdownscale.pointwise_models`` (pangeo-data/scikit-downscale):
``PointWiseDownscaler.fit/predict`` over ``BcsdTemperature``, ``BcsdPrecipitation``,
``QuantileMapper``, ``PureAnalog`` and ``AnalogRegression``, executed for all grid cells at
once by hand-written sm_100a CUDA kernels behind the C ABI of ``include/sdb.h``
(``csrc/libsdb.so``).  There is no CPU fallback: using an estimator without the built
library, or without a CUDA device, raises.
"""

__version__ = '0.1.0'

from . import pointwise_models  # noqa: F401,E402
