"""Import shim: the package lives in ``scikit-downscale_b200/`` (a directory name Python
cannot import directly); this module loads it under the importable name
``skdownscale_b200`` so that ``import skdownscale_b200.pointwise_models`` works from the
repository root."""

import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'scikit-downscale_b200')
_spec = importlib.util.spec_from_file_location(
    'skdownscale_b200', os.path.join(_pkg_dir, '__init__.py'), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['skdownscale_b200'] = _mod
_spec.loader.exec_module(_mod)
