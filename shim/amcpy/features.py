"""amcpy.features -> amcpy_b200.features (reference: src/amcpy/features.py:17-232)."""
from amcpy_b200.features import *  # noqa: F401,F403
from amcpy_b200.features import (  # noqa: F401  (underscore names are part of the reference's surface)
    _FEATURE_FUNCTIONS,
    FEATURE_NAMES,
    InstantaneousValues,
    MomentValues,
    calculate_features,
    calculate_features_batch,
)
from amcpy_b200 import features as _impl

globals().update({name: getattr(_impl, name) for name in _impl.FEATURE_NAMES})
