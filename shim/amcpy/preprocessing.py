"""amcpy.preprocessing -> amcpy_b200.consumer (reference: src/amcpy/preprocessing.py:13-75, transpose of :55 fixed)."""
from amcpy_b200.consumer import load_feature_set as preprocess_data  # noqa: F401
from amcpy_b200.consumer import load_feature_set_device, stack_features  # noqa: F401
