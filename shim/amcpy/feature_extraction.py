"""amcpy.feature_extraction -> amcpy_b200.feature_extraction (reference: src/amcpy/feature_extraction.py:42-99)."""
from amcpy_b200.feature_extraction import _modulation_process, extract_all, run_extraction  # noqa: F401
