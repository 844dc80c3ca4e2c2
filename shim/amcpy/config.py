"""amcpy.config -> amcpy_b200.config (reference: src/amcpy/config.py:15-186)."""
from amcpy_b200.config import Config, FeatureConfig, Paths, SignalConfig, TrainingConfig  # noqa: F401
