"""`amcpy` import name for the B200 feature-extraction path.

A user of ronnymilleo/amcpy keeps `from amcpy.features import calculate_features`,
`from amcpy.feature_extraction import run_extraction`, `from amcpy.config import Config` and the `amcpy` console
script (/root/reference/pyproject.toml:56-57); every name resolves to `amcpy_b200` (CUDA library, no CPU path).
Only the modules on the extraction path and its two consumers exist here; plotting / quantisation are out of scope.
"""

__version__ = "2.0.0+b200"
