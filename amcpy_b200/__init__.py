"""amcpy_b200 - B200-native feature extraction for Automatic Modulation Classification.

A from-scratch sm_100a implementation of amcpy's feature-extraction hot path (the 18 per-frame
statistical features), behind the reference's own operator / stage interface:

    from amcpy_b200.features import calculate_features, _FEATURE_FUNCTIONS
    from amcpy_b200.feature_extraction import run_extraction
    from amcpy_b200.ops import extract_features            # batched (n_snr, n_frames, N) entry
"""

__version__ = "0.1.0"
