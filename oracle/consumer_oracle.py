"""CPU restatement of the reference's feature-matrix consumer - TEST INFRASTRUCTURE ONLY (imported by tests/).

`preprocess_data` (/root/reference/src/amcpy/preprocessing.py:13-75) with the one-line fix that makes it runnable: line 55
assigns `mod_data[snr, :, list(f.used)]` - a (n_used, n_frames) block under numpy's advanced-indexing rule - into an
(n_frames, n_used) slot and raises ValueError; here the block is transposed.  Everything else is
the reference's own sequence of calls: zeros float32 (n, n_used) (:41), per modulation / per SNR row blocks (:44-58),
sklearn `StandardScaler().fit/transform` (:61-63), `train_test_split(test_size, random_state, stratify=y)` (:65-71).

Parity unpinned against the live reference: the unmodified function cannot run (it raises), so there are no reference
outputs to pin to; the restatement is pinned to sklearn itself, which IS the reference's arithmetic here.
"""

from __future__ import annotations

import numpy as np
from sklearn.model_selection import train_test_split
from sklearn.preprocessing import StandardScaler


def preprocess_data_fixed(cfg, matrices: dict, mode: str = "training"):
    """matrices: {MOD: float32 (n_snr, n_frames, 18)} (what loadmat returns for calculated-features/*.mat)."""
    s, t, f = cfg.signals, cfg.training, cfg.features
    snr_axis = t.training_snr if mode == "training" else t.all_snr            # preprocessing.py:38
    n_samples = s.num_frames * len(snr_axis) * len(s.modulations_with_noise)  # :40
    x = np.zeros((n_samples, f.num_used), dtype=np.float32)                   # :41
    y = np.zeros(n_samples, dtype=np.int64)                                   # :42
    for mod_idx, mod_name in enumerate(s.modulations_with_noise):             # :44-45 (feature_files order)
        mod_data = matrices[mod_name]                                         # :48-49
        base = mod_idx * s.num_frames * len(snr_axis)                         # :51
        for snr_i, snr in enumerate(snr_axis):                                # :53
            row_start = base + snr_i * s.num_frames
            row_end = row_start + s.num_frames
            x[row_start:row_end, :] = mod_data[snr, : s.num_frames, list(f.used)].T   # :55 with the transpose fixed
        y[base: base + s.num_frames * len(snr_axis)] = s.labels[mod_idx]      # :58
    scaler = StandardScaler()                                                 # :61
    scaler.fit(x)
    scaled_x = scaler.transform(x)
    x_train, x_test, y_train, y_test = train_test_split(                      # :65-71
        scaled_x, y, test_size=t.test_size, random_state=t.random_state, stratify=y)
    return x_train, x_test, y_train, y_test, scaler
