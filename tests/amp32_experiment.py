"""EXPERIMENT report (not a test, not product): the N = 2048 kernel with float32 amplitude statistics
(csrc/experiments/amc_fused16a.cuh, built as a variant library and selected with AMCPY_B200_LIB) against the oracle -
worst relative error per feature over BPSK ... 64QAM at a ladder of SNRs.  Lives beside the tests because it imports the oracle.
usage: AMCPY_B200_LIB=amcpy_b200/_lib/exp/libamcpy_b200_amp32.so python tests/amp32_experiment.py"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from amcpy_b200 import ops, synth  # noqa: E402
from oracle import amc_oracle as orc  # noqa: E402

for snr in (-10.0, 0.0, 10.0, 20.0, 30.0, 40.0, 50.0):
    x = np.concatenate([synth.cell(m, snr, 3, range(4), 2048, seed=7) for m in range(5)])
    want = orc.features_batch(x)
    got = ops.extract_features(torch.from_numpy(x).cuda()).cpu().numpy()
    rel = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300), axis=0)
    print(json.dumps({"snr_db": snr, "frames": len(x), "worst_rel_err": {str(f): float(f"{rel[f - 1]:.2e}") for f in range(1, 19)}}))
