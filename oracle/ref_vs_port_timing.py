"""One-off check that the CPU baseline `bench.py` reports (the oracle's faithful restatement, kind "port") runs at the
speed of the UNMODIFIED reference: both are timed here, single-threaded, on the same frames.  Needs /root/reference
(only present in the build container - not on the GPU box), so the result is committed under profiles/.
usage: PYTHONDONTWRITEBYTECODE=1 python oracle/ref_vs_port_timing.py [frames] > profiles/r1n_ref_vs_port_cpu.json"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

from amcpy import features as ref  # noqa: E402  (the unmodified reference)

from amcpy_b200 import synth  # noqa: E402
from oracle import amc_oracle as orc  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 150
x = np.concatenate([synth.cell(m, 10.0, 10, range(frames // 6), 2048, seed=5) for m in range(6)])
ids = list(range(1, 19))
res = {}
with np.errstate(all="ignore"):
    for name, fn in (("reference", ref.calculate_features), ("port", orc.calculate_features_faithful),
                     ("reference", ref.calculate_features), ("port", orc.calculate_features_faithful)):
        t0 = time.perf_counter()
        out = [fn(ids, f) for f in x]
        dt = time.perf_counter() - t0
        res.setdefault(name, []).append(dt / len(x))
        res[name + "_out"] = np.asarray(out, dtype=np.float64)
same = bool(np.array_equal(res.pop("reference_out"), res.pop("port_out"), equal_nan=True))
ref_ms, port_ms = 1e3 * min(res["reference"]), 1e3 * min(res["port"])
print(json.dumps({"frames": len(x), "frame_size": 2048, "dtype": "complex128", "threads": 1,
                  "reference_ms_per_frame": round(ref_ms, 3), "port_ms_per_frame": round(port_ms, 3),
                  "port_over_reference_speed": round(ref_ms / port_ms, 3), "outputs_bitwise_equal": same,
                  "what": "amcpy.features.calculate_features(range(1,19), frame) vs oracle.amc_oracle."
                          "calculate_features_faithful on the same frames, best of 2 passes"}))
