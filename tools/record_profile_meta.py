"""Tie the committed measurement constants to the kernels they were taken from.
usage: python tools/record_profile_meta.py <bench-line.json> [<ncu summary .txt of the hot kernel>]
Writes profiles/config3_sha256.json (the N = 1 SHA-256 of BASELINE config 3's feature matrix: bench.py's
`strong_scaling.matches_n1` compares against it at every N) and, when the ncu summary is given,
profiles/traffic_per_launch.json (DRAM bytes per launch: bench.py's `roofline.traffic`).  Both carry a hash of
amcpy_b200/csrc/*; bench.py ignores them when the sources have changed since."""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

src = bench._kernel_source_hash()
line = json.loads(Path(sys.argv[1]).read_text().strip().splitlines()[-1])
ss = line.get("strong_scaling")
if ss and line.get("n_gpus") == 1:
    (ROOT / "profiles" / "config3_sha256.json").write_text(json.dumps({
        "sha256_n1": ss["sha256_of_features"], "frames_per_cell": bench.CONFIG3_FRAMES, "frames": ss["frames"],
        "kernel_source_sha256_16": src, "source": f"{sys.argv[1]} (python bench.py, 1 GPU)"}, indent=1) + "\n")
    print("config3 sha256 recorded:", ss["sha256_of_features"])
if len(sys.argv) > 2:
    txt = Path(sys.argv[2]).read_text()
    m = re.search(r"dram traffic per launch = (\d+) bytes", txt)
    (ROOT / "profiles" / "traffic_per_launch.json").write_text(json.dumps({
        "dram_bytes_per_launch": int(m.group(1)), "kernel_source_sha256_16": src,
        "source": f"{sys.argv[2]} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch of "
                  "fused16_features_kernel<2048,double2,15>)"}, indent=1) + "\n")
    print("traffic recorded:", m.group(1))
