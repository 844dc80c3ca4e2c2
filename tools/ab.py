"""A/B timing of library variants on the bench workload (48,000 frames x 2048 c128, device-resident):
usage: python tools/ab.py [--steps 100] [--rounds 3] default=amcpy_b200/_lib/libamcpy_b200.so name=path ...
Runs every variant in its own process (AMCPY_B200_LIB), interleaved `rounds` times, prints ms per launch."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
args = sys.argv[1:]
steps, rounds = 100, 3
specs = []
i = 0
while i < len(args):
    if args[i] == "--steps":
        steps = int(args[i + 1]); i += 2
    elif args[i] == "--rounds":
        rounds = int(args[i + 1]); i += 2
    else:
        specs.append(args[i]); i += 1
res = {}
for r in range(rounds):
    for spec in specs:
        name, _, path = spec.partition("=")
        env = dict(os.environ, AMCPY_B200_LIB=str((ROOT / path).resolve()))
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-e2e", "--steps", str(steps), "--warmup", "5"],
                             env=env, capture_output=True, text=True, cwd=str(ROOT))
        try:
            ms = json.loads(out.stdout.strip().splitlines()[-1])["kernel_ms_per_launch"]
        except Exception:  # noqa: BLE001
            ms = None
            print(name, "FAILED", out.stderr[-400:])
        res.setdefault(name, []).append(ms)
for name, v in res.items():
    ok = [x for x in v if x is not None]
    print(json.dumps({"variant": name, "ms": v, "best": min(ok) if ok else None}))
