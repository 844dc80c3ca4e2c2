"""Build A/B variants of the library next to the product build (amcpy_b200/_lib/exp/libamcpy_b200_<name>.so).
A variant is selected at run time with AMCPY_B200_LIB=<path> (amcpy_b200/_native.py); none of them is the product.

usage: python tools/exp/build_variants.py name=DEF1,DEF2 [name2=DEF3 ...]
       e.g.  exp=AMC_EXPERIMENTS  twv2=AMC_TW_V2
"""
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from amcpy_b200 import _native as nat  # noqa: E402


def one(spec: str):
    name, _, defs = spec.partition("=")
    out = nat.LIB_DIR / "exp" / f"libamcpy_b200_{name}.so"
    nat.build(defines=tuple(d for d in defs.split(",") if d), out=out)
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(max_workers=4) as pool:
        for path in pool.map(one, sys.argv[1:]):
            print("built", path)
