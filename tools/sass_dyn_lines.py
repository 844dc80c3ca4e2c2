"""Dynamic warp-instruction count per CUDA source line: joins `ncu --page source --csv --print-source sass`
(executed counts per SASS address) with nvdisasm's line table of the SAME build of the library.
usage: python tools/sass_dyn_lines.py src_sass.csv <kernel-mangled-substring> [frames] [min_per_frame]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "amcpy_b200" / "_lib" / "libamcpy_b200.so"
csv_path, sub = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 else 48000.0
minpf = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    elf = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
    idx = None
    for line in elf.splitlines():
        m = re.match(r"\s*(0x[0-9a-f]+)\s+0\s+0\s+0x3\s+0\s+0x[0-9a-f]+\s+\.text\.(\S+)", line)
        if m and sub in m.group(2):
            idx = m.group(1)
            break
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-fun", idx, str(cubin)], capture_output=True, text=True).stdout
off2line = {}
cur = None
started = False
for line in txt.splitlines():
    if ".section" in line and ".text." in line:
        started = sub in line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(\S+)", line)
    if started and m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
i = 0
while rows[i][0] != "Address":
    i += 1
hdr = rows[i]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[i + 1:] if len(r) > iex and r[0].startswith("0x")]
base = int(body[0][0], 16)
dyn = collections.Counter()
smp = collections.Counter()
for r in body:
    off = int(r[0], 16) - base
    ln = off2line.get(off)
    dyn[ln] += int(r[iex])
    smp[ln] += int(r[ismp])
tot = sum(dyn.values())
tots = sum(smp.values())
print(f"total {tot / frames:.1f} warp-instr/frame, {tots} samples")
for k, v in sorted(dyn.items(), key=lambda kv: (kv[0] or ("", 0))):
    if v / frames >= minpf:
        print(f"{k[0] if k else '?':22s} {k[1] if k else 0:5d} {v / frames:8.1f}/frame  samples {100 * smp[k] / tots:5.2f}%")
