"""Static SASS instruction count per source line of one kernel in the built library.
usage: python tools/sass_lines.py <kernel-mangled-substring> [min_count]   (needs cuobjdump/nvdisasm, -lineinfo build)"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "amcpy_b200" / "_lib" / "libamcpy_b200.so"
sub = sys.argv[1]
minc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    elf = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
    idx = None
    for line in elf.splitlines():
        m = re.match(r"\s*(0x[0-9a-f]+)\s+0\s+0\s+0x3\s+0\s+0x[0-9a-f]+\s+\.text\.(\S+)", line)
        if m and sub in m.group(2):
            idx = m.group(1)
            print("kernel", m.group(2), "index", idx)
            break
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-fun", idx, str(cubin)], capture_output=True, text=True).stdout
cur = None
cnt = collections.Counter()
n = 0
started = False
for line in txt.splitlines():
    if line.startswith("\t.section\t.text.") or line.startswith(".section	.text."):
        started = sub in line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if started and re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S+", line):
        n += 1
        cnt[cur] += 1
print("instructions:", n, "bytes:", n * 16)
for k, v in sorted(cnt.items(), key=lambda kv: (kv[0] or ("", 0))):
    if v >= minc:
        print(f"{k[0]:24s} {k[1]:5d} {v:6d}")
