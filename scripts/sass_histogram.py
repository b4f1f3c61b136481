"""Tensor-core / TMA opcode histogram of every kernel in the shipped library (cuobjdump -sass), so the claims "this
contraction runs on mma.sync / tcgen05" and "tiles move by TMA bulk copies" are checkable from the binary:
  python scripts/sass_histogram.py [libtensorgame_b200.so] > profiles/r02_sass_histogram.txt
HMMA / IMMA = legacy mma.sync (f16 / int8); UTCHMMA / UTCIMMA etc. = tcgen05.mma; LDTM / STTM = tcgen05.ld / st;
UBLKCP = cp.async.bulk (TMA 1-D); LDSM = ldmatrix; SYNCS = mbarrier."""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parents[1] / "mat_mul_b200" / "libtensorgame_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
WATCH = ("HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCOMMA", "UTCMXQMMA", "LDTM", "STTM", "UBLKCP", "UBLKPF", "LDSM", "SYNCS", "IDP", "REDUX", "ATOMS")
per = defaultdict(Counter)
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["total"] += 1
        if op in WATCH:
            per[name][op] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib.name}: SASS instruction counts per kernel (static); columns: total, then the tensor-core / TMA / mbarrier opcodes present")
rows = []
for mangled, pretty in zip(per, demangled):
    c = per[mangled]
    short = re.sub(r"\(.*", "", pretty).replace("tg::", "").replace("(anonymous namespace)::", "")
    rows.append((short, c))
for short, c in sorted(rows):
    ops = " ".join(f"{k}={c[k]}" for k in WATCH if c[k])
    print(f"{short:70s} total={c['total']:5d}  {ops}")
