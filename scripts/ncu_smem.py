"""Shared-memory wavefronts per SASS instruction of the first kernel in an .ncu-rep: python scripts/ncu_smem.py rep units [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
units = float(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if "Instructions Executed" in r)
hdr = rows[h]
isrc, iex, iw, ii = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for k, r in enumerate(rows[h + 1:]):
    try:
        data.append((k, int(r[iex]), int(r[iw]), int(r[ii]), r[isrc].strip()[:70]))
    except (ValueError, IndexError):
        pass
tw = sum(d[2] for d in data)
print(f"shared wavefronts per unit {tw / units:.1f} (ideal {sum(d[3] for d in data) / units:.1f})")
for d in sorted(data, key=lambda x: -x[2])[:n]:
    print(f"  line {d[0]:5d} exec/unit {d[1] / units:6.2f} wavefronts/unit {d[2] / units:6.2f} ({d[2] / max(d[1], 1):5.2f} per instr, ideal {d[3] / max(d[1], 1):4.2f})  {d[4]}")
