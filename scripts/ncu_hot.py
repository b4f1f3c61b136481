"""Hottest SASS lines of the first kernel in an .ncu-rep (needs --import-source on): python scripts/ncu_hot.py rep [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if "Instructions Executed" in r)
hdr = rows[h]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[h + 1:]):
    if len(r) <= iex or r[0] == "Kernel Name" or r[0] == "Address":
        if r and r[0] == "Kernel Name" and data:
            break
        continue
    try:
        data.append((k, int(r[iex]), int(r[ist]), r[isrc].strip()[:80]))
    except ValueError:
        pass
ti, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"instructions executed {ti}, stall samples {ts}, SASS lines {len(data)}")
print("-- by stall samples")
for d in sorted(data, key=lambda x: -x[2])[:n]:
    print(f"  line {d[0]:5d}  exec {d[1]:9d} ({100 * d[1] / ti:4.1f}%)  samples {d[2]:7d} ({100 * d[2] / max(ts, 1):4.1f}%)  {d[3]}")
print("-- by executed count")
for d in sorted(data, key=lambda x: -x[1])[:n // 2]:
    print(f"  line {d[0]:5d}  exec {d[1]:9d} ({100 * d[1] / ti:4.1f}%)  samples {d[2]:7d}  {d[3]}")
