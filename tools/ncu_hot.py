"""Top sampled SASS instructions of a captured kernel (ncu --import-source on). Usage: ncu_hot.py REP [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
isrc, ins, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ins]) for r in data)
print(rows[0][1][:120]); print("total samples", tot)
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][ins]))[:n]:
    print(f"{idx:5d} {int(r[ins]):6d} {100*int(r[ins])/tot:5.1f}% exec={r[iex]:>8s}  {r[isrc].strip()[:100]}")
