"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA (B200_PROFILING.md: tcgen05.mma = UTCHMMA*,
tcgen05.ld / st = LDTM / STTM, cp.async.bulk.tensor = UTMALDG / UTMASTG / UTMAREDG, mbarrier = SYNCS, legacy mma.sync = HMMA)
of the built library.   python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "vfmseg_b200" / "lib" / "libvfmseg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "LDS", "STS", "LDL", "STL"]
per = OrderedDict()
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(.*", "", fn)
        per[fn] = Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[fn]["instructions"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[fn][k] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"# {Path(lib).name}: {len(per)} kernels, arch {sorted(set(arch))}; counts of SASS instructions per kernel (cuobjdump -sass)")
print("# " + " ".join(f"{k:>8s}" for k in ["instr"] + KEYS) + "  kernel")
tot = Counter()
for fn, c in per.items():
    tot.update(c)
    print("  " + " ".join(f"{c.get(k, 0):8d}" for k in ["instructions"] + KEYS) + "  " + fn)
print("# total " + " ".join(f"{k}={tot.get(k, 0)}" for k in KEYS))
