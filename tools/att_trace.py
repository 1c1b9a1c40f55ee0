"""Debug: event timelines of the two co-resident first-wave attention CTAs of SM 5 (needs a -DVFM_EPI_TIMING build)."""
import ctypes, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import _C, build
build.LIB_PATH = ROOT / "vfmseg_b200" / "lib" / (sys.argv[1] if len(sys.argv) > 1 else "libvfmseg_b200_dbg.so")
_C.LIB_PATH = build.LIB_PATH
from vfmseg_b200 import ops
lib = _C.load()
qkv = (torch.randn(18 * 1025, 3072, device="cuda") * 0.7).to(torch.bfloat16)
for _ in range(2):
    ops.attention_fwd(qkv, 18, 1025, 16, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
tr = (ctypes.c_longlong * 644)()
lib.vfm_debug_att_trace(tr)
t = [[[tr[(s * 20 + j) * 16 + e] for e in range(16)] for j in range(20)] for s in range(2)]
t0 = min(v for sl in t for row in sl for v in row if v > 0)
names = ["mma:top", "mma:p_ok", "-", "mma:iss", "-", "-", "-", "-",
         "wait_s", "s_ok", "S_ld", "max", "jchk", "arrived", "-", "exp_end"]
print(build.LIB_PATH.name, "event timelines (cycles from first event), SM 5, slots = the two co-resident CTAs")
for s in range(2):
    print(f"slot {s}")
    print("tile " + " ".join(f"{n:>9s}" for n in names if n != "-") + "   exp_len  period")
    prev = None
    for j in range(18):
        r = t[s][j]
        line = f"{j:4d} " + " ".join(f"{(r[e]-t0) if r[e] > 0 else -1:9d}" for e in range(16) if names[e] != "-")
        line += f" {r[15]-r[12]:9d}" + (f" {r[8]-prev:7d}" if prev else "")
        prev = r[8]
        print(line)
clk, ns = tr[642] - tr[640], tr[643] - tr[641]
if ns > 0:
    print(f"slot-0 CTA life: {clk} clk in {ns} ns -> SM clock {clk / ns:.3f} GHz under this kernel")
