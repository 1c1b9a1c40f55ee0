"""Debug: event timeline of one attention CTA (needs the -DVFM_EPI_TIMING build in lib/libvfmseg_b200_dbg.so)."""
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
tr = (ctypes.c_longlong * 384)()
lib.vfm_debug_att_trace(tr)
t = [[tr[j * 24 + e] for e in range(24)] for j in range(16)]
t0 = min(v for row in t for v in row if v > 0)
names = ["mma:top", "mma:p_ok", "mma:pv_iss", "mma:s_iss", "-", "-", "-", "-",
         "A:wait_s", "A:s_ok", "A:S_ld", "A:max", "A:jchk", "A:arrived", "A:pp_go", "A:exp_end",
         "B:wait_s", "B:s_ok", "B:S_ld", "B:max", "B:jchk", "B:arrived", "B:pp_go", "B:exp_end"]
print("event timeline of CTA 150 (cycles from first event)")
print("tile " + " ".join(f"{n:>9s}" for n in names if n != "-"))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{(t[j][e]-t0) if t[j][e] > 0 else -1:9d}" for e in range(24) if names[e] != "-"))
