"""Debug: per-phase cycle counts of the GEMM epilogue (needs the -DVFM_EPI_TIMING build)."""
import ctypes, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import _C, build
build.LIB_PATH = ROOT / "vfmseg_b200" / "lib" / "libvfmseg_b200_dbg.so"
_C.LIB_PATH = build.LIB_PATH
from vfmseg_b200 import ops
lib = _C.load()
dev = "cuda"
M, C, H = 18450, 1024, 4096
bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
f32 = lambda *s: torch.randn(*s, device=dev)
a, a4 = bf(M, C), bf(M, H)
cases = {
    "qkv": lambda: ops.gemm_bias_bf16(a, bf(3 * C, C), f32(3 * C)),
    "proj": lambda: ops.gemm_bias_ls_residual_(f32(M, C), a, bf(C, C), f32(C), f32(C)),
    "fc1": lambda: ops.gemm_bias_gelu_bf16(a, bf(H, C), f32(H)),
    "fc2": lambda: ops.gemm_bias_ls_residual_(f32(M, C), a4, bf(C, H), f32(C), f32(C)),
}
Mp = 36 * 1024
ap = bf(Mp, 768)
cases["patch_embed"] = lambda: ops.gemm_patch_embed(ap, bf(C, 768), f32(C), f32(1025, C), 36, 1024)
cases["convt1"] = lambda: ops.gemm_convt2x2_gelu(bf(Mp, C), bf(2048, C), f32(2048), 512, 32, 32)
cases["convt2"] = lambda: ops.gemm_convt2x2_gelu(bf(4 * Mp, 512), bf(1024, 512), f32(1024), 256, 64, 64)
out = (ctypes.c_ulonglong * 5)()
for name, fn in cases.items():
    fn(); lib.vfm_debug_epi(out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); lib.vfm_debug_epi(out)
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms (includes operand allocation)")
    v = list(out)
    tiles = max(v[4], 1)
    print(f"{name}: per tile (cycles, warp 4 avg): wait_full {v[0]/tiles:.0f}  tmem_ld {v[1]/tiles:.0f}  transpose {v[2]/tiles:.0f}  elem+global {v[3]/tiles:.0f}  tiles {tiles}")

qkv = bf(M, 3 * C)
for _ in range(2):
    ops.attention_fwd(qkv, 18, 1025, 16)
tr = (ctypes.c_longlong * 192)()
lib.vfm_debug_att_trace(tr)
t = [[tr[j * 12 + e] for e in range(12)] for j in range(16)]
t0 = min(v for row in t[:16] for v in row[3:9] if v > 0)
names = ["-", "-", "-", "wait_s(sm)", "s_ready(sm)", "S_loaded(sm)", "max_done(sm)", "P_buf_free(sm)", "exp_done(sm)"]
print("event timeline of CTA 300 (cycles from first event):")
print("tile " + " ".join(f"{n:>17s}" for n in names))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{(t[j][e]-t0) if t[j][e] > 0 else -1:17d}" for e in range(9)))
