import sys, ctypes, json
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import _C, build
lib = sys.argv[1]
build.LIB_PATH = ROOT / "vfmseg_b200" / "lib" / lib
_C.LIB_PATH = build.LIB_PATH
from vfmseg_b200 import ops
qkv = (torch.randn(18 * 1025, 3072, device="cuda") * 0.7).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = ops.attention_fwd(qkv, 18, 1025, 16)
q, k, v = qkv.float().view(18, 1025, 3, 16, 64).permute(2, 0, 3, 1, 4)
ref = ((q[:2] @ k[:2].transpose(-1, -2)).softmax(-1) @ v[:2]).transpose(1, 2).reshape(2 * 1025, 1024)
err = (out[:2050].float() - ref).abs().max().item()
# wide score range: the running max jumps by more than 2^24 after the first tile (exercises the rescale / redo path)
qkv2 = qkv.clone(); qkv2[:, :2048] *= 2.2
out2 = ops.attention_fwd(qkv2, 18, 1025, 16)
q, k, v = qkv2.float().view(18, 1025, 3, 16, 64).permute(2, 0, 3, 1, 4)
ref2 = ((q[:2] @ k[:2].transpose(-1, -2)).softmax(-1) @ v[:2]).transpose(1, 2).reshape(2 * 1025, 1024)
err2 = (out2[:2050].float() - ref2).abs().max().item()
print("wide-range max abs err", err2, "finite", bool(torch.isfinite(out2.float()).all()))
for _ in range(3): ops.attention_fwd(qkv, 18, 1025, 16)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_(); s.record(); ops.attention_fwd(qkv, 18, 1025, 16); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print(lib, "median ms", sorted(ts)[5], "max abs err", err)
# per-kernel split (CUDA events around each launch, C ABI profiler)
import ctypes
lib_ = _C.load()
lib_.vfm_prof_enable(1)
for _ in range(10):
    flush.zero_(); ops.attention_fwd(qkv, 18, 1025, 16)
buf = ctypes.create_string_buffer(1 << 16)
lib_.vfm_prof_report(buf, len(buf))
lib_.vfm_prof_enable(0)
print(buf.value.decode().strip().replace("\n", " | "))
for mode in (1, 2, 3):
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); s.record(); ops.attention_fwd(qkv, 18, 1025, 16, mode); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print("mode", mode, "median ms", sorted(ts)[5])
