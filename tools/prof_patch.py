"""Isolated timing of the patch-embed GEMM (36 windows: M = 36864, N = 1024, K = 768). Development tool."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vfmseg_b200 import ops  # noqa: E402

n, P, C, K = 36, 1024, 1024, 768
a = torch.randn(n * P, K, device="cuda").bfloat16()
w = (torch.randn(C, K, device="cuda") * 0.02).bfloat16()
b = torch.randn(C, device="cuda")
pos = torch.randn(P + 1, C, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.gemm_patch_embed(a, w, b, pos, n, P)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_patch_embed(a, w, b, pos, n, P); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
print(f"patch_embed GEMM: {ms * 1e3:.1f} us median, {2.0 * n * P * C * K / ms / 1e9:.1f} TFLOP/s")
