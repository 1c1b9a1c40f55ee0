"""One config-2-shaped attention call (ncu target)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import ops
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
crops = int(sys.argv[2]) if len(sys.argv) > 2 else 18
qkv = (torch.randn(crops * 1025, 3072, device="cuda") * 0.7).to(torch.bfloat16)
for _ in range(3):
    ops.attention_fwd(qkv, crops, 1025, 16, mode)
torch.cuda.synchronize()
print("ok")
