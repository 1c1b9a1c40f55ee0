"""A/B of the two tiled merge kernels on BASELINE config 2's tail (2 images of 1024x2048, 18 windows each, 19 classes):
VFM_MERGE_MODE=1 = round-1 window-major tile kernel, default = class-major kernel. CUDA events, L2 flushed between launches.
Prints one JSON line per case. Optional arguments: comma-separated VFM_MERGE_MODE values for the slide merge, then for the stage-1 merge
of ms_inference ("none" skips that part)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfmseg_b200 import ops
from vfmseg_b200.engine import slide_boxes

bx = torch.tensor(slide_boxes(1024, 2048, (512, 512), (341, 341)), dtype=torch.int32).cuda()
low = torch.randn(36, 19, 128, 128, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
algo_bytes = low.numel() * 4 + 2 * 1024 * 2048   # low-res logits read once + uint8 labels written
slide_modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "3", "0"]
for mode in ([] if slide_modes == ["none"] else slide_modes):
    os.environ["VFM_MERGE_MODE"] = mode
    for want in (False, True):
        f = lambda: ops.slide_merge_argmax(low, bx, 2, (512, 512), (1024, 2048), want_logits=want)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        med = sorted(ts)[len(ts) // 2]
        b = algo_bytes + (2 * 19 * 1024 * 2048 * 4 if want else 0)
        print(json.dumps({"kernel": "slide_merge_tile (window-major)" if mode == "1" else "slide_merge_class (class-major), CTAs per SM >= " + {"2": "2", "3": "3", "0": "4"}[mode],
                          "want_logits": want, "median_us": round(med * 1e3, 1), "min_us": round(min(ts) * 1e3, 1),
                          "algorithmic_GBps": round(b / med / 1e6, 1)}))
os.environ.pop("VFM_MERGE_MODE", None)

# Stage-1 merge of ms_inference (BASELINE config 3): 2 images, coarse logits 128x256 (x8), 24 of 36 windows refined (32x32, x16)
ms_modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "3", "0"]
if ms_modes != ["none"]:
    low0 = torch.randn(2, 19, 128, 256, device="cuda")
    g = torch.Generator().manual_seed(1)
    mask = torch.rand(2, 18, generator=g) < 0.67
    ref_index = torch.full((2, 18), -1, dtype=torch.int32)
    ref_index[mask] = torch.arange(int(mask.sum()), dtype=torch.int32)
    refined = torch.randn(int(mask.sum()), 19, 32, 32, device="cuda")
    ri = ref_index.cuda()
    names = {"1": "ms_merge_argmax_kernel (per-pixel gather)", "3": "ms_merge_class_kernel, CTAs per SM >= 3", "0": "ms_merge_class_kernel, CTAs per SM >= 2"}
    for mode in ms_modes:
        name = names[mode]
        os.environ["VFM_MERGE_MODE"] = mode
        f = lambda: ops.ms_merge_argmax(low0, refined, ri, bx, (512, 512), (1024, 2048))
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        print(json.dumps({"kernel": name, "refined_windows": int(mask.sum()), "median_us": round(sorted(ts)[10] * 1e3, 1),
                          "min_us": round(min(ts) * 1e3, 1)}))
    os.environ.pop("VFM_MERGE_MODE", None)
