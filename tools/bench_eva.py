"""Throughput of BASELINE config 4 (EVA02-L/16 + LinearHead slide inference, 1024x2048, crop 512 / stride 320) on one GPU.
Development tool; bench.py is the contract benchmark (config 2)."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vfmseg_b200
from vfmseg_b200 import synthetic, _C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synthetic.eva_model_config()
model = vfmseg_b200.MODELS.build(dict(cfg))
model.load_state_dict(synthetic.synthetic_eva_state_dict(cfg, seed=0), strict=False)
model = model.cuda().eval()
img = synthetic.synthetic_images(B, 1024, 2048, seed=11).cuda()
for _ in range(3):
    model.predict_labels(img)
torch.cuda.synchronize()
lib = _C.load()
n0 = lib.vfm_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    model.predict_labels(img)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
flop_img = 18 * (24 * (2 * 1025 * 1024 * (3072 + 1024 + 3 * 2730) + 4 * 16 * 1025 * 1025 * 64) + 2 * 1024 * 768 * 1024 + 17.34e9)
print(json.dumps({"config": "EVA02-L/16 + LinearHead slide 1024x2048 crop 512 stride 320", "images_per_step": B, "ms_per_step": round(ms, 3),
                  "images_per_s": round(B / ms * 1e3, 2), "tflops": round(flop_img * B / ms / 1e9, 1),
                  "launches_per_step": (lib.vfm_launch_count() - n0) // K}))
