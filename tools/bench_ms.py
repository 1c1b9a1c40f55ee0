"""Throughput of BASELINE config 3 (MsVFMEncoderDecoder ms_slide_inference, ViT-L/16, 1024x2048) on one GPU for a given
fraction of refined windows (random-init weights never pass the shipped 0.968 gate, so the fraction is forced through
test_cfg.conf). Development tool; bench.py is the contract benchmark (config 2)."""
import json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vfmseg_b200
from vfmseg_b200 import synthetic, _C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = synthetic.ms_model_config()
model = vfmseg_b200.MODELS.build(dict(cfg))
model.load_state_dict(synthetic.synthetic_ms_state_dict(cfg, seed=0), strict=False)
model = model.cuda().eval()
img = synthetic.synthetic_images(B, 1024, 2048, seed=11).cuda()
for conf, name in ((1.5, "all 18 windows refined"), (0.0, "no window refined (stage 0 only)")):
    model.test_cfg.conf = conf
    for _ in range(3):
        model.predict_labels(img)
    torch.cuda.synchronize()
    n0 = _C.load().vfm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        model.predict_labels(img)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(json.dumps({"config": "ms_slide_inference ViT-L 1024x2048 crop 512 stride 320", "case": name, "images_per_step": B,
                      "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 2),
                      "launches_per_step": (_C.load().vfm_launch_count() - n0) // K}))
