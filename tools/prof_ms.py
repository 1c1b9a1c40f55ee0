"""Per-kernel split (C-ABI profiler) of one ms_slide_inference step, all windows refined. Development tool."""
import ctypes, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vfmseg_b200
from vfmseg_b200 import synthetic, _C
cfg = synthetic.ms_model_config()
model = vfmseg_b200.MODELS.build(dict(cfg))
model.load_state_dict(synthetic.synthetic_ms_state_dict(cfg, seed=0), strict=False)
model = model.cuda().eval()
model.test_cfg.conf = 1.5
img = synthetic.synthetic_images(2, 1024, 2048, seed=11).cuda()
for _ in range(2): model.predict_labels(img)
torch.cuda.synchronize()
lib = _C.load(); lib.vfm_prof_enable(1)
model.predict_labels(img)
buf = ctypes.create_string_buffer(1 << 16); lib.vfm_prof_report(buf, len(buf)); lib.vfm_prof_enable(0)
tot = 0
for line in buf.value.decode().strip().splitlines():
    n, c, t = line.split(","); tot += float(t); print(f"{n:28s} {int(c):4d} {float(t):8.3f} ms")
print("total", round(tot, 3))
