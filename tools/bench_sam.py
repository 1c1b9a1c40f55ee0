"""Throughput of BASELINE config 5 at the shipped crop (SAM ViT-H/16 + LinearHead slide inference, 1024x2048, crop 512 /
stride 320, configs/_base_/models/lora_sam_linear.py) on one GPU, with the per-kernel split of the C-ABI profiler.
Development tool; bench.py is the contract benchmark (config 2)."""
import ctypes, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vfmseg_b200
from vfmseg_b200 import synthetic, _C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
CROP = int(sys.argv[2]) if len(sys.argv) > 2 else 512          # 512: the shipped config; 1024: BASELINE config 5 (stride 682)
STRIDE = 320 if CROP == 512 else 682
cfg = synthetic.sam_model_config(img_size=CROP, crop_size=(CROP, CROP), stride=(STRIDE, STRIDE))
model = vfmseg_b200.MODELS.build(dict(cfg))
model.load_state_dict(synthetic.synthetic_sam_state_dict(cfg, seed=0), strict=False)
model = model.cuda().eval()
img = synthetic.synthetic_images(B, 1024, 2048, seed=11).cuda()
for _ in range(2):
    model.predict_labels(img)
torch.cuda.synchronize()
lib = _C.load()
n0 = lib.vfm_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K):
    model.predict_labels(img)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
launches = (lib.vfm_launch_count() - n0) // K
# algorithmic FLOPs per 512^2 crop: 1024 tokens, C = 1280, 32 blocks (28 windowed: 9 windows x 196 tokens incl. padding
# are NOT counted, only the 1024 real tokens' projections; attention counted on the real key/query pairs it must form)
from oracle import torch_ref
g = CROP // 16
C, T, D = 1280, g * g, 32
n_win = ((g + 13) // 14) ** 2
n_crops = len(torch_ref.slide_boxes(1024, 2048, (CROP, CROP), (STRIDE, STRIDE)))
gemm = D * 2 * T * C * (3 * C + C + 4 * C + 4 * C)
att = 28 * 4 * 16 * n_win * 196 * 196 * 80 + 4 * 4 * 16 * T * T * 80
head = 2 * T * 4 * C * C + 2 * T * C * 2 * C + 2 * 4 * T * (C // 2) * C + 2 * 16 * T * (C // 4) * 19
flop_img = n_crops * (gemm + att + 2 * T * 768 * C + head)
lib.vfm_prof_enable(1)
model.predict_labels(img)
buf = ctypes.create_string_buffer(1 << 16)
lib.vfm_prof_report(buf, len(buf))
lib.vfm_prof_enable(0)
prof = {}
for line in buf.value.decode().strip().splitlines():
    name, cnt, t = line.split(",")
    prof[name] = (int(cnt), round(float(t), 3))
print(json.dumps({"config": f"SAM ViT-H/16 + LinearHead slide 1024x2048 crop {CROP} stride {STRIDE} ({n_crops} crops)", "images_per_step": B, "ms_per_step": round(ms, 3),
                  "images_per_s": round(B / ms * 1e3, 2), "tflops": round(flop_img * B / ms / 1e9, 1), "flop_per_image": flop_img,
                  "launches_per_step": launches, "kernels_ms(one step, profiled)": prof}))
