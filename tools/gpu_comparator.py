"""GPU-side comparator: the SAME modules as the reference's hot path — DINOv2 ViT-L/16 with
the LoRA merged + LinearHead — written as plain PyTorch eager ops in bf16 on the B200 (cuBLASLt GEMMs through F.linear,
F.scaled_dot_product_attention with the cuDNN backend, F.layer_norm / F.gelu / F.group_norm), the 18 windows of an image
batched into one pass, merge + argmax in torch. This is what "run the reference on the GPU with library kernels" gives; the
hand-written path has to beat it to count.

    python tools/gpu_comparator.py [--images 2] [--steps 10]

Prints one JSON line: images/s, ms per image, and the TFLOP/s of its qkv GEMM and attention calls timed in isolation at
the same shapes. bench.py --gpu-comparator embeds that line under the key "gpu_comparator". No kernel of this repo is used.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _boxes(H, W, crop, stride):
    hg = max(H - crop + stride - 1, 0) // stride + 1
    wg = max(W - crop + stride - 1, 0) // stride + 1
    out = []
    for hi in range(hg):
        for wi in range(wg):
            y2, x2 = min(hi * stride + crop, H), min(wi * stride + crop, W)
            out.append((max(y2 - crop, 0), max(x2 - crop, 0)))
    return out


class EagerVitL:
    """bf16 weights on the device, LoRA merged into qkv (what any deployment of the reference would do)."""

    def __init__(self, sd, dev, depth=24, heads=16, lora_scale=1.0, out_indices=(7, 11, 15, 23)):
        p = "backbone.base_model.model."
        g = lambda k: sd[p + k].to(dev)
        self.depth, self.heads, self.out_indices = depth, heads, out_indices
        self.pe_w, self.pe_b = g("patch_embed.proj.weight").bfloat16(), g("patch_embed.proj.bias").bfloat16()
        self.cls, self.pos = g("cls_token").bfloat16(), g("pos_embed").bfloat16()
        self.blocks = []
        for i in range(depth):
            b = f"blocks.{i}."
            w = g(b + "attn.qkv.base_layer.weight") + lora_scale * g(b + "attn.qkv.lora_B.default.weight") @ g(b + "attn.qkv.lora_A.default.weight")
            self.blocks.append(dict(
                n1w=g(b + "norm1.weight").bfloat16(), n1b=g(b + "norm1.bias").bfloat16(), qkv_w=w.bfloat16(), qkv_b=g(b + "attn.qkv.base_layer.bias").bfloat16(),
                proj_w=g(b + "attn.proj.weight").bfloat16(), proj_b=g(b + "attn.proj.bias").bfloat16(), ls1=g(b + "ls1.gamma").bfloat16(),
                n2w=g(b + "norm2.weight").bfloat16(), n2b=g(b + "norm2.bias").bfloat16(), fc1_w=g(b + "mlp.fc1.weight").bfloat16(),
                fc1_b=g(b + "mlp.fc1.bias").bfloat16(), fc2_w=g(b + "mlp.fc2.weight").bfloat16(), fc2_b=g(b + "mlp.fc2.bias").bfloat16(),
                ls2=g(b + "ls2.gamma").bfloat16()))
        h = "decode_head."
        hd = lambda k: sd[h + k].to(dev).bfloat16()
        self.head = {k: hd(k) for k in ("fusion_conv.conv.weight", "fusion_conv.gn.weight", "fusion_conv.gn.bias", "output_upscaling.0.weight",
                                        "output_upscaling.0.bias", "output_upscaling.1.weight", "output_upscaling.1.bias",
                                        "output_upscaling.1.running_mean", "output_upscaling.1.running_var", "output_upscaling.3.weight",
                                        "output_upscaling.3.bias", "conv_seg.weight", "conv_seg.bias")}

    def crops_lowres(self, x):   # x: bf16 [n, 3, 512, 512] normalised
        n = x.shape[0]
        t = F.conv2d(x, self.pe_w, self.pe_b, stride=16).flatten(2).transpose(1, 2)
        t = torch.cat((self.cls.expand(n, -1, -1), t), 1) + self.pos
        feats = []
        for i, b in enumerate(self.blocks):
            h = F.layer_norm(t, (1024,), b["n1w"], b["n1b"], 1e-6)
            qkv = F.linear(h, b["qkv_w"], b["qkv_b"]).view(n, -1, 3, self.heads, 64).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(n, -1, 1024)
            t = t + F.linear(a, b["proj_w"], b["proj_b"]) * b["ls1"]
            h = F.layer_norm(t, (1024,), b["n2w"], b["n2b"], 1e-6)
            t = t + F.linear(F.gelu(F.linear(h, b["fc1_w"], b["fc1_b"])), b["fc2_w"], b["fc2_b"]) * b["ls2"]
            if i in self.out_indices:
                feats.append(t[:, 1:].transpose(1, 2).reshape(n, 1024, 32, 32))
        hd = self.head
        y = F.conv2d(torch.cat(feats, 1), hd["fusion_conv.conv.weight"])
        y = F.relu(F.group_norm(y, 32, hd["fusion_conv.gn.weight"], hd["fusion_conv.gn.bias"], 1e-5))
        y = F.conv_transpose2d(y, hd["output_upscaling.0.weight"], hd["output_upscaling.0.bias"], stride=2)
        y = F.gelu(F.batch_norm(y, hd["output_upscaling.1.running_mean"], hd["output_upscaling.1.running_var"], hd["output_upscaling.1.weight"],
                                hd["output_upscaling.1.bias"], False, 0.0, 1e-5))
        y = F.gelu(F.conv_transpose2d(y, hd["output_upscaling.3.weight"], hd["output_upscaling.3.bias"], stride=2))
        return F.conv2d(y, hd["conv_seg.weight"], hd["conv_seg.bias"])   # [n, 19, 128, 128]

    def slide_labels(self, img_u8, crop=512, stride=341):
        B, _, H, W = img_u8.shape
        mean = torch.tensor(MEAN, device=img_u8.device).view(1, 3, 1, 1)
        std = torch.tensor(STD, device=img_u8.device).view(1, 3, 1, 1)
        x = ((img_u8.flip(1).float() - mean) / std).bfloat16()
        boxes = _boxes(H, W, crop, stride)
        crops = torch.cat([x[:, :, y:y + crop, xx:xx + crop] for (y, xx) in boxes], 0)   # window-major: [n_win * B, ...]
        low = self.crops_lowres(crops).float()
        up = F.interpolate(low, size=(crop, crop), mode="bilinear", align_corners=False)
        preds = torch.zeros(B, up.shape[1], H, W, device=x.device)
        count = torch.zeros(B, 1, H, W, device=x.device)
        for k, (y, xx) in enumerate(boxes):
            preds[:, :, y:y + crop, xx:xx + crop] += up[k * B:(k + 1) * B]
            count[:, :, y:y + crop, xx:xx + crop] += 1
        return (preds / count).argmax(1).to(torch.uint8)


def _time(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def run(images: int = 2, steps: int = 10, dev=None) -> dict:
    from torch.nn.attention import SDPBackend, sdpa_kernel
    from vfmseg_b200 import synthetic
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    cfg = synthetic.model_config()
    net = EagerVitL(synthetic.synthetic_state_dict(cfg, seed=0), dev)
    img = synthetic.synthetic_images(images, 1024, 2048, seed=1000).to(dev)
    gt = synthetic.synthetic_labels(images, 1024, 2048, 19, seed=2000).to(dev)

    def step():
        with torch.no_grad(), sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION]):
            lab = net.slide_labels(img)
            keep = gt != 255   # confusion matrix, as the timed step of bench.py has it
            torch.bincount(gt[keep].long() * 19 + lab[keep].long(), minlength=19 * 19)
    ms = _time(step, steps)
    n = 18 * images
    M = n * 1025
    a = torch.randn(M, 1024, device=dev, dtype=torch.bfloat16)
    w = torch.randn(3072, 1024, device=dev, dtype=torch.bfloat16)
    bq = torch.randn(3072, device=dev, dtype=torch.bfloat16)
    ms_qkv = _time(lambda: F.linear(a, w, bq), 20)
    q = torch.randn(n, 16, 1025, 64, device=dev, dtype=torch.bfloat16)
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        ms_att = _time(lambda: F.scaled_dot_product_attention(q, q, q), 20)
    return {"impl": "PyTorch eager bf16 on the same GPU: cuBLASLt F.linear + cuDNN SDPA, LoRA merged, 18 windows x images batched per pass, torch merge/argmax/bincount",
            "images_per_step": images, "value": round(images / ms * 1e3, 3), "unit": "images/s", "ms_per_step": round(ms, 3),
            "qkv_gemm_tflops": round(2.0 * M * 1024 * 3072 / ms_qkv / 1e9, 1), "attention_tflops": round(4.0 * n * 16 * 1025 * 1025 * 64 / ms_att / 1e9, 1),
            "torch": torch.__version__}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(run(a.images, a.steps)))
