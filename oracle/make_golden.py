"""ORACLE — TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz by running the REFERENCE's own
modules (imported unmodified from /root/reference through oracle/ref_shim.py) on the deterministic
synthetic weights/images of vfmseg_b200.synthetic. Run here (the reference tree does not exist on the
GPU box):   python -m oracle.make_golden

Files
  tiny_slide.npz : tiny model (dim 256, depth 4), 1 image 80x112, crop 64 / stride 43 (2x3 windows):
                   reference slide_inference logits, argmax labels, IoUMetric areas, mIoU summary.
  tiny_whole.npz : same model, whole_inference on a non-square 64x96 input (bicubic pos-embed path).
  vitl_crop.npz  : config 1 — DINOv2 ViT-L/16 + LoRA + LinearHead, one 512x512 crop, fp32 CPU:
                   low-res logits subsampled [:, ::2, ::2] and per-tap statistics.
  vitl_crop_probe.npz / vitl_full_probe.npz : the "trained-network-like" recipe (synthetic.region_images + a classifier fitted by
                   oracle/probe.py): fitted conv_seg, the reference's slide_inference labels for the WHOLE image and its
                   logits subsampled; vitl_full_probe is BASELINE config 2 at full size (one 1024x2048 image, 18 windows,
                   ViT-L). The >= 99.9 % raw label-agreement bar of north_star is asserted against these.
Inputs and weights are NOT stored: they are regenerated from seeds (torch CPU generators).
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shim, torch_ref  # noqa: E402
from vfmseg_b200 import synthetic  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def build_reference(cfg, sd):
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if "num_batches_tracked" not in m], missing
    return model.eval()


def tiny_slide():
    cfg = synthetic.tiny_config()
    sd = synthetic.synthetic_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd)
    img = synthetic.synthetic_images(1, 80, 112, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    with torch.no_grad():
        logits = model.inference(x, [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)])
        samples = model.postprocess_result(logits.clone())
    pred = samples[0].pred_sem_seg.data  # int64 [1,H,W]
    gt = synthetic.synthetic_labels(1, 80, 112, 19, seed=4321)
    metric_mod = ref_shim.load("dg_metrics")
    m = metric_mod.DGIoUMetric(dataset_keys=["citys"], ignore_index=255, iou_metrics=["mIoU"])
    m.dataset_meta = dict(classes=list(range(19)))
    m.process({}, [dict(pred_sem_seg=dict(data=pred), gt_sem_seg=dict(data=gt[0:1].long()), seg_map_path="x/citys/y.png", img_path="a.png")])
    key, ai, au, ap, al = m.results[0]
    summary = m.compute_metrics(m.results)
    np.savez_compressed(GOLDEN / "tiny_slide.npz", logits=logits.numpy(), labels=pred.numpy().astype(np.uint8),
                        area_intersect=ai.numpy(), area_union=au.numpy(), area_pred=ap.numpy(), area_label=al.numpy(),
                        mIoU=np.float64(summary["citys_mIoU"]), mAcc=np.float64(summary["citys_mAcc"]), aAcc=np.float64(summary["citys_aAcc"]))
    print("tiny_slide", logits.shape, summary)


def tiny_whole():
    cfg = synthetic.tiny_config(mode="whole")
    sd = synthetic.synthetic_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd)
    img = synthetic.synthetic_images(1, 64, 96, seed=77)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    with torch.no_grad():
        logits = model.inference(x, [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)])
    np.savez_compressed(GOLDEN / "tiny_whole.npz", logits=logits.numpy())
    print("tiny_whole", logits.shape)


def vitl_crop():
    cfg = synthetic.model_config()
    sd = synthetic.synthetic_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd)
    img = synthetic.synthetic_images(1, 512, 512, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    t0 = time.time()
    with torch.no_grad():
        feats = model.extract_feat(x)
        low = model.decode_head(feats)
    print(f"vitl_crop reference forward {time.time() - t0:.1f}s", low.shape)
    stats = np.array([[f.mean().item(), f.std().item(), f.abs().max().item()] for f in feats], dtype=np.float64)
    np.savez_compressed(GOLDEN / "vitl_crop.npz", lowres_sub=low[0, :, ::2, ::2].numpy(), tap_stats=stats,
                        lowres_argmax=low[0].argmax(0).numpy().astype(np.uint8))


def _oracle_cfg(cfg):
    bb, lc = cfg["backbone"], cfg["Lora_config"]
    return dict(depth=bb["depth"], num_heads=bb["num_heads"], patch=bb["patch_size"], out_indices=tuple(bb["out_indices"]),
                lora_scale=lc["lora_alpha"] / lc["r"], groups=cfg["decode_head"]["norm_cfg"]["num_groups"])


def _probe_golden(name, cfg, H, W, cell, crop, stride, sub_logits, img_seed=11):
    """Fit the probe on the oracle's merged features, then run the REFERENCE's slide_inference with it."""
    from oracle import probe
    sd = synthetic.synthetic_state_dict(cfg, seed=0)
    img, planted = synthetic.region_images(1, H, W, seed=img_seed, cell=cell)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    t0 = time.time()
    feats = probe.merged_features(x, sd, _oracle_cfg(cfg), crop, stride)
    w, b = probe.fit_probe(feats, planted)
    del feats
    sd = synthetic.with_probe_classifier(sd, w, b)
    model = build_reference(cfg, sd)
    with torch.no_grad():
        logits = model.inference(x, [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)])
        with torch.autocast("cpu", dtype=torch.bfloat16):   # how far a bf16 run of the reference itself lands from its fp32 labels
            lo = torch_ref.slide_inference(x, torch_ref.split_state_dict(sd), _oracle_cfg(cfg), crop=crop, stride=stride).float()
    labels = logits.argmax(1)
    agree_planted = (labels == planted.long()).float().mean().item()
    agree_bf16 = (labels == lo.argmax(1)).float().mean().item()
    print(f"{name}: {time.time() - t0:.0f}s; {probe.margin_report(logits)}; labels == planted {agree_planted:.4f}; "
          f"reference fp32 vs oracle under bf16 autocast: label agreement {agree_bf16:.5f}")
    np.savez_compressed(GOLDEN / f"{name}.npz", conv_seg_weight=w.numpy(), conv_seg_bias=b.numpy(),
                        labels=labels[0].numpy().astype(np.uint8), logits_sub=logits[0, :, ::sub_logits, ::sub_logits].numpy().astype(np.float16),
                        sub=np.int64(sub_logits), img_seed=np.int64(img_seed), cell=np.int64(cell),
                        agree_bf16_autocast=np.float64(agree_bf16), agree_planted=np.float64(agree_planted))


def vitl_crop_probe():   # one 512 x 512 window of the real architecture (smoke(), single-crop tests)
    _probe_golden("vitl_crop_probe", synthetic.model_config(), 512, 512, 128, (512, 512), (341, 341), 4)


def vitl_full_probe():
    _probe_golden("vitl_full_probe", synthetic.model_config(), 1024, 2048, 256, (512, 512), (341, 341), 8)


def build_reference_ms(cfg, sd):
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_ms_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if "num_batches_tracked" not in m], missing
    return model.eval()


def ms_gate_constants(model, x, cfg):
    """threshold / conf for the tiny model: random weights never reach the shipped 0.968 / 0.8, so pick a threshold at
    the median max-softmax and a conf in the widest gap of the per-window fractions — both branches of
    Ms_VFM_encoder_decoder.py:449-452 run and no decision is borderline."""
    import torch.nn.functional as F
    with torch.no_grad():
        lr = F.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
        seg = model.whole_inference(lr, [dict(img_shape=x.shape[2:], ori_shape=x.shape[2:])])
    thr = float(torch.softmax(seg, 1).max(1)[0].median())
    fr = []
    for (y1, y2, x1, x2) in torch_ref.slide_boxes(x.shape[2], x.shape[3], cfg["test_cfg"]["crop_size"], cfg["test_cfg"]["stride"]):
        c = seg[:, :, y1:y2, x1:x2]
        fr.append((torch.softmax(c, 1).max(1)[0] > thr).float().mean().item())
    srt = sorted(fr)
    gaps = [(srt[i + 1] - srt[i], i) for i in range(len(srt) - 1)]
    g, i = max(gaps[len(gaps) // 4: 3 * len(gaps) // 4 + 1] or gaps)
    return thr, 0.5 * (srt[i] + srt[i + 1]), fr


def tiny_ms():
    """tiny_ms.npz: MsVFMEncoderDecoder (tiny backbone + LinearHead + VFMHead/MaskTransformerDecoder depth 2),
    mode ms_slide_inference on one 128x192 image (crop 64 / stride 43 -> 3x4 windows): reference logits, labels, the
    per-window confidence fractions and refine decisions, and one standalone VFMHead.forward call."""
    cfg = synthetic.tiny_ms_config()
    sd = synthetic.synthetic_ms_state_dict(cfg, seed=0)
    model = build_reference_ms(cfg, sd)
    img = synthetic.synthetic_images(1, 128, 192, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    thr, conf, fr = ms_gate_constants(model, x, cfg)
    model.test_cfg["threadshod"], model.test_cfg["conf"] = thr, conf
    metas = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)]
    with torch.no_grad():
        logits = model.inference(x, metas)
        # standalone head call (VFMHead.forward with the decoder's mask off), window (0, 0)
        model.aux_decoder.transformer_decoder.mask_enable = False
        lr = torch.nn.functional.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
        seg = model.whole_inference(lr, metas)
        feats = model.extract_feat(x[:, :, :64, :64])
        head_out = model.aux_decoder(feats, seg[:, :, :64, :64])
    refined = [f < conf for f in fr]
    print("tiny_ms thr", thr, "conf", conf, "refined", refined)
    assert any(refined) and not all(refined)
    np.savez_compressed(GOLDEN / "tiny_ms.npz", logits=logits.numpy().astype(np.float16), labels=logits.argmax(1).numpy().astype(np.uint8),
                        threshold=np.float64(thr), conf=np.float64(conf), fracs=np.array(fr), refined=np.array(refined),
                        head_out=head_out.numpy())


def tiny_eva():
    """tiny_eva.npz: EncoderDecoder(LoRABackbone(EVA2 dim 256 / depth 4 / 4 heads, 4x4 grid), LinearHead), slide inference on
    one 80x112 image (crop 64 / stride 43): reference logits + the four feature maps of window (0, 0)."""
    cfg = synthetic.tiny_eva_config()
    sd = synthetic.synthetic_eva_state_dict(cfg, seed=0)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_eva_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "rope" not in m and "num_batches_tracked" not in m], (missing, unexpected)
    img = synthetic.synthetic_images(1, 80, 112, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    metas = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)]
    with torch.no_grad():
        logits = model.inference(x, metas)
        feats = model.extract_feat(x[:, :, :64, :64])
    np.savez_compressed(GOLDEN / "tiny_eva.npz", logits=logits.numpy().astype(np.float16), feats=torch.stack(list(feats)).numpy())
    print("tiny_eva", logits.shape, logits.std())


def tiny_sam():
    """tiny_sam.npz: EncoderDecoder(LoRABackbone(SAMViT dim 640 / depth 4 / 8 heads x 80, 16x16 tokens, window 14, global
    blocks 1 and 3), LinearHead), slide inference on one 256x384 image (crop 256 / stride 171): reference logits + the four
    feature maps of window (0, 0)."""
    cfg = synthetic.tiny_sam_config()
    sd = synthetic.synthetic_sam_state_dict(cfg, seed=0)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_sam_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m], (missing, unexpected)
    img = synthetic.synthetic_images(1, 256, 384, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    metas = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)]
    with torch.no_grad():
        logits = model.inference(x, metas)
        feats = model.extract_feat(x[:, :, :256, :256])
    np.savez_compressed(GOLDEN / "tiny_sam.npz", logits=logits.numpy().astype(np.float16), feats=torch.stack(list(feats)).numpy().astype(np.float16))
    print("tiny_sam", logits.shape, logits.std(), [float(f.std()) for f in feats])


def _real_crop_golden(name, cfg, sd, builder, crop):
    """One real-dimension window through the reference's backbone + LinearHead: low-res logits [:, ::2, ::2] and argmax."""
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = builder(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "rope" not in m and "num_batches_tracked" not in m], (missing, unexpected)
    model.eval()
    img = synthetic.synthetic_images(1, crop, crop, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    t0 = time.time()
    with torch.no_grad():
        low = model.decode_head(model.extract_feat(x))
    # yardstick: the SAME reference modules under CPU bf16 autocast (what "bf16 tolerance" means for this architecture)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        lo = model.decode_head(model.extract_feat(x)).float()
    rms = low.pow(2).mean().sqrt()
    err = (lo - low).abs()
    within = (err <= 2e-2 * low.abs() + 2e-2 * rms).float().mean().item()
    rel = ((lo - low).pow(2).mean().sqrt() / rms).item()
    agree = (lo.argmax(1) == low.argmax(1)).float().mean().item()
    print(f"{name}: reference forward {time.time() - t0:.1f}s", tuple(low.shape), float(low.std()),
          f"| reference under bf16 autocast: within band {within:.4f}, rel rms {rel:.4f}, label agreement {agree:.4f}")
    np.savez_compressed(GOLDEN / f"{name}.npz", lowres_sub=low[0, :, ::2, ::2].numpy().astype(np.float16),
                        lowres_argmax=low[0].argmax(0).numpy().astype(np.uint8), autocast_within=np.float64(within),
                        autocast_rel_rms=np.float64(rel), autocast_label_agreement=np.float64(agree))


def eva_crop():   # BASELINE config 4 at real dimensions: EVA02-L/16, one 512 x 512 window
    cfg = synthetic.eva_model_config()
    _real_crop_golden("eva_crop", cfg, synthetic.synthetic_eva_state_dict(cfg, seed=0), ref_shim.build_reference_eva_segmentor, 512)


def sam_crop():   # BASELINE config 5 at real dimensions: SAM ViT-H/16, one 1024 x 1024 window (4096 tokens, 25 padded 14 x 14 windows)
    cfg = synthetic.sam_model_config(img_size=1024, crop_size=(1024, 1024), stride=(682, 682))
    _real_crop_golden("sam_crop", cfg, synthetic.synthetic_sam_state_dict(cfg, seed=0), ref_shim.build_reference_sam_segmentor, 1024)


def ms_crop():
    """BASELINE config 3 at real dimensions: the VFMHead refinement of ONE 512 x 512 window of a 1024 x 2048 image — coarse
    whole-image pass at 512 x 1024, context = the window's slice of the up-sampled coarse logits, ViT-L features of the
    window, VFMHead.forward with the decoder's mask off: refined low-res logits [19, 32, 32] + the coarse logits."""
    import torch.nn.functional as F
    cfg = synthetic.ms_model_config()
    sd = synthetic.synthetic_ms_state_dict(cfg, seed=0)
    model = build_reference_ms(cfg, sd)
    img = synthetic.synthetic_images(1, 1024, 2048, seed=1234)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    metas = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)]
    t0 = time.time()
    with torch.no_grad():
        model.aux_decoder.transformer_decoder.mask_enable = False
        lr = F.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
        seg = model.whole_inference(lr, metas)                      # [1, 19, 1024, 2048]
        y1, x1 = 320, 640                                           # window (1, 2) of the stride-320 grid
        feats = model.extract_feat(x[:, :, y1:y1 + 512, x1:x1 + 512])
        head_out = model.aux_decoder(feats, seg[:, :, y1:y1 + 512, x1:x1 + 512])
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        seg_b = model.whole_inference(lr, metas).float()
        head_b = model.aux_decoder(model.extract_feat(x[:, :, y1:y1 + 512, x1:x1 + 512]), seg_b[:, :, y1:y1 + 512, x1:x1 + 512]).float()
    rms = head_out.pow(2).mean().sqrt()
    within = ((head_b - head_out).abs() <= 2e-2 * head_out.abs() + 2e-2 * rms).float().mean().item()
    rel = ((head_b - head_out).pow(2).mean().sqrt() / rms).item()
    print(f"ms_crop: {time.time() - t0:.1f}s", tuple(head_out.shape), float(head_out.std()),
          f"| reference under bf16 autocast: within band {within:.4f}, rel rms {rel:.4f}")
    np.savez_compressed(GOLDEN / "ms_crop.npz", head_out=head_out[0].numpy().astype(np.float32), window=np.array([y1, x1]),
                        coarse_sub=seg[0, :, ::16, ::16].numpy().astype(np.float16), autocast_within=np.float64(within),
                        autocast_rel_rms=np.float64(rel))


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["tiny_slide", "tiny_whole", "vitl_crop", "tiny_ms", "tiny_eva", "tiny_sam", "vitl_crop_probe", "vitl_full_probe", "eva_crop", "sam_crop", "ms_crop"]
    for w in which:
        globals()[w]()
