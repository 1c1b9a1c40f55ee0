"""CPU: the oracle restatement (oracle/torch_ref.py) against the golden vectors produced by the
reference's OWN modules (oracle/make_golden.py), and against the live reference when the tree exists."""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_shim, torch_ref
from vfmseg_b200 import synthetic

GOLDEN = Path(__file__).parent / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _cfg_of(cfg):
    bb, lc = cfg["backbone"], cfg["Lora_config"]
    return dict(depth=bb["depth"], num_heads=bb["num_heads"], patch=bb["patch_size"], out_indices=tuple(bb["out_indices"]),
                lora_scale=lc["lora_alpha"] / lc["r"], groups=cfg["decode_head"]["norm_cfg"]["num_groups"])


def _oracle_model(cfg, seed=0):
    sd = synthetic.synthetic_state_dict(cfg, seed=seed)
    return torch_ref.split_state_dict(sd), _cfg_of(cfg)


def test_slide_boxes_match_survey():
    boxes = torch_ref.slide_boxes(1024, 2048, (512, 512), (341, 341))
    assert len(boxes) == 18
    assert sorted({b[0] for b in boxes}) == [0, 341, 512]
    assert sorted({b[2] for b in boxes}) == [0, 341, 682, 1023, 1364, 1536]
    assert len(torch_ref.slide_boxes(1024, 2048, (512, 512), (320, 320))) == 18
    assert torch_ref.slide_boxes(64, 64, (64, 64), (43, 43)) == [(0, 64, 0, 64)]


def test_tiny_slide_against_reference_golden():
    g = np.load(GOLDEN / "tiny_slide.npz")
    cfg = synthetic.tiny_config()
    sd, oc = _oracle_model(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 80, 112, seed=1234), MEAN, STD, True)
    with torch.no_grad():
        logits = torch_ref.slide_inference(x, sd, oc, crop=(64, 64), stride=(43, 43))
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-4)
    labels = torch_ref.postprocess(logits)[0].numpy().astype(np.uint8)
    assert (labels == g["labels"]).mean() >= 0.9995
    # metric: identical prediction map -> bit-exact areas and summary
    gt = synthetic.synthetic_labels(1, 80, 112, 19, seed=4321)
    pred = torch.from_numpy(g["labels"].astype(np.int64))
    ai, au, ap, al = torch_ref.intersect_and_union(pred, gt.long(), 19, 255)
    for a, k in ((ai, "area_intersect"), (au, "area_union"), (ap, "area_pred"), (al, "area_label")):
        assert np.array_equal(a.numpy(), g[k]), k
    cm = torch_ref.confusion_matrix_np(g["labels"], gt.numpy(), 19, 255)
    ci, cu, cp, cl = torch_ref.areas_from_confusion(cm, 19)
    assert np.array_equal(ci, g["area_intersect"].astype(np.int64)) and np.array_equal(cu, g["area_union"].astype(np.int64))
    assert np.array_equal(cp, g["area_pred"].astype(np.int64)) and np.array_equal(cl, g["area_label"].astype(np.int64))
    s = torch_ref.total_area_to_metrics(ci, cu, cp, cl)
    assert s["mIoU"] == pytest.approx(float(g["mIoU"]), abs=1e-6)
    assert s["mAcc"] == pytest.approx(float(g["mAcc"]), abs=1e-6)
    assert s["aAcc"] == pytest.approx(float(g["aAcc"]), abs=1e-6)


def test_tiny_whole_nonsquare_against_reference_golden():
    g = np.load(GOLDEN / "tiny_whole.npz")
    cfg = synthetic.tiny_config(mode="whole")
    sd, oc = _oracle_model(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 64, 96, seed=77), MEAN, STD, True)
    with torch.no_grad():
        logits = torch_ref.whole_inference(x, sd, oc)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-4)


def test_vitl_crop_against_reference_golden():
    """Config 1: DINOv2 ViT-L/16 + LoRA + LinearHead, one 512x512 crop, fp32 on CPU."""
    g = np.load(GOLDEN / "vitl_crop.npz")
    cfg = synthetic.model_config()
    (bb, hd), oc = _oracle_model(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 512, 512, seed=1234), MEAN, STD, True)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        feats = torch_ref.dino_forward(x, bb, depth=oc["depth"], num_heads=oc["num_heads"], out_indices=oc["out_indices"],
                                       lora_scale=oc["lora_scale"])
        low = torch_ref.linear_head_forward(feats, hd)
    stats = np.array([[f.mean().item(), f.std().item(), f.abs().max().item()] for f in feats])
    np.testing.assert_allclose(stats, g["tap_stats"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(low[0, :, ::2, ::2].numpy(), g["lowres_sub"], rtol=2e-3, atol=2e-3)
    assert (low[0].argmax(0).numpy() == g["lowres_argmax"]).mean() >= 0.999


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference_modules():
    """Different seed / shapes than the goldens, straight against the reference's classes."""
    import tempfile
    cfg = synthetic.tiny_config(stride=(32, 32))
    sd = synthetic.synthetic_state_dict(cfg, seed=5)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "bb.pth")
        torch.save(synthetic.backbone_checkpoint_from(sd), ck)
        ref = ref_shim.build_reference_segmentor(cfg, ck)
    ref.load_state_dict(sd, strict=False)
    x = torch_ref.preprocess(synthetic.synthetic_images(2, 96, 96, seed=9), MEAN, STD, True)
    meta = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)] * 2
    with torch.no_grad():
        want = ref.inference(x, meta)
        got = torch_ref.slide_inference(x, torch_ref.split_state_dict(sd), _cfg_of(cfg), crop=(64, 64), stride=(32, 32))
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ coarse-to-fine path (BASELINE config 3)
def _ms_cfg_of(cfg):
    bb, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    tr = cfg["aux_head"]["transformer"]
    return dict(depth=bb["depth"], num_heads=bb["num_heads"], patch=bb["patch_size"], out_indices=tuple(bb["out_indices"]),
                lora_scale=lc["lora_alpha"] / lc["r"], groups=32, aux_heads=tr["n_heads"], aux_depth=tr["depth"])


def test_tiny_ms_against_reference_golden():
    """oracle ms_inference / vfm_head_forward vs the reference's MsVFMEncoderDecoder + VFMHead (tiny_ms.npz)."""
    g = np.load(GOLDEN / "tiny_ms.npz")
    cfg = synthetic.tiny_ms_config()
    sd3 = torch_ref.split_ms_state_dict(synthetic.synthetic_ms_state_dict(cfg, seed=0))
    oc = _ms_cfg_of(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 128, 192, seed=1234), MEAN, STD, True)
    with torch.no_grad():
        logits, info = torch_ref.ms_inference(x, sd3, oc, crop=(64, 64), stride=(43, 43), threshold=float(g["threshold"]),
                                              conf=float(g["conf"]), return_info=True)
    assert info["refined"] == list(g["refined"])
    np.testing.assert_allclose(np.array(info["fracs"]), g["fracs"], atol=2e-3)
    np.testing.assert_allclose(logits.numpy(), g["logits"].astype(np.float32), rtol=2e-3, atol=2e-3)   # golden stored as fp16
    assert (logits.argmax(1).numpy() == g["labels"]).mean() >= 0.999
    # standalone head
    with torch.no_grad():
        seg = torch.nn.functional.interpolate(info["low0"], size=x.shape[2:], mode="bilinear", align_corners=False)
        feats = torch_ref.dino_forward(x[:, :, :64, :64], sd3[0], depth=oc["depth"], num_heads=oc["num_heads"], patch=16,
                                       out_indices=oc["out_indices"], lora_scale=oc["lora_scale"])
        head = torch_ref.vfm_head_forward(feats, seg[:, :, :64, :64], sd3[2], heads=oc["aux_heads"], depth=oc["aux_depth"])
    np.testing.assert_allclose(head.numpy(), g["head_out"], rtol=1e-4, atol=1e-4)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_ms_oracle_against_live_reference():
    import tempfile
    cfg = synthetic.tiny_ms_config(aux_depth=1)
    sd = synthetic.synthetic_ms_state_dict(cfg, seed=3)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "b.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_ms_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 96, 128, seed=5), MEAN, STD, True)
    thr, conf = 0.09, 0.5
    model.test_cfg["threadshod"], model.test_cfg["conf"] = thr, conf
    metas = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)]
    with torch.no_grad():
        ref = model.inference(x, metas)
        got = torch_ref.ms_inference(x, torch_ref.split_ms_state_dict(sd), _ms_cfg_of(cfg), crop=(64, 64), stride=(43, 43),
                                     threshold=thr, conf=conf)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ EVA02 backbone (BASELINE config 4)
def _eva_parts(cfg, seed=0):
    sd = synthetic.synthetic_eva_state_dict(cfg, seed=seed)
    pre = "backbone.model.base_model.model."
    bb = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    hd = {k[len("decode_head."):]: v for k, v in sd.items() if k.startswith("decode_head.")}
    b, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    oc = dict(depth=b["depth"], num_heads=b["num_heads"], out_indices=tuple(b["out_indices"]), lora_scale=lc["lora_alpha"] / lc["r"], groups=32)
    return sd, bb, hd, oc


def test_tiny_eva_against_reference_golden():
    """oracle eva_forward / slide loop vs the reference's own EVA2 + LoRABackbone + LinearHead (tiny_eva.npz)."""
    g = np.load(GOLDEN / "tiny_eva.npz")
    cfg = synthetic.tiny_eva_config()
    _, bb, hd, oc = _eva_parts(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 80, 112, seed=1234), MEAN, STD, True)
    with torch.no_grad():
        feats = torch_ref.eva_forward(x[:, :, :64, :64], bb, depth=oc["depth"], num_heads=oc["num_heads"], out_indices=oc["out_indices"],
                                      lora_scale=oc["lora_scale"])
        logits = torch_ref.eva_slide_inference(x, bb, hd, oc, crop=(64, 64), stride=(43, 43))
    np.testing.assert_allclose(torch.stack(feats).numpy(), g["feats"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(logits.numpy(), g["logits"].astype(np.float32), rtol=2e-3, atol=2e-3)


# ------------------------------------------------------------------ SAM ViT backbone (BASELINE config 5)
def _sam_parts(cfg, seed=0):
    sd = synthetic.synthetic_sam_state_dict(cfg, seed=seed)
    pre = "backbone.model.base_model.model."
    bb = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    hd = {k[len("decode_head."):]: v for k, v in sd.items() if k.startswith("decode_head.")}
    b, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    oc = dict(depth=b["depth"], num_heads=b["num_heads"], out_indices=tuple(b["out_indices"]), window_size=b["window_size"],
              global_attn_indexes=tuple(b["global_attn_indexes"]), lora_scale=lc["lora_alpha"] / lc["r"], groups=32)
    return sd, bb, hd, oc


def test_tiny_sam_against_reference_golden():
    """oracle sam_forward / slide loop vs the reference's own SAMViT + LoRABackbone + LinearHead (tiny_sam.npz): windowed
    blocks with zero-padded windows, global blocks with interpolated rel-pos tables, decomposed rel-pos bias on unscaled q."""
    g = np.load(GOLDEN / "tiny_sam.npz")
    cfg = synthetic.tiny_sam_config()
    _, bb, hd, oc = _sam_parts(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 256, 384, seed=1234), MEAN, STD, True)
    with torch.no_grad():
        feats = torch_ref.sam_forward(x[:, :, :256, :256], bb, depth=oc["depth"], num_heads=oc["num_heads"], window_size=oc["window_size"],
                                      global_attn_indexes=oc["global_attn_indexes"], out_indices=oc["out_indices"], lora_scale=oc["lora_scale"])
        logits = torch_ref.sam_slide_inference(x, bb, hd, oc, crop=(256, 256), stride=(171, 171))
    np.testing.assert_allclose(torch.stack(feats).numpy(), g["feats"].astype(np.float32), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(logits.numpy(), g["logits"].astype(np.float32), rtol=2e-3, atol=2e-3)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_sam_oracle_against_live_reference():
    """Same restatement against the reference's sam_vit.py imported here, fp32 tolerance, on a second seed and a
    non-square token grid is impossible (fixed pos_embed) — so a different window layout: 20x20 tokens (pad to 28)."""
    import os, tempfile
    cfg = synthetic.tiny_sam_config(img_size=320, crop_size=(320, 320), depth=2, global_attn_indexes=(1,), out_indices=(0, 1))
    sd, bb, hd, oc = _sam_parts(cfg, seed=3)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "backbone.pth")
        torch.save(synthetic.ms_backbone_checkpoint_from(sd), ck)
        model = ref_shim.build_reference_sam_segmentor(cfg, ck)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m]
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 320, 320, seed=9), MEAN, STD, True)
    with torch.no_grad():
        ref = model.extract_feat(x)
        got = torch_ref.sam_forward(x, bb, depth=oc["depth"], num_heads=oc["num_heads"], window_size=oc["window_size"],
                                    global_attn_indexes=oc["global_attn_indexes"], out_indices=oc["out_indices"], lora_scale=oc["lora_scale"])
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ flip test-time augmentation
@pytest.mark.skipif(not ref_shim.available(), reason="reference tree absent")
@pytest.mark.parametrize("flip", [True, False])
def test_tta_flip_oracle_against_live_reference(flip):
    """oracle.tta_flip_combine against the reference's own HRDAEncoderDecoder.slide_inference
    (hrda_encoder_decoder.py:194-229, unmodified), the per-window encode_decode supplied by the oracle."""
    import torch.nn as nn
    hrda = ref_shim.load("models.segmentors.hrda_encoder_decoder")
    cfg = synthetic.tiny_config(stride=(32, 32))
    sd, ocfg = _oracle_model(cfg, seed=3)
    seg = hrda.HRDAEncoderDecoder.__new__(hrda.HRDAEncoderDecoder)
    nn.Module.__init__(seg)
    seg.test_cfg = ref_shim.ConfigDict(mode="slide", crop_size=(64, 64), stride=(32, 32), test_time_aug=True, flip=flip)
    seg.test_time_aug, seg.flip = True, flip           # what __init__ copies out of test_cfg (:114-115)
    seg.num_classes = seg.out_channels = cfg["decode_head"]["num_classes"]
    seg.align_corners = False
    seg.encode_decode = lambda crop, metas: torch_ref.encode_decode(crop, sd, ocfg)[0]
    x = torch_ref.preprocess(synthetic.synthetic_images(2, 80, 112, seed=21), MEAN, STD, True)
    meta = [dict(ori_shape=x.shape[2:], img_shape=x.shape[2:], pad_shape=x.shape[2:], padding_size=[0] * 4)] * 2
    with torch.no_grad():
        want = seg.slide_inference(x, meta)
        got = torch_ref.tta_flip_combine(lambda im: torch_ref.slide_inference(im, sd, ocfg, crop=(64, 64), stride=(32, 32)), x, flip)
    assert torch.equal(got, want)
    if flip:   # the augmentation is not a no-op for this model, and its result is mirror-equivariant by construction
        plain = torch_ref.slide_inference(x, sd, ocfg, crop=(64, 64), stride=(32, 32))
        assert not torch.allclose(got, plain, atol=1e-3)
        got_m = torch_ref.tta_flip_combine(lambda im: torch_ref.slide_inference(im, sd, ocfg, crop=(64, 64), stride=(32, 32)),
                                           torch.flip(x, [3]), True)
        torch.testing.assert_close(torch.flip(got_m, [3]), got, rtol=0, atol=1e-6)


def test_probe_golden_pins_the_oracle_at_real_dimensions():
    """vitl_crop_probe.npz comes from the REFERENCE's modules (oracle/make_golden.py); the oracle restatement with the same
    fitted classifier must reproduce its labels (fp32 vs fp32: differences only from summation order) — this pins the
    oracle at ViT-L dimensions, not just on the tiny model — and the golden must have the trained-like margins the
    >= 99.9 % label bar relies on."""
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "vitl_crop_probe.npz")
    cfg = synthetic.model_config()
    sd = synthetic.with_probe_classifier(synthetic.synthetic_state_dict(cfg, seed=0), g["conv_seg_weight"], g["conv_seg_bias"])
    img, planted = synthetic.region_images(1, 512, 512, seed=int(g["img_seed"]), cell=int(g["cell"]))
    x = torch_ref.preprocess(img, [123.675, 116.28, 103.53], [58.395, 57.12, 57.375], True)
    bb, lc = cfg["backbone"], cfg["Lora_config"]
    oc = dict(depth=bb["depth"], num_heads=bb["num_heads"], patch=16, out_indices=tuple(bb["out_indices"]),
              lora_scale=lc["lora_alpha"] / lc["r"], groups=32)
    with torch.no_grad():
        ref = torch_ref.slide_inference(x, torch_ref.split_state_dict(sd), oc, crop=(512, 512), stride=(341, 341))
    assert (ref[0].argmax(0).numpy() == g["labels"]).mean() >= 0.9999
    sub = int(g["sub"])
    gl = torch.from_numpy(g["logits_sub"].astype(np.float32))
    assert (ref[0, :, ::sub, ::sub] - gl).abs().max().item() <= 2e-3 * gl.abs().max().item() + 2e-3   # fp16 storage
    assert float(g["agree_planted"]) > 0.8 and float(g["agree_bf16_autocast"]) > 0.999


def _ms_head_conditioning():
    """Config-3 refinement head on the tiny model: output error of the fp32 oracle head when only its INPUTS carry bf16
    error (backbone features and coarse logits computed under CPU bf16 autocast), and of the whole chain under autocast —
    the reference's own modules at the precision the GPU path computes in."""
    import torch.nn.functional as F
    cfg = synthetic.tiny_ms_config()
    sd = synthetic.synthetic_ms_state_dict(cfg, seed=0)
    sd3 = torch_ref.split_ms_state_dict(sd)
    bb = cfg["backbone"]["backbone"]
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 128, 192, seed=1234), [123.675, 116.28, 103.53], [58.395, 57.12, 57.375], True)
    kw = dict(depth=bb["depth"], num_heads=bb["num_heads"], out_indices=tuple(bb["out_indices"]), lora_scale=2.0)

    def inputs(bf16):
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=bf16):
            lr = F.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
            low0 = torch_ref.linear_head_forward(torch_ref.dino_forward(lr, sd3[0], **kw), sd3[1]).float()
            feats = [f.float() for f in torch_ref.dino_forward(x[:, :, :64, :64], sd3[0], **kw)]
        return feats, F.interpolate(low0, size=x.shape[2:], mode="bilinear", align_corners=False)[:, :, :64, :64]

    def head(feats, ctx, bf16):
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=bf16):
            return torch_ref.vfm_head_forward(feats, ctx, sd3[2], heads=8, depth=2).float()

    f32, c32 = inputs(False)
    fb, cb = inputs(True)
    ref = head(f32, c32, False)
    rms = ref.pow(2).mean().sqrt()

    def score(got):
        e = (got - ref).abs()
        return (e <= 2e-2 * ref.abs() + 2e-2 * rms).float().mean().item(), ((got - ref).pow(2).mean().sqrt() / rms).item()
    return dict(inputs_only=score(head(fb, cb, False)), all_bf16=score(head(fb, cb, True)),
                ctx_err=((cb - c32).pow(2).mean().sqrt() / c32.pow(2).mean().sqrt()).item())


def test_ms_head_amplifies_bf16_input_error_beyond_the_logit_band():
    """Why tests/test_ms_e2e_gpu.py cannot assert 99.9 % of the REFINED logits inside the 2e-2 band:
    the reference's VFMHead with random weights roughly doubles the relative error of its coarse-logit input (0.7 % in,
    1.2-1.3 % out). With an EXACT fp32 head fed by a bf16 backbone only ~97 % of the outputs stay inside the band, and the
    reference's own modules run end to end under bf16 autocast keep ~89 %. No precision choice inside the head can beat the
    first number; the GPU tests assert the path is at least as accurate as the second (measured live there)."""
    r = _ms_head_conditioning()
    assert 0.003 < r["ctx_err"] < 0.02
    assert r["inputs_only"][0] < 0.99 and r["inputs_only"][1] > 1.5 * r["ctx_err"]
    assert r["all_bf16"][0] < r["inputs_only"][0]
