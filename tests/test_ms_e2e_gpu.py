"""GPU parity of the coarse-to-fine path (BASELINE config 3): registered MsVFMEncoderDecoder / LoRABackbone / VFMHead /
MaskTransformerDecoder -> engine -> C ABI -> sm_100a kernels, against the golden vectors produced by the reference's own
modules (tests/golden/tiny_ms.npz, oracle/make_golden.py) and the oracle restatement (oracle/torch_ref.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from test_e2e_gpu import _check_labels, _check_logits

pytestmark = pytest.mark.gpu

# The refinement path stacks ~40 bf16 round trips behind the backbone (k2s2 convs + GroupNorms over as few as 2
# channels x 64 pixels, LayerNorms, two attentions and a GEGLU per decoder block); tools/ms_head_debug.py shows every
# stage adding 0.1-0.2 % relative rms error and no stage standing out (head output: 1.4 % of rms). The elementwise
# band of test_e2e_gpu.py (2e-2*|ref| + 2e-2*rms) therefore holds for ~95-97 % of the refined logits, not 99.9 %, on
# random-init weights; asserted here: that fraction >= 0.95, relative rms error <= 2 %; labels: 99.9 % agreement on the
# pixels the reference decides by more than the logit tolerance (measured 100 %), raw agreement >= 98.5 % (measured
# 98.9-99.2 %: random weights leave 16-23 % of the pixels with a top-2 margin inside the tolerance).
MS_FRAC = 0.95
# Round 2: the error is a property of the reference's head, not of this implementation — tests/test_oracle.py::
# test_ms_head_amplifies_bf16_input_error_beyond_the_logit_band measures, on this very input, that the fp32 oracle head fed
# with a bf16-autocast backbone keeps only ~97 % of its outputs inside the band (the head doubles the relative error of its
# coarse-logit input) and that the reference's modules under end-to-end bf16 autocast keep ~89 % (1.9-2.0 % rms). The
# asserts below therefore also compare against that yardstick, computed live: at least as many logits in the band and no
# larger rms error than the reference itself at bf16.


def _rel_rms(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()

GOLDEN = Path(__file__).parent / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _build_ms(cfg, seed=0):
    import vfmseg_b200
    from vfmseg_b200 import synthetic
    sd = synthetic.synthetic_ms_state_dict(cfg, seed=seed)
    model = vfmseg_b200.MODELS.build(dict(cfg))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m], (missing, unexpected)
    return model.cuda().eval(), sd


def _ms_cfg_of(cfg):
    bb, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    tr = cfg["aux_head"]["transformer"]
    return dict(depth=bb["depth"], num_heads=bb["num_heads"], patch=bb["patch_size"], out_indices=tuple(bb["out_indices"]),
                lora_scale=lc["lora_alpha"] / lc["r"], groups=32, aux_heads=tr["n_heads"], aux_depth=tr["depth"])


def test_tiny_ms_vs_reference_golden():
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "tiny_ms.npz")
    cfg = synthetic.tiny_ms_config(threshold=float(g["threshold"]), conf=float(g["conf"]))
    model, _ = _build_ms(cfg)
    img = synthetic.synthetic_images(1, 128, 192, seed=1234)
    labels, logits, info = model._ms(img.cuda(), True, "image")
    # the gate: same per-window fractions, same refine / reuse decisions as the reference loop
    fr = info["counts"].float().cpu().numpy()[0] / (64 * 64)
    np.testing.assert_allclose(fr, g["fracs"], atol=0.03)
    assert ((info["ref_index"].cpu().numpy()[0] >= 0) == g["refined"]).all()
    assert info["n_refined"] == int(g["refined"].sum())
    ref = torch.from_numpy(g["logits"].astype(np.float32))
    _check_logits(logits, ref, "tiny ms_slide_inference vs reference golden", frac=MS_FRAC)
    assert _rel_rms(logits, ref) <= 0.02
    _check_labels(labels, ref, "tiny ms labels vs reference golden", raw_min=0.985, top2_min=0.999)
    # mmseg contract: test_step -> SegDataSample with int64 labels and fp32 logits
    out = model.test_step(dict(inputs=[img[0]]))
    assert out[0].pred_sem_seg.data.dtype == torch.int64 and out[0].seg_logits.data.shape == (19, 128, 192)
    assert model.aux_decoder.transformer_decoder.mask_enable is True   # restored, Ms_VFM_encoder_decoder.py:463-464


def test_vfm_head_module_vs_reference_golden():
    """VFMHead.forward(inputs, seg_logits) used standalone (registry API), decoder mask off."""
    from oracle import torch_ref
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "tiny_ms.npz")
    cfg = synthetic.tiny_ms_config()
    model, sd = _build_ms(cfg)
    sd3 = torch_ref.split_ms_state_dict(sd)
    oc = _ms_cfg_of(cfg)
    x = torch_ref.preprocess(synthetic.synthetic_images(1, 128, 192, seed=1234), MEAN, STD, True)
    with torch.no_grad():
        lr = torch.nn.functional.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
        low0 = torch_ref.linear_head_forward(torch_ref.dino_forward(lr, sd3[0], depth=oc["depth"], num_heads=oc["num_heads"],
                                                                    out_indices=oc["out_indices"], lora_scale=oc["lora_scale"]), sd3[1])
        seg = torch.nn.functional.interpolate(low0, size=x.shape[2:], mode="bilinear", align_corners=False)
        feats = torch_ref.dino_forward(x[:, :, :64, :64], sd3[0], depth=oc["depth"], num_heads=oc["num_heads"],
                                       out_indices=oc["out_indices"], lora_scale=oc["lora_scale"])
    head = model.aux_decoder
    with pytest.raises(NotImplementedError):
        head([f.cuda() for f in feats], seg[:, :, :64, :64].cuda())     # mask_enable=True is a random augmentation
    head.transformer_decoder.mask_enable = False
    out = head([f.cuda() for f in feats], seg[:, :, :64, :64].cuda().contiguous())
    assert out.shape == (1, 19, 4, 4)
    _check_logits(out, torch.from_numpy(g["head_out"]), "VFMHead.forward vs reference golden", frac=0.90)
    assert _rel_rms(out, torch.from_numpy(g["head_out"])) <= 0.02
    # yardstick: the reference's modules under bf16 autocast on the same input (see the note at MS_FRAC)
    from test_oracle import _ms_head_conditioning
    y = _ms_head_conditioning()
    ref = torch.from_numpy(g["head_out"]).float()
    got = out.float().cpu()
    rms = ref.pow(2).mean().sqrt()
    ours = ((got - ref).abs() <= 2e-2 * ref.abs() + 2e-2 * rms).float().mean().item()
    print(f"VFMHead: within band {ours:.4f} / rel rms {_rel_rms(out, ref):.4f}; reference under bf16 autocast {y['all_bf16'][0]:.4f} / "
          f"{y['all_bf16'][1]:.4f}; exact head on bf16 inputs {y['inputs_only'][0]:.4f} / {y['inputs_only'][1]:.4f}")
    assert ours >= y["all_bf16"][0] and _rel_rms(out, ref) <= y["all_bf16"][1]


def test_ms_real_dimension_window_vs_reference_golden():
    """BASELINE config 3 at REAL dimensions: coarse pass of a 1024 x 2048 image at 512 x 1024 and the VFMHead refinement of
    one 512 x 512 window (ViT-L features, 3-block MaskTransformerDecoder, 256 channels); golden = the reference's own
    Ms_VFM_encoder_decoder.py / VFMHead.py / Transformer.py modules in fp32 (oracle/make_golden.py:ms_crop). Bars as for the
    tiny model (see the note at MS_FRAC): the head amplifies the bf16 error of its coarse-logit input."""
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "ms_crop.npz")
    cfg = synthetic.ms_model_config()
    model, _ = _build_ms(cfg)
    img = synthetic.synthetic_images(1, 1024, 2048, seed=1234).cuda()
    model.test_cfg.conf = 1.5                     # refine every window
    labels, logits, info = model._ms(img, True, "image")
    from vfmseg_b200.engine import slide_boxes
    boxes = slide_boxes(1024, 2048, (512, 512), (320, 320))
    assert info["n_refined"] == len(boxes) == 18 and info["refined"].shape == (18, 19, 32, 32)
    y1, x1 = [int(v) for v in g["window"]]
    k = boxes.index((y1, x1))
    ref = torch.from_numpy(g["head_out"])
    got = info["refined"][k]
    # yardstick stored with the golden: the reference's own modules under CPU bf16 autocast keep 92.9 % inside the band (rel rms 1.70 %)
    _check_logits(got, ref, "config 3 real dimensions: refined window vs reference golden", frac=max(0.90, float(g["autocast_within"])))
    assert _rel_rms(got, ref) <= float(g["autocast_rel_rms"])
    up = torch.nn.functional.interpolate(info["low0"], size=(1024, 2048), mode="bilinear", align_corners=False)
    _check_logits(up[0, :, ::16, ::16], torch.from_numpy(g["coarse_sub"].astype(np.float32)), "config 3 real dimensions: coarse logits vs reference golden")


def test_ms_batch_and_modes_vs_oracle():
    from oracle import torch_ref
    from vfmseg_b200 import synthetic
    cfg = synthetic.tiny_ms_config(threshold=0.09, conf=0.5, aux_depth=1)
    model, sd = _build_ms(cfg, seed=3)
    sd3 = torch_ref.split_ms_state_dict(sd)
    oc = _ms_cfg_of(cfg)
    img = synthetic.synthetic_images(2, 96, 128, seed=5)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    # per-image gating == the reference run once per image (its test loop has batch_size 1)
    labels, logits = model.predict_labels(img.cuda(), want_logits=True)
    with torch.no_grad():
        refs = [torch_ref.ms_inference(x[b:b + 1], sd3, oc, crop=(64, 64), stride=(43, 43), threshold=0.09, conf=0.5,
                                       return_info=True) for b in range(2)]
    ref = torch.cat([r[0] for r in refs])
    margins = np.abs(np.array([r[1]["fracs"] for r in refs]) - 0.5)
    if margins.min() > 0.03:      # decisions not borderline for bf16 logits
        _check_logits(logits, ref, "tiny ms batch 2 vs oracle", frac=MS_FRAC)
        _check_labels(labels, ref, "tiny ms batch 2 labels vs oracle", raw_min=0.985, top2_min=0.999)
    # literal batch semantics of Ms_VFM_encoder_decoder.py:448 (one decision per window from the batch mean)
    with torch.no_grad():
        refb, infob = torch_ref.ms_inference(x, sd3, oc, crop=(64, 64), stride=(43, 43), threshold=0.09, conf=0.5, return_info=True)
    if np.abs(np.array(infob["fracs"]) - 0.5).min() > 0.03:
        _check_logits(model.inference(x.cuda(), None), refb, "tiny ms inference() batch gate vs oracle", frac=MS_FRAC)
    # hr_slide_inference = plain slide with the main head (:287)
    model.test_cfg.mode = "hr_slide_inference"
    cfg_lin = dict(depth=oc["depth"], num_heads=oc["num_heads"], patch=16, out_indices=oc["out_indices"], lora_scale=oc["lora_scale"], groups=32)
    with torch.no_grad():
        ref_hr = torch_ref.slide_inference(x, (sd3[0], sd3[1]), cfg_lin, crop=(64, 64), stride=(43, 43))
    _check_logits(model.predict_labels(img.cuda(), want_logits=True)[1], ref_hr, "hr_slide_inference vs oracle")
    # lr_slide_inference (:282-285)
    model.test_cfg.mode = "lr_slide_inference"
    img_big = synthetic.synthetic_images(1, 160, 192, seed=6)
    xb = torch_ref.preprocess(img_big, MEAN, STD, True)
    with torch.no_grad():
        xl = torch.nn.functional.interpolate(xb, scale_factor=0.5, mode="bilinear", align_corners=False)
        ref_lr = torch.nn.functional.interpolate(torch_ref.slide_inference(xl, (sd3[0], sd3[1]), cfg_lin, crop=(64, 64), stride=(43, 43)),
                                                 scale_factor=2, mode="bilinear", align_corners=False)
    _check_logits(model.predict_labels(img_big.cuda(), want_logits=True)[1], ref_lr, "lr_slide_inference vs oracle")
    model.test_cfg.mode = "msfull_slide_inference"
    with pytest.raises(NotImplementedError):
        model.predict_labels(img.cuda())


def test_full_size_ms_properties():
    """BASELINE config 3 at full size (ViT-L/16, 1024x2048, crop 512 / stride 320 -> 18 windows): size-independent
    properties of the coarse-to-fine merge. conf = 0 -> no window is refined and the result is exactly the x8 bilinear
    upsampling of the stage-0 logits; conf > 1 -> every window is refined and the merge equals the plain slide merge
    of the refined window logits."""
    import time
    from vfmseg_b200 import ops, synthetic
    cfg = synthetic.ms_model_config()
    model, _ = _build_ms(cfg)
    img = synthetic.synthetic_images(1, 1024, 2048, seed=11).cuda()
    model.test_cfg.conf = 0.0
    labels0, logits0, info0 = model._ms(img, True, "image")
    assert info0["n_refined"] == 0 and info0["low0"].shape == (1, 19, 128, 256)
    up = ops.resize_bilinear(info0["low0"], (1024, 2048))
    assert torch.equal(logits0, up)
    assert torch.equal(labels0.long(), up.argmax(1))
    model.test_cfg.conf = 1.5
    torch.cuda.synchronize()
    t0 = time.time()
    labels1, logits1, info1 = model._ms(img, True, "image")
    torch.cuda.synchronize()
    print(f"full-size ms_slide_inference, all 18 windows refined: {time.time() - t0:.3f} s (includes first-use allocations)")
    assert info1["n_refined"] == 18 and info1["refined"].shape == (18, 19, 32, 32)
    from vfmseg_b200.engine import slide_boxes
    bx = torch.tensor(slide_boxes(1024, 2048, (512, 512), (320, 320)), dtype=torch.int32).cuda()
    lab_ref, log_ref = ops.slide_merge_argmax(info1["refined"], bx, 1, (512, 512), (1024, 2048), want_logits=True)
    assert torch.equal(logits1, log_ref) and torch.equal(labels1, lab_ref)
    assert torch.isfinite(logits1).all()
    # the shipped gate constants (0.968 / 0.8): random weights are never that confident -> everything is refined
    model.test_cfg.conf = 0.8
    assert model._ms(img, False, "image")[2]["n_refined"] == 18
