"""GPU parity of the EVA02 backbone path (BASELINE config 4): RoPE / SwiGLU+LayerNorm kernels against torch, and the
registered EncoderDecoder(LoRABackbone(EVA2), LinearHead) against the golden vectors produced by the reference's own
modules (tests/golden/tiny_eva.npz) and the oracle restatement."""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_e2e_gpu import _check_labels, _check_logits

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _rand(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def test_rope_qk():
    from oracle import torch_ref
    from vfmseg_b200 import ops
    n, grid, heads = 2, 4, 3
    T, C = grid * grid + 1, heads * 64
    qkv = _rand(n * T, 3 * C, seed=1, dtype=torch.bfloat16)
    cos, sin = torch_ref.eva_rope_tables(64, 16, grid)
    got = ops.rope_qk_(qkv.clone(), heads, T, cos.cuda().contiguous(), sin.cuda().contiguous()).float().cpu()
    x = qkv.float().cpu().view(n, T, 3, heads, 64)
    ref = x.clone()
    for w in (0, 1):
        t = x[:, 1:, w]                                            # [n, P, heads, 64]
        ref[:, 1:, w] = t * cos[None, :, None, :] + torch_ref._rotate_half(t) * sin[None, :, None, :]
    ref = ref.view(n * T, 3 * C)
    assert ((got - ref).abs() <= 1e-2 + 1e-2 * ref.abs()).all()
    assert torch.equal(got.view(n, T, 3, heads, 64)[:, 0], x[:, 0])           # cls rows untouched
    assert torch.equal(got.view(n, T, 3, heads, 64)[:, :, 2], x[:, :, 2])     # v untouched


@pytest.mark.parametrize("H", [682, 2730, 96])
def test_swiglu_layernorm(H):
    from vfmseg_b200 import ops
    M, Hp = 70, (H + 7) // 8 * 8
    x12 = _rand(M, 2 * Hp, seed=2, dtype=torch.bfloat16)
    g, b = _rand(H, seed=3), _rand(H, seed=4)
    got = ops.swiglu_layernorm(x12, g, b, H, 1e-5).float()
    x1, x2 = x12[:, :H].float(), x12[:, Hp:Hp + H].float()
    ref = F.layer_norm(F.silu(x1) * x2, (H,), g, b, 1e-5)
    assert ((got[:, :H] - ref).abs() <= 1e-2 + 1e-2 * ref.abs()).all()
    assert (got[:, H:] == 0).all()


def _build_eva(cfg, seed=0):
    import vfmseg_b200
    from vfmseg_b200 import synthetic
    sd = synthetic.synthetic_eva_state_dict(cfg, seed=seed)
    model = vfmseg_b200.MODELS.build(dict(cfg))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m], (missing, unexpected)
    return model.cuda().eval(), sd


def test_tiny_eva_vs_reference_golden():
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "tiny_eva.npz")
    cfg = synthetic.tiny_eva_config()
    model, _ = _build_eva(cfg)
    img = synthetic.synthetic_images(1, 80, 112, seed=1234)
    labels, logits = model.predict_labels(img.cuda(), want_logits=True)
    ref = torch.from_numpy(g["logits"].astype(np.float32))
    _check_logits(logits, ref, "tiny EVA02 slide vs reference golden")
    _check_labels(labels, ref, "tiny EVA02 labels vs reference golden", raw_min=0.985, top2_min=0.999)
    # backbone contract: four [B, C, h, w] maps of the un-normalised residual stream
    from oracle import torch_ref
    x = torch_ref.preprocess(img, MEAN, STD, True)[:, :, :64, :64].contiguous()
    feats = model.extract_feat(x.cuda())
    assert len(feats) == 4 and all(f.shape == (1, 256, 4, 4) for f in feats)
    for i, f in enumerate(feats):
        _check_logits(f, torch.from_numpy(g["feats"][i]), f"EVA02 tap {i} vs reference golden")
    # the fixed-grid pos_embed / RoPE only accept the grid the model was built for (eva_02.py:825-826)
    from vfmseg_b200 import _C
    with pytest.raises(_C.VfmError):
        model.extract_feat(torch.zeros(1, 3, 64, 96).cuda())


def test_full_size_eva_runs():
    """BASELINE config 4 shapes (EVA02-L/16, 1024x2048, crop 512 / stride 320): finite logits, batching invariance."""
    from vfmseg_b200 import synthetic
    cfg = synthetic.eva_model_config()
    model, _ = _build_eva(cfg)
    img = synthetic.synthetic_images(2, 1024, 2048, seed=21).cuda()
    labels, logits = model.predict_labels(img[:1], want_logits=True)
    assert torch.isfinite(logits).all() and labels.shape == (1, 1024, 2048)
    labels2, _ = model.predict_labels(img)
    assert torch.equal(labels2[0], labels[0])


def test_eva_real_dimension_crop_vs_reference_golden():
    """BASELINE config 4 at REAL dimensions: EVA02-L/16 (1024 wide, 24 blocks, RoPE, SwiGLU + sub-LN) + LinearHead, one
    512 x 512 window; golden = the reference's own eva_02.py + linear_head.py in fp32 (oracle/make_golden.py:eva_crop)."""
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "eva_crop.npz")
    cfg = synthetic.eva_model_config()
    model, _ = _build_eva(cfg)
    img = synthetic.synthetic_images(1, 512, 512, seed=1234).cuda()
    low = model.engine().crops_lowres(img, torch.tensor([[0, 0, 0, 0]], dtype=torch.int32, device="cuda"), (512, 512))
    assert low.shape == (1, 19, 128, 128)
    ref = torch.from_numpy(g["lowres_sub"].astype(np.float32))
    # EVA02 has no LayerScale (init_values=None): every branch enters the residual stream at full weight and the bf16 error
    # of 24 blocks adds up to ~1 % of the logit rms at real depth (the tiny 4-block golden sits at 99.98 % inside the band).
    # Yardstick stored with the golden: the reference's own modules under CPU bf16 autocast keep 95.8 % inside the band
    # (rel rms 1.47 %); asserted: at least that, and >= 97.5 % (measured 98.2 %).
    _check_logits(low[0, :, ::2, ::2], ref, "EVA02-L crop low-res logits vs reference golden", frac=max(0.975, float(g["autocast_within"])))
    got = low[0, :, ::2, ::2].float().cpu()
    assert ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item() <= float(g["autocast_rel_rms"])
    _check_labels(low[0].argmax(0)[::2, ::2], ref, "EVA02-L crop low-res labels vs reference golden", raw_min=float(g["autocast_label_agreement"]), top2_min=0.999)


def test_gemm_bias_rope_matches_gemm_then_rope():
    """RoPE in the qkv GEMM epilogue (fp32 accumulator, one rounding) against the GEMM followed by the in-place RoPE kernel
    (two roundings) and against torch fp32."""
    from oracle import torch_ref
    from vfmseg_b200 import ops
    n, grid, heads = 3, 4, 4
    T, C = grid * grid + 1, heads * 64
    a = _rand(n * T, C, seed=11, dtype=torch.bfloat16)
    w = _rand(3 * C, C, scale=0.05, seed=12, dtype=torch.bfloat16)
    b = _rand(3 * C, scale=0.1, seed=13)
    cos, sin = torch_ref.eva_rope_tables(64, 16, grid)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    got = ops.gemm_bias_rope_bf16(a, w, b, cos, sin, 2 * C, T).float().cpu()
    two = ops.rope_qk_(ops.gemm_bias_bf16(a, w, b), heads, T, cos, sin).float().cpu()
    y = (a.float().cpu() @ w.float().cpu().t() + b.cpu()).view(n, T, 3, heads, 64)
    ref = y.clone()
    for which in (0, 1):
        t = y[:, 1:, which]
        ref[:, 1:, which] = t * cos.cpu()[None, :, None, :] + torch_ref._rotate_half(t) * sin.cpu()[None, :, None, :]
    ref = ref.view(n * T, 3 * C)
    assert ((got - ref).abs() <= 1e-2 + 1e-2 * ref.abs()).all()
    assert ((got - two).abs() <= 2e-2 + 2e-2 * two.abs()).all()


@pytest.mark.parametrize("real_dims", [False, True])
@pytest.mark.parametrize("fold", ["3", "0", "1", "2"])
def test_eva_c_driver_equals_python_driver(monkeypatch, real_dims, fold):
    """vfm_eva_forward (one C call for EVA2.forward_features, eva_02.py:816-849) against the same launch sequence issued operator
    by operator from Python (VFM_EVA_DRIVER=py): bit-identical taps, with the LayerNorm folding on, off and half on (VFM_LN_FOLD),
    uint8 and pre-normalised fp32 images, several windows of two images."""
    from vfmseg_b200 import synthetic
    if real_dims and fold in ("1", "2"):
        pytest.skip("half-folded sequences are covered at the tiny size")
    cfg = synthetic.eva_model_config() if real_dims else synthetic.tiny_eva_config()
    model, _ = _build_eva(cfg)
    eng = model.engine()
    crop = 512 if real_dims else 64
    g = crop // 16
    img = synthetic.synthetic_images(2, crop + 32, crop + 48, seed=77).cuda()
    crops = torch.tensor([[0, 0, 0, 0], [1, 32, 48, 0], [0, 16, 8, 0]], dtype=torch.int32, device="cuda")
    monkeypatch.setenv("VFM_LN_FOLD", fold)
    for x in (img, (img.float() - 110.0) / 60.0):
        monkeypatch.setenv("VFM_EVA_DRIVER", "py")
        ref = eng.backbone_taps(x.contiguous(), crops, g, g)
        monkeypatch.setenv("VFM_EVA_DRIVER", "c")
        got = eng.backbone_taps(x.contiguous(), crops, g, g)
        assert got.shape == ref.shape and torch.isfinite(got.float()).all()
        assert torch.equal(got, ref)
