"""GPU parity of the SAM ViT backbone path (BASELINE config 5): rel-pos terms, row gather and the rel-pos attention kernel
against torch, and the registered EncoderDecoder(LoRABackbone(SAMViT), LinearHead) against the golden vectors produced by
the reference's own modules (tests/golden/tiny_sam.npz) and the oracle restatement."""
from pathlib import Path

import numpy as np
import pytest
import torch

from test_e2e_gpu import _check_labels, _check_logits

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden"
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def _rand(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


@pytest.mark.parametrize("q_h,q_w,d", [(14, 14, 80), (32, 32, 80), (5, 9, 64)])
def test_relpos_terms(q_h, q_w, d):
    from vfmseg_b200 import ops
    n, heads = 3, 2
    qkv = _rand(n * q_h * q_w, 3 * heads * d, seed=1, dtype=torch.bfloat16)
    Rh, Rw = _rand(q_h, q_h, d, seed=2), _rand(q_w, q_w, d, seed=3)
    got = ops.relpos_terms(qkv, Rh, Rw, n, heads, d).cpu()
    q = qkv.float().cpu().view(n, q_h, q_w, 3, heads, d)[:, :, :, 0].permute(0, 3, 1, 2, 4)       # [n, heads, q_h, q_w, d]
    rel_h = torch.einsum("bnhwc,hkc->bnhwk", q, Rh.cpu())
    rel_w = torch.einsum("bnhwc,wkc->bnhwk", q, Rw.cpu())
    ref = torch.cat((rel_h, rel_w), dim=-1).reshape(n, heads, q_h * q_w, q_h + q_w)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-3)


def test_rows_gather_window_partition():
    """The engine's row maps against the reference's own window_partition / window_unpartition arithmetic
    (sam_vit.py:292-346), restated with torch ops."""
    import torch.nn.functional as F
    from vfmseg_b200 import ops
    from vfmseg_b200.sam_engine import PackedSam
    n, gh, gw, ws, C = 2, 16, 16, 14, 64
    x = _rand(n * gh * gw, C, seed=4, dtype=torch.bfloat16)
    part, unpart, n_win = PackedSam._window_maps(type("E", (), {"_maps": {}, "device": "cuda"})(), n, gh, gw, ws)
    xs = x.float().cpu().view(n, gh, gw, C)
    ph, pw = (ws - gh % ws) % ws, (ws - gw % ws) % ws
    p = F.pad(xs, (0, 0, 0, pw, 0, ph))
    Hp, Wp = gh + ph, gw + pw
    win = p.view(n, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, C)
    got = ops.rows_gather(x, part)
    assert n_win == 8 and torch.equal(got.float().cpu(), win)
    back = ops.rows_gather(got, unpart)
    assert torch.equal(back, x)


@pytest.mark.parametrize("n_seq,k_h,k_w,heads,d,bias", [(5, 14, 14, 2, 80, True), (2, 32, 32, 2, 80, True), (3, 20, 20, 1, 80, True),
                                                        (2, 16, 16, 2, 80, False), (2, 14, 14, 2, 64, True), (1, 64, 64, 1, 80, True)])
def test_attention_relpos(n_seq, k_h, k_w, heads, d, bias):
    from vfmseg_b200 import ops
    S = k_h * k_w
    qkv = _rand(n_seq * S, 3 * heads * d, scale=0.7, seed=5, dtype=torch.bfloat16)
    rel = _rand(n_seq, heads, S, k_h + k_w, scale=1.5, seed=6) if bias else None
    scale = d ** -0.5
    got = ops.attention_relpos(qkv, rel, n_seq, S, heads, d, k_h, k_w, scale).float().cpu()
    q, k, v = qkv.float().cpu().view(n_seq, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    att = (q * scale) @ k.transpose(-1, -2)
    if bias:
        r = rel.cpu()
        att = (att.view(n_seq, heads, S, k_h, k_w) + r[..., :k_h, None] + r[..., None, k_h:]).view(n_seq, heads, S, S)
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * S, heads * d)
    err = (got - ref).abs()
    assert (err <= 2e-2 + 2e-2 * ref.abs()).all(), err.max()


def _build_sam(cfg, seed=0):
    import vfmseg_b200
    from vfmseg_b200 import synthetic
    sd = synthetic.synthetic_sam_state_dict(cfg, seed=seed)
    model = vfmseg_b200.MODELS.build(dict(cfg))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m], (missing, unexpected)
    return model.cuda().eval(), sd


def test_tiny_sam_vs_reference_golden():
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "tiny_sam.npz")
    cfg = synthetic.tiny_sam_config()
    model, _ = _build_sam(cfg)
    img = synthetic.synthetic_images(1, 256, 384, seed=1234)
    # backbone contract first: four [B, C, h, w] maps of the raw block outputs
    from oracle import torch_ref
    x = torch_ref.preprocess(img, MEAN, STD, True)[:, :, :256, :256].contiguous()
    feats = model.extract_feat(x.cuda())
    assert len(feats) == 4 and all(f.shape == (1, 640, 16, 16) for f in feats)
    for i, f in enumerate(feats):
        _check_logits(f, torch.from_numpy(g["feats"][i].astype(np.float32)), f"SAM ViT tap {i} vs reference golden")
    labels, logits = model.predict_labels(img.cuda(), want_logits=True)
    ref = torch.from_numpy(g["logits"].astype(np.float32))
    _check_logits(logits, ref, "tiny SAM ViT slide vs reference golden")
    _check_labels(labels, ref, "tiny SAM ViT labels vs reference golden", raw_min=0.985, top2_min=0.999)
    # the fixed pos_embed only accepts the grid the model was built for (sam_vit.py:131-132)
    from vfmseg_b200 import _C
    with pytest.raises(_C.VfmError):
        model.extract_feat(torch.zeros(1, 3, 256, 320).cuda())


def test_sam_real_dimension_crop_vs_reference_golden():
    """BASELINE config 5 at REAL dimensions: SAM ViT-H/16 (1280 wide, 32 blocks, 16 heads x 80; 28 windowed blocks over 25
    padded 14 x 14 windows, 4 global blocks over 4096 tokens with decomposed rel-pos bias) + LinearHead, one 1024 x 1024
    window; golden = the reference's own sam_vit.py + linear_head.py in fp32 (oracle/make_golden.py:sam_crop)."""
    from vfmseg_b200 import synthetic
    g = np.load(GOLDEN / "sam_crop.npz")
    cfg = synthetic.sam_model_config(img_size=1024, crop_size=(1024, 1024), stride=(682, 682))
    model, _ = _build_sam(cfg)
    img = synthetic.synthetic_images(1, 1024, 1024, seed=1234).cuda()
    low = model.engine().crops_lowres(img, torch.tensor([[0, 0, 0, 0]], dtype=torch.int32, device="cuda"), (1024, 1024))
    assert low.shape == (1, 19, 256, 256)
    ref = torch.from_numpy(g["lowres_sub"].astype(np.float32))
    # 32 blocks without LayerScale: the reference's own modules under CPU bf16 autocast keep 96.8 % of these logits inside the
    # band (rel rms 1.36 %, stored with the golden); asserted: at least that, and >= 99 % (measured 99.4 %).
    _check_logits(low[0, :, ::2, ::2], ref, "SAM ViT-H 1024 crop low-res logits vs reference golden", frac=max(0.99, float(g["autocast_within"])))
    got = low[0, :, ::2, ::2].float().cpu()
    assert ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item() <= float(g["autocast_rel_rms"])
    _check_labels(low[0].argmax(0)[::2, ::2], ref, "SAM ViT-H 1024 crop low-res labels vs reference golden", raw_min=float(g["autocast_label_agreement"]), top2_min=0.999)


def test_full_size_sam_runs():
    """BASELINE config 5 shapes at the shipped crop (SAM ViT-H/16: 1280 wide, 32 blocks, 16 heads x 80, 1024x2048 image,
    crop 512 / stride 320 as configs/_base_/models/lora_sam_linear.py:49-54): finite logits, batching invariance."""
    from vfmseg_b200 import synthetic
    cfg = synthetic.sam_model_config()
    model, _ = _build_sam(cfg)
    img = synthetic.synthetic_images(2, 1024, 2048, seed=21).cuda()
    labels, logits = model.predict_labels(img[:1], want_logits=True)
    assert torch.isfinite(logits).all() and labels.shape == (1, 1024, 2048)
    labels2, _ = model.predict_labels(img)
    assert torch.equal(labels2[0], labels[0])


@pytest.mark.parametrize("n_seq,k,heads,bias", [(5, 14, 2, True), (162, 14, 16, True), (3, 14, 2, False), (4, 12, 3, True), (2, 8, 1, True)])
def test_attention_window_tc(n_seq, k, heads, bias):
    """tcgen05 window attention (bias added by the tensor core through one-hot key columns) against torch fp32 on the same
    packed rows: q | k | v | G_h | G_w with random table terms, rel_h[q, kh] = G_h[qh - kh + k - 1]."""
    from vfmseg_b200 import ops
    d, S, L = 80, k * k, 2 * k - 1
    C = heads * d
    ld = 3 * C + 2 * heads * L
    ld += (-ld) % 32
    qkv = _rand(n_seq * S, ld, scale=0.7, seed=7, dtype=torch.bfloat16)
    scale = d ** -0.5
    got = ops.attention_window_tc(qkv, n_seq, S, heads, d, k, k, scale, 3 * C if bias else -1).float().cpu()
    x = qkv.float().cpu()
    q, kk_, v = x[:, :3 * C].view(n_seq, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    att = (q * scale) @ kk_.transpose(-1, -2)
    if bias:
        Gh = x[:, 3 * C:3 * C + heads * L].view(n_seq, k, k, heads, L)
        Gw = x[:, 3 * C + heads * L:3 * C + 2 * heads * L].view(n_seq, k, k, heads, L)
        idx = torch.arange(k)[:, None] - torch.arange(k)[None, :] + k - 1                      # [q, key]
        rel_h = torch.gather(Gh, 4, idx[None, :, None, None, :].expand(n_seq, k, k, heads, k))  # [n, qh, qw, head, kh]
        rel_w = torch.gather(Gw, 4, idx[None, None, :, None, :].expand(n_seq, k, k, heads, k))  # [n, qh, qw, head, kw]
        b = rel_h[..., :, None] + rel_w[..., None, :]                                          # [n, qh, qw, head, kh, kw]
        att = att + b.permute(0, 3, 1, 2, 4, 5).reshape(n_seq, heads, S, S)
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * S, C)
    err = (got - ref).abs()
    assert (err <= 2e-2 + 2e-2 * ref.abs()).all(), err.max()


def _relpos_ref(qkv, n_seq, k_h, k_w, heads, d, scale):
    """torch fp32 attention on packed rows q | k | v | G_h | G_w with rel_h[q, kh] = G_h[qh - kh + k_h - 1] (and likewise w)."""
    S, Lh, Lw = k_h * k_w, 2 * k_h - 1, 2 * k_w - 1
    C = heads * d
    x = qkv.float().cpu()
    q, kk_, v = x[:, :3 * C].view(n_seq, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    att = (q * scale) @ kk_.transpose(-1, -2)
    Gh = x[:, 3 * C:3 * C + heads * Lh].view(n_seq, k_h, k_w, heads, Lh)
    Gw = x[:, 3 * C + heads * Lh:3 * C + heads * (Lh + Lw)].view(n_seq, k_h, k_w, heads, Lw)
    ih = torch.arange(k_h)[:, None] - torch.arange(k_h)[None, :] + k_h - 1
    iw = torch.arange(k_w)[:, None] - torch.arange(k_w)[None, :] + k_w - 1
    rel_h = torch.gather(Gh, 4, ih[None, :, None, None, :].expand(n_seq, k_h, k_w, heads, k_h))
    rel_w = torch.gather(Gw, 4, iw[None, None, :, None, :].expand(n_seq, k_h, k_w, heads, k_w))
    b = rel_h[..., :, None] + rel_w[..., None, :]
    att = att + b.permute(0, 3, 1, 2, 4, 5).reshape(n_seq, heads, S, S)
    return (att.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * S, C)


@pytest.mark.parametrize("n_seq,k_h,k_w,heads", [(2, 32, 32, 2), (1, 64, 64, 1), (3, 20, 20, 2), (2, 16, 16, 3), (1, 24, 40, 1)])
def test_attention_global_tc(n_seq, k_h, k_w, heads):
    """tcgen05 grid attention (key-tile loop, online softmax, bias through one-hot key columns by TMA) against torch fp32:
    full tiles, a ragged last tile (400, 960 keys), one and two bias atoms, non-square grid."""
    from vfmseg_b200 import ops
    d, S = 80, k_h * k_w
    C = heads * d
    ld = 3 * C + heads * (2 * k_h - 1 + 2 * k_w - 1)
    ld += (-ld) % 32
    qkv = _rand(n_seq * S, ld, scale=0.7, seed=8, dtype=torch.bfloat16)
    scale = d ** -0.5
    e = ops.relpos_onehot(k_h, k_w, "cuda")
    got = ops.attention_global_tc(qkv, e, n_seq, S, heads, d, k_h, k_w, scale, 3 * C).float().cpu()
    ref = _relpos_ref(qkv, n_seq, k_h, k_w, heads, d, scale)
    err = (got - ref).abs()
    assert (err <= 2e-2 + 2e-2 * ref.abs()).all(), err.max()
    # and against the mma.sync kernel on the same rows
    if k_h % 2 == 0 and k_w % 2 == 0:
        other = ops.attention_relpos_terms(qkv, n_seq, S, heads, d, k_h, k_w, scale, 3 * C).float().cpu()
        assert ((got - other).abs() <= 2e-2 + 2e-2 * other.abs()).all()


def test_sam_1024_crop_two_blocks_vs_oracle():
    """BASELINE config 5's crop (1024^2 -> 64 x 64 tokens) at ViT-H width with two blocks (one windowed: 25 windows, 64 -> 70
    zero padding; one global: 4096 tokens, two 64-column bias atoms, rel-pos tables interpolated 255 -> 127) against the oracle
    restatement on the CPU."""
    import vfmseg_b200
    from oracle import torch_ref
    from vfmseg_b200 import synthetic
    cfg = synthetic.sam_model_config(img_size=1024, depth=2, global_attn_indexes=(1,), out_indices=(0, 1), crop_size=(1024, 1024),
                                     stride=(682, 682))
    sd = synthetic.synthetic_sam_state_dict(cfg, seed=2)
    model = vfmseg_b200.MODELS.build(dict(cfg))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "num_batches_tracked" not in m]
    model = model.cuda().eval()
    img = synthetic.synthetic_images(1, 1024, 1024, seed=5)
    x = torch_ref.preprocess(img, MEAN, STD, True)
    feats = model.extract_feat(x.cuda())
    pre = "backbone.model.base_model.model."
    bb = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    lc = cfg["backbone"]["Lora_config"]
    with torch.no_grad():
        ref = torch_ref.sam_forward(x, bb, depth=2, num_heads=16, window_size=14, global_attn_indexes=(1,), out_indices=(0, 1),
                                    lora_scale=lc["lora_alpha"] / lc["r"])
    for i, (f, r) in enumerate(zip(feats, ref)):
        assert f.shape == (1, 1280, 64, 64)
        _check_logits(f, r, f"SAM ViT-H 1024 crop, block {i} output vs oracle")


@pytest.mark.parametrize("variant,fold", [("tiny80", "3"), ("tiny64", "3"), ("tiny64", "0"), ("real", "3"), ("real", "0")])
def test_sam_c_driver_equals_python_driver(monkeypatch, variant, fold):
    """vfm_sam_forward (one C call for SAMViT.forward, sam_vit.py:123-147) against the same launch sequence issued operator by
    operator from Python (VFM_SAM_DRIVER=py): bit-identical taps. tiny80 = head_dim 80 (tcgen05 window / global attention, width
    640: no LayerNorm folding), tiny64 = head_dim 64 (CUDA-core rel-pos attention + row gather, width 512: folding on / off),
    real = SAM ViT-H dimensions (1280 wide, 32 blocks, 32 x 32 tokens, 14 x 14 windows)."""
    from vfmseg_b200 import synthetic
    if variant == "real":
        cfg, crop = synthetic.sam_model_config(), 512
    elif variant == "tiny64":
        cfg, crop = synthetic.tiny_sam_config(embed_dim=512, num_heads=8), 256
    else:
        cfg, crop = synthetic.tiny_sam_config(), 256
    model, _ = _build_sam(cfg)
    eng = model.engine()
    g = crop // 16
    img = synthetic.synthetic_images(2, crop + 32, crop + 48, seed=78).cuda()
    crops = torch.tensor([[0, 0, 0, 0], [1, 32, 48, 0], [0, 16, 8, 0]], dtype=torch.int32, device="cuda")
    monkeypatch.setenv("VFM_LN_FOLD", fold)
    for x, cr in ((img, crops), ((img.float() - 110.0) / 60.0, crops[:2].contiguous())):
        monkeypatch.setenv("VFM_SAM_DRIVER", "py")
        ref = eng.backbone_taps(x.contiguous(), cr, g, g)
        monkeypatch.setenv("VFM_SAM_DRIVER", "c")
        got = eng.backbone_taps(x.contiguous(), cr, g, g)
        assert got.shape == ref.shape and torch.isfinite(got.float()).all()
        assert torch.equal(got, ref)
