"""Per-operator parity of the CUDA kernels (through the C ABI) against plain torch fp32.

Tolerances: operands are bf16-rounded before both paths, accumulation is fp32 on both sides, so
GEMM-only outputs agree to ~1e-3 relative; bf16 outputs add one rounding (2^-9 relative).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vfmseg_b200 import _C, ops
    _C.check(_C.load().vfm_device_check())
    return ops


def _rand(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _close(a, b, rtol, atol, what):
    a = a.float()
    b = b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} out of tolerance, max abs err {err.max().item():.4g}, ref max {b.abs().max().item():.4g}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (1025, 1024, 1024), (18450, 3072, 1024),
                                   (300, 384, 128), (2050, 1024, 4096), (77, 96, 72)])
def test_gemm_f32(ops, M, N, K):
    a = _rand(M, K, seed=1, dtype=torch.bfloat16)
    w = _rand(N, K, scale=K ** -0.5, seed=2, dtype=torch.bfloat16)
    b = _rand(N, seed=3)
    out = ops.gemm_f32(a, w, b)
    ref = a.float() @ w.float().t() + b
    _close(out, ref, 2e-3, 2e-3, f"gemm_f32 {M}x{N}x{K}")


def test_gemm_bias_bf16_and_gelu(ops):
    M, N, K = 2050, 4096, 1024
    a = _rand(M, K, seed=4, dtype=torch.bfloat16)
    w = _rand(N, K, scale=K ** -0.5, seed=5, dtype=torch.bfloat16)
    b = _rand(N, seed=6)
    ref = a.float() @ w.float().t() + b
    _close(ops.gemm_bias_bf16(a, w, b), ref, 1e-2, 1e-2, "gemm_bias_bf16")
    _close(ops.gemm_bias_bf16(a, w, None), ref - b, 1e-2, 1e-2, "gemm_bias_bf16(no bias)")
    _close(ops.gemm_bias_gelu_bf16(a, w, b), F.gelu(ref), 1e-2, 1e-2, "gemm_bias_gelu_bf16")


def test_gemm_residual_with_tap(ops):
    n_crops, T, C, K = 3, 65, 256, 512
    M = n_crops * T
    a = _rand(M, K, seed=7, dtype=torch.bfloat16)
    w = _rand(C, K, scale=K ** -0.5, seed=8, dtype=torch.bfloat16)
    b = _rand(C, seed=9)
    g = _rand(C, seed=10)
    x0 = _rand(M, C, seed=11)
    ref = x0 + g * (a.float() @ w.float().t() + b)
    x = x0.clone()
    tap = torch.zeros(n_crops * (T - 1), 2 * C, device="cuda", dtype=torch.bfloat16)
    ops.gemm_bias_ls_residual_(x, a, w, b, g, tap=tap, tap_col0=C, tokens_per_crop=T)
    _close(x, ref, 2e-3, 2e-3, "residual")
    ref_tap = ref.view(n_crops, T, C)[:, 1:].reshape(-1, C)
    _close(tap[:, C:], ref_tap, 1e-2, 1e-2, "tap")
    assert tap[:, :C].abs().max().item() == 0
    x2 = x0.clone()
    ops.gemm_bias_ls_residual_(x2, a, w, b, g)
    assert torch.equal(x, x2)


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (2050, 1024, 1024), (1025, 1024, 4096), (36900, 1024, 1024), (37, 512, 64)])
def test_gemm_residual_stats(ops, M, N, K):
    """Residual GEMM that also emits bf16(x) and the LayerNorm statistics: x bit-identical to the reduce-add epilogue."""
    a = _rand(M, K, seed=41, dtype=torch.bfloat16)
    w = _rand(N, K, scale=K ** -0.5, seed=42, dtype=torch.bfloat16)
    b = _rand(N, seed=43)
    g = _rand(N, seed=44) * 0.3
    x0 = _rand(M, N, seed=45) * 2 + 0.25
    x_ref = ops.gemm_bias_ls_residual_(x0.clone(), a, w, b, g)
    x = x0.clone()
    xb, stats = ops.gemm_bias_ls_residual_stats_(x, a, w, b, g)
    assert torch.equal(x, x_ref), f"x differs from the reduce-add path: max {(x - x_ref).abs().max().item():.3g}"
    assert torch.equal(xb, x.to(torch.bfloat16)), "xb != bf16(x)"
    xs = x.double().view(M, N // 128, 128)
    _close(stats[..., 0], xs.sum(-1).float(), 1e-5, 1e-3, "slot sums")
    _close(stats[..., 1], (xs * xs).sum(-1).float(), 1e-5, 1e-3, "slot sums of squares")
    _close(x, x0 + g * (a.float() @ w.float().t() + b), 2e-3, 2e-3, "residual value")


@pytest.mark.parametrize("M,C,N,gelu", [(300, 256, 768, False), (2050, 1024, 3072, False), (2050, 1024, 4096, True), (36900, 1024, 256, True)])
def test_gemm_lnfold(ops, M, C, N, gelu):
    """Linear(LayerNorm(x)) with the norm folded into the weights and applied by the epilogue from the emitted statistics,
    against torch fp32 LayerNorm -> Linear (-> GELU) and against the unfolded kernel pair."""
    K0 = 512
    a = _rand(M, K0, seed=51, dtype=torch.bfloat16)
    w0 = _rand(C, K0, scale=K0 ** -0.5, seed=52, dtype=torch.bfloat16)
    x = _rand(M, C, seed=53) * 1.5 + 0.3
    x[:, 7] += 6.0   # an outlier channel, as ViT residual streams have
    xb, stats = ops.gemm_bias_ls_residual_stats_(x, a, w0, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda") * 0.5)
    ln_w = 1 + 0.2 * _rand(C, seed=54)
    ln_b = 0.1 * _rand(C, seed=55)
    w = _rand(N, C, scale=C ** -0.5, seed=56)
    b = _rand(N, seed=57) * 0.1
    wf, bf, cs = ops.fold_layernorm(w, b, ln_w, ln_b)
    got = ops.gemm_lnfold_bf16(xb, stats, wf, bf, cs, 1e-6, gelu=gelu).float()
    ref = F.linear(F.layer_norm(x, (C,), ln_w, ln_b, 1e-6), w, b)
    if gelu:
        ref = F.gelu(ref)
    xn = ops.layernorm(x, ln_w, ln_b, 1e-6)
    wb = w.to(torch.bfloat16)
    old = (ops.gemm_bias_gelu_bf16(xn, wb, b) if gelu else ops.gemm_bias_bf16(xn, wb, b)).float()
    rms = ref.pow(2).mean().sqrt().item()
    e_new = (got - ref).pow(2).mean().sqrt().item() / rms
    e_old = (old - ref).pow(2).mean().sqrt().item() / rms
    assert e_new < 6e-3, f"folded LayerNorm GEMM: rms error {e_new:.3g} of the output rms"
    assert e_new < 1.5 * e_old + 1e-4, f"folded path ({e_new:.3g}) is less accurate than LayerNorm kernel + GEMM ({e_old:.3g})"
    _close(got, ref, 3e-2, 3e-2 * rms, "gemm_lnfold_bf16")


def test_patch_gather_and_embed(ops):
    B, H, W, C = 2, 96, 128, 256
    img = _rand(B, 3, H, W, seed=12)
    crops = torch.tensor([[0, 0, 0, 0], [0, 32, 61, 0], [1, 13, 64, 0]], dtype=torch.int32).cuda()
    gh = gw = 4
    a = ops.patch_gather(img, crops, gh, gw)
    ref_rows = []
    for (b_, y1, x1, _) in crops.tolist():
        crop = img[b_, :, y1:y1 + 64, x1:x1 + 64]
        ref_rows.append(F.unfold(crop[None], kernel_size=16, stride=16)[0].t())  # [P, 3*256] (c, py, px)
    ref_a = torch.cat(ref_rows).to(torch.bfloat16)
    assert torch.equal(a, ref_a), "patch_gather fp32"
    # uint8 BGR + normalisation
    g = torch.Generator().manual_seed(13)
    img8 = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).cuda()
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    a8 = ops.patch_gather(img8, crops, gh, gw, ops.pixel_norm(mean, std, True))
    imgf = (img8[:, [2, 1, 0]].float() - torch.tensor(mean).view(1, 3, 1, 1).cuda()) / torch.tensor(std).view(1, 3, 1, 1).cuda()
    ref8 = torch.cat([F.unfold(imgf[b_, :, y1:y1 + 64, x1:x1 + 64][None], 16, stride=16)[0].t() for (b_, y1, x1, _) in crops.tolist()])
    _close(a8, ref8, 1e-2, 1e-2, "patch_gather uint8")
    # embed GEMM + pos + cls
    w = _rand(C, 768, scale=768 ** -0.5, seed=14, dtype=torch.bfloat16)
    bias = _rand(C, seed=15)
    pos = _rand(gh * gw + 1, C, seed=16)
    cls = _rand(C, seed=17)
    x = ops.gemm_patch_embed(a, w, bias, pos, crops.shape[0], gh * gw)
    ops.cls_rows_(x, cls, pos, crops.shape[0], gh * gw + 1)
    tok = (ref_a.float() @ w.float().t() + bias).view(3, gh * gw, C)
    ref_x = torch.cat([cls.expand(3, 1, C), tok], 1) + pos
    _close(x, ref_x.reshape(-1, C), 2e-3, 2e-3, "patch_embed")


@pytest.mark.parametrize("M,C", [(1025, 1024), (37, 128), (18450, 1024), (100, 384)])
def test_layernorm(ops, M, C):
    x = _rand(M, C, seed=18) * 3 + 0.5
    g = _rand(C, seed=19)
    b = _rand(C, seed=20)
    out = ops.layernorm(x, g, b, 1e-6)
    _close(out, F.layer_norm(x, (C,), g, b, 1e-6), 1e-2, 1e-2, "layernorm")


def _attn_ref(qkv, n_seq, S, heads):
    q, k, v = qkv.float().view(n_seq, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    return ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(n_seq * S, heads * 64)


# mode 0 = automatic, 1 = tensor tiles over every token (ragged tail tiles), 2 = token 0 split off (cls token)
@pytest.mark.parametrize("n_seq,S,heads,mode", [
    (1, 128, 1, 1), (2, 256, 2, 1), (1, 1025, 16, 0), (3, 1025, 4, 1), (3, 1025, 4, 2), (2, 197, 2, 1), (2, 197, 2, 2),
    (1, 2049, 2, 0), (1, 2049, 2, 1), (2, 17, 2, 0), (2, 17, 2, 2), (2, 2, 1, 2), (1, 385, 3, 0), (2, 641, 2, 1),
    # more units than persistent CTAs (2 x 148): every CTA streams several units back to back — odd and even key-tile
    # counts per unit (buffer / phase parity across unit boundaries), one- and two-tile units, extra-token mode
    (20, 257, 12, 0), (20, 257, 12, 2), (40, 65, 8, 1), (600, 17, 1, 0), (300, 129, 2, 2), (7, 1025, 16, 0),
    # mode 3: the serial four-CTAs-per-SM kernel
    (1, 128, 1, 3), (3, 1025, 4, 3), (2, 197, 2, 3), (1, 2049, 2, 3), (2, 17, 2, 3), (20, 257, 12, 3),
    # modes 4 / 5: the ping-pong kernel (256-query units, 128-key tiles, one CTA per SM; 5 = extra-token split, what mode 0
    # picks for ViT windows). Ragged query / key tails, several units per CTA, ONE key tile per unit (40 x 65: a softmax warp
    # can run two units ahead of the output warps — the barrier-parity case that dead-locked during development), 36 windows
    (1, 128, 1, 4), (2, 256, 2, 4), (1, 1024, 2, 4), (3, 1025, 4, 4), (3, 1025, 4, 5), (2, 197, 2, 4), (2, 197, 2, 5),
    (1, 2049, 2, 4), (1, 2049, 2, 5), (2, 17, 2, 4), (2, 17, 2, 5), (2, 2, 1, 5), (1, 385, 3, 5), (2, 641, 2, 4),
    (20, 257, 12, 4), (20, 257, 12, 5), (40, 65, 8, 4), (40, 65, 8, 5), (600, 17, 1, 5), (300, 129, 2, 4), (300, 129, 2, 5),
    (7, 1025, 16, 5), (36, 1025, 16, 0),
    # modes 6 / 7: the per-half softmax pipeline over the same units (attention_ph_sm100.cuh; experiment, slower than 4 / 5): half
    # tiles that are empty or ragged at the end of a sequence, one key tile per unit, several units per CTA, the peaked case below
    (1, 128, 1, 6), (2, 197, 2, 6), (2, 197, 2, 7), (1, 2049, 2, 7), (2, 17, 2, 7), (2, 2, 1, 7), (3, 300, 4, 6), (2, 641, 2, 6),
    (20, 257, 12, 6), (40, 65, 8, 6), (40, 65, 8, 7), (300, 129, 2, 7), (7, 1025, 16, 7)])
def test_attention(ops, n_seq, S, heads, mode):
    qkv = _rand(n_seq * S, 3 * heads * 64, seed=21, dtype=torch.bfloat16)
    out = ops.attention_fwd(qkv, n_seq, S, heads, mode)
    _close(out, _attn_ref(qkv, n_seq, S, heads), 2e-2, 2e-2, f"attention S={S} heads={heads} mode={mode}")


@pytest.mark.parametrize("mode", [1, 2, 4, 5, 6, 7])
def test_attention_peaked(ops, mode):
    # large-magnitude scores: exercises the online-softmax rescaling (and the rescaling of the split-off key's weight)
    n_seq, S, heads = 1, 1025, 2
    C = heads * 64
    qkv = _rand(n_seq * S, 3 * C, seed=22, dtype=torch.bfloat16)
    qkv[:, :C] *= 4
    out = ops.attention_fwd(qkv, n_seq, S, heads, mode)
    _close(out, _attn_ref(qkv, n_seq, S, heads), 3e-2, 3e-2, "attention peaked")


def test_attention_modes_agree(ops):
    # the split-off cls token is the same softmax attention, not an approximation
    n_seq, S, heads = 2, 1025, 3
    qkv = _rand(n_seq * S, 3 * heads * 64, seed=31, dtype=torch.bfloat16)
    a = ops.attention_fwd(qkv, n_seq, S, heads, 1).float()
    for mode in (2, 4, 5, 6, 7):
        b = ops.attention_fwd(qkv, n_seq, S, heads, mode).float()
        assert (a - b).abs().max().item() <= 2e-2 * a.abs().max().item(), mode


@pytest.mark.parametrize("n_seq,Lq,Lkv,heads", [(2, 1024, 1024, 8), (1, 100, 333, 2), (3, 256, 64, 1)])
def test_attention_cross(ops, n_seq, Lq, Lkv, heads):
    C = heads * 64
    q = _rand(n_seq * Lq, C, seed=32, dtype=torch.bfloat16)
    kv = _rand(n_seq * Lkv, 2 * C, seed=33, dtype=torch.bfloat16)
    out = ops.attention_cross(q, kv, n_seq, Lq, Lkv, heads)
    qf = q.float().view(n_seq, Lq, heads, 64).transpose(1, 2)
    kf, vf = kv.float().view(n_seq, Lkv, 2, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((qf @ kf.transpose(-1, -2)).softmax(-1) @ vf).transpose(1, 2).reshape(n_seq * Lq, C)
    _close(out, ref, 2e-2, 2e-2, f"cross attention {Lq}x{Lkv}")


def test_groupnorm_relu(ops):
    n_crops, P, C, G = 3, 64, 256, 32
    x = (_rand(n_crops * P, C, seed=23) * 2 + 0.3).to(torch.bfloat16)
    g = _rand(C, seed=24)
    b = _rand(C, seed=25)
    out = ops.groupnorm_relu(x, g, b, n_crops, G, 1e-5)
    xr = x.float().view(n_crops, P, C).permute(0, 2, 1)  # [n, C, P]
    ref = F.relu(F.group_norm(xr, G, g, b, 1e-5)).permute(0, 2, 1).reshape(-1, C)
    _close(out, ref, 1e-2, 1e-2, "groupnorm_relu")


@pytest.mark.parametrize("n,h,w,Cin,Cout", [(2, 8, 8, 256, 128),      # manual pixel-shuffle epilogue
                                           (2, 32, 32, 256, 128),    # 4-D TMA store epilogue (w % 32 == 0)
                                           (3, 64, 64, 128, 64)])
def test_convt_and_cls(ops, n, h, w, Cin, Cout):
    x = _rand(n * h * w, Cin, seed=26, dtype=torch.bfloat16)
    wt = _rand(Cin, Cout, 2, 2, scale=Cin ** -0.5, seed=27)  # ConvTranspose2d weight [Cin, Cout, 2, 2]
    bias = _rand(Cout, seed=28)
    wg = wt.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).to(torch.bfloat16).contiguous()  # row (dy*2+dx)*Cout + co
    out = ops.gemm_convt2x2_gelu(x, wg, bias.repeat(4).contiguous(), Cout, h, w)
    xin = x.float().view(n, h, w, Cin).permute(0, 3, 1, 2)
    ref = F.gelu(F.conv_transpose2d(xin, wt.to(torch.bfloat16).float(), bias, stride=2))  # [n, Cout, 2h, 2w]
    ref_tok = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    _close(out, ref_tok, 1e-2, 1e-2, "convt2x2_gelu")
    nc = 19
    wc = _rand(nc, Cout, scale=Cout ** -0.5, seed=29)
    bc = _rand(nc, seed=30)
    w32 = torch.zeros(32, Cout, device="cuda", dtype=torch.bfloat16)
    w32[:nc] = wc.to(torch.bfloat16)
    logits = ops.gemm_cls_nchw(out, w32, bc, nc, 4 * h * w)
    ref_l = (out.float() @ w32[:nc].float().t() + bc).view(n, 4 * h * w, nc).permute(0, 2, 1)
    _close(logits, ref_l, 2e-3, 2e-3, "cls_nchw")


def _slide_boxes(H, W, ch, cw, sh, sw):
    hg = max(H - ch + sh - 1, 0) // sh + 1
    wg = max(W - cw + sw - 1, 0) // sw + 1
    boxes = []
    for hi in range(hg):
        for wi in range(wg):
            y2 = min(hi * sh + ch, H); x2 = min(wi * sw + cw, W)
            boxes.append((max(y2 - ch, 0), max(x2 - cw, 0)))
    return boxes


@pytest.mark.parametrize("H,W,crop,stride,n_img", [(1024, 2048, 512, 341, 1), (160, 224, 64, 43, 2), (64, 64, 64, 43, 1)])
def test_slide_merge_argmax(ops, H, W, crop, stride, n_img):
    nc = 19
    boxes = _slide_boxes(H, W, crop, crop, stride, stride)
    lh = crop // 4
    low = _rand(n_img * len(boxes), nc, lh, lh, seed=31)
    labels, logits = ops.slide_merge_argmax(low, torch.tensor(boxes, dtype=torch.int32).cuda(), n_img, (crop, crop), (H, W), want_logits=True)
    preds = torch.zeros(n_img, nc, H, W, device="cuda")
    count = torch.zeros(n_img, 1, H, W, device="cuda")
    lv = low.view(n_img, len(boxes), nc, lh, lh)
    for k, (y1, x1) in enumerate(boxes):
        up = F.interpolate(lv[:, k], size=(crop, crop), mode="bilinear", align_corners=False)
        preds += F.pad(up, (x1, W - x1 - crop, y1, H - y1 - crop))
        count[:, :, y1:y1 + crop, x1:x1 + crop] += 1
    assert (count == 0).sum() == 0
    ref = preds / count
    _close(logits, ref, 1e-5, 1e-5, "merged logits")
    ref_lab = ref.argmax(1)
    agree = (labels.long() == ref_lab).float().mean().item()
    assert agree >= 0.9999, f"label agreement {agree}"
    # labels must be the argmax of the kernel's own logits (first max wins), exactly
    assert torch.equal(labels.long(), logits.argmax(1))
    labels2, none = ops.slide_merge_argmax(low, torch.tensor(boxes, dtype=torch.int32).cuda(), n_img, (crop, crop), (H, W))
    assert none is None and torch.equal(labels, labels2)


@pytest.mark.parametrize("H,W,crop,stride,n_img", [(1024, 2048, 512, 341, 1), (160, 224, 64, 43, 2), (64, 100, 64, 43, 1)])
@pytest.mark.parametrize("want_logits", [True, False])
def test_slide_merge_flip_argmax(ops, H, W, crop, stride, n_img, want_logits):
    """The flip-TTA second pass fused into the merge kernel == merge of the mirrored pass + tta_flip_mean_argmax, bit for bit
    (labels and averaged logits), incl. a width that is not a multiple of the 64-pixel tile."""
    from vfmseg_b200.engine import slide_boxes
    nc, lh = 19, crop // 4
    boxes = torch.tensor(slide_boxes(H, W, (crop, crop), (stride, stride)), dtype=torch.int32).cuda()
    low_a = _rand(n_img * boxes.shape[0], nc, lh, lh, seed=61)
    low_b = _rand(n_img * boxes.shape[0], nc, lh, lh, seed=62)      # windows of the mirrored image
    _, a = ops.slide_merge_argmax(low_a, boxes, n_img, (crop, crop), (H, W), want_logits=True)
    _, b = ops.slide_merge_argmax(low_b, boxes, n_img, (crop, crop), (H, W), want_logits=True)
    ref_labels, ref_logits = ops.tta_flip_mean_argmax(a.clone(), b, want_logits=True)
    assert ops.flip_merge_supported(low_b, (crop, crop), W)
    labels, logits = ops.slide_merge_flip_argmax(low_b, boxes, n_img, (crop, crop), a, want_logits=want_logits)
    assert torch.equal(labels, ref_labels)
    if want_logits:
        assert logits is a and torch.equal(logits, ref_logits)
    else:
        assert logits is None
    assert torch.equal(ref_labels.long(), ref_logits.argmax(1))


@pytest.mark.parametrize("H,W,crop,stride,n_img", [(1024, 2048, 512, 341, 2), (1024, 1820, 512, 341, 1), (160, 224, 64, 43, 2),
                                                   (160, 224, 64, 16, 1), (64, 100, 64, 43, 1), (96, 132, 64, 21, 1)])
def test_slide_merge_class_major_equals_window_major(ops, monkeypatch, H, W, crop, stride, n_img):
    """The class-major merge kernel (slide_merge_class_kernel: all footprints staged behind one barrier, one class at a time,
    integer x4 geometry) against the round-1 window-major tile kernel (VFM_MERGE_MODE=1): labels and logits bit for bit, plain and
    fused flip-TTA. The small crops put more than
    four windows over a tile (per-pixel gather inside the class loop), 1820 / 100 / 132 leave edge tiles partly outside."""
    from vfmseg_b200.engine import slide_boxes
    nc, lh = 19, crop // 4
    boxes = torch.tensor(slide_boxes(H, W, (crop, crop), (stride, stride)), dtype=torch.int32).cuda()
    low = _rand(n_img * boxes.shape[0], nc, lh, lh, seed=71)
    low_b = _rand(n_img * boxes.shape[0], nc, lh, lh, seed=72)
    monkeypatch.setenv("VFM_MERGE_MODE", "1")
    lab_old, log_old = ops.slide_merge_argmax(low, boxes, n_img, (crop, crop), (H, W), want_logits=True)
    flab_old, flog_old = ops.slide_merge_flip_argmax(low_b, boxes, n_img, (crop, crop), log_old.clone(), want_logits=True)
    for mode in ("0", "2", "3"):   # 4 (default) / 2 / 3 resident CTAs per SM: three register allocations of the same kernel
        monkeypatch.setenv("VFM_MERGE_MODE", mode)
        lab_new, log_new = ops.slide_merge_argmax(low, boxes, n_img, (crop, crop), (H, W), want_logits=True)
        lab_new2, _ = ops.slide_merge_argmax(low, boxes, n_img, (crop, crop), (H, W))
        flab_new, flog_new = ops.slide_merge_flip_argmax(low_b, boxes, n_img, (crop, crop), log_new.clone(), want_logits=True)
        flab_new2, _ = ops.slide_merge_flip_argmax(low_b, boxes, n_img, (crop, crop), log_new.clone(), want_logits=False)
        assert torch.equal(log_new, log_old) and torch.equal(lab_new, lab_old) and torch.equal(lab_new2, lab_old), mode
        assert torch.equal(flog_new, flog_old) and torch.equal(flab_new, flab_old) and torch.equal(flab_new2, flab_old), mode
    monkeypatch.delenv("VFM_MERGE_MODE")
    assert torch.equal(lab_new.long(), log_new.argmax(1))


@pytest.mark.parametrize("n,P,C,cls", [(3, 64, 256, 1), (2, 1024, 1024, 1), (3, 96, 128, 0), (5, 16, 256, 1), (2, 24, 64, 0)])
def test_patch_embed_gemm_paths(ops, n, P, C, cls):
    """Patch-embed GEMM + bias + pos-embed: the fp32 TMA-store epilogue (patches % 32 == 0: boxes shifted past the cls
    rows in front of them) and the column-per-lane fallback, with and without a cls row (DINOv2 / SAM)."""
    a = _rand(n * P, 768, seed=41, dtype=torch.bfloat16)
    w = _rand(C, 768, scale=768 ** -0.5, seed=42, dtype=torch.bfloat16)
    bias, pos = _rand(C, seed=43), _rand(P + cls, C, seed=44)
    if cls:
        x = ops.gemm_patch_embed(a, w, bias, pos, n, P)
        x.view(n, P + 1, C)[:, 0] = 7.0                      # cls rows are not written by the GEMM
        got = x.view(n, P + 1, C)[:, 1:]
    else:
        got = ops.gemm_patch_embed_nocls(a, w, bias, pos, n, P).view(n, P, C)
    ref = (a.float() @ w.float().t() + bias).view(n, P, C) + pos[cls:]
    _close(got.reshape(-1, C), ref.reshape(-1, C), 2e-3, 2e-3, f"patch_embed n={n} P={P} C={C} cls={cls}")


@pytest.mark.parametrize("B,nc,H,W", [(2, 19, 16, 64), (1, 19, 7, 13), (1, 3, 5, 4), (1, 19, 1024, 2048)])
@pytest.mark.parametrize("want_logits", [True, False])
def test_tta_flip_mean_argmax(ops, B, nc, H, W, want_logits):
    """Bit-exact against the reference's statements (hrda_encoder_decoder.py:199-229, scales = [1]) on the CPU;
    quantised values force ties, which must resolve to the first maximum like torch.argmax on the CPU."""
    g = torch.Generator().manual_seed(7)
    a = (torch.randn(B, nc, H, W, generator=g) * 4).round() / 4
    b = (torch.randn(B, nc, H, W, generator=g) * 4).round() / 4
    a[0, :, 0, :min(W, 4)] = 1.25        # whole-pixel ties across every class
    b[0, :, 0, W - min(W, 4):] = 0.75
    res = torch.zeros_like(a)
    res += a
    res += torch.flip(b, [3])
    want = res / 2
    labels, logits = ops.tta_flip_mean_argmax(a.cuda(), b.cuda(), want_logits=want_logits)
    assert labels.dtype == torch.uint8 and labels.shape == (B, H, W)
    assert torch.equal(labels.cpu().long(), want.argmax(1))
    assert (labels[0, 0, :min(W, 4)] == 0).all()
    if want_logits:
        assert torch.equal(logits.cpu(), want)
    else:
        assert logits is None


@pytest.mark.parametrize("n", [1024 * 2048, 16 * 1000 + 7, 5, 0])
def test_confusion_matrix(ops, n):
    nc = 19
    g = torch.Generator().manual_seed(32)
    pred = torch.randint(0, nc, (n,), generator=g, dtype=torch.uint8)
    label = torch.randint(0, nc + 3, (n,), generator=g, dtype=torch.uint8)
    label[label == nc + 2] = 255
    # piecewise-constant stretch to exercise the run-length path
    if n > 4096:
        pred[:2048] = 3; label[:2048] = 3; label[2048:4096] = 255
    cm = torch.zeros(nc + 1, nc, dtype=torch.int64, device="cuda")
    ops.confusion_matrix_(cm, pred.cuda(), label.cuda(), nc, 255)
    ops.confusion_matrix_(cm, pred.cuda(), label.cuda(), nc, 255)  # accumulates
    keep = label != 255
    p, l = pred[keep].long(), label[keep].long().clamp(max=nc)
    ref = torch.zeros(nc + 1, nc, dtype=torch.int64)
    ref.view(-1).index_add_(0, l * nc + p, torch.ones_like(p))
    assert torch.equal(cm.cpu(), 2 * ref)
