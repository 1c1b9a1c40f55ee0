"""Per-operator parity of the coarse-to-fine (MsVFMEncoderDecoder / VFMHead) kernels, through the C ABI, against
plain torch fp32 restatements of the reference lines each one replaces."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


@pytest.fixture(scope="module")
def ops():
    from vfmseg_b200 import _C, ops
    _C.check(_C.load().vfm_device_check())
    return ops


def _rand(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def _boxes(H, W, crop, stride):
    from vfmseg_b200.engine import slide_boxes
    return slide_boxes(H, W, crop, stride)


@pytest.mark.parametrize("H,W,h,w", [(64, 128, 32, 64), (96, 160, 48, 64), (40, 40, 24, 56)])
def test_image_resize_norm(ops, H, W, h, w):
    # Ms_VFM_encoder_decoder.py:413: resize(inputs, size=..., mode='bilinear', align_corners=False) after SegDataPreProcessor
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 3, H, W), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(MEAN).view(1, 3, 1, 1), torch.tensor(STD).view(1, 3, 1, 1)
    x = (u8[:, [2, 1, 0]].float() - mean) / std
    ref = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
    got_u8 = ops.image_resize_norm(u8.cuda(), (h, w), ops.pixel_norm(MEAN, STD, True)).cpu()
    got_f = ops.image_resize_norm(x.cuda().contiguous(), (h, w)).cpu()
    assert (got_u8 - ref).abs().max() < 2e-5 * 4
    assert (got_f - ref).abs().max() < 1e-5


def test_ms_confidence(ops):
    # Ms_VFM_encoder_decoder.py:443-448 on the upsampled coarse logits (:417-420)
    H, W, crop, stride, nc = 96, 160, (64, 64), (40, 40), 19
    low0 = _rand(2, nc, H // 8, W // 8, scale=3.0, seed=5)
    boxes = _boxes(H, W, crop, stride)
    bt = torch.tensor(boxes, dtype=torch.int32).cuda()
    thr = 0.6
    got = ops.ms_confidence(low0, bt, crop, (H, W), thr).cpu()
    U = F.interpolate(low0.cpu(), size=(H, W), mode="bilinear", align_corners=False)
    for b in range(2):
        for k, (y1, x1) in enumerate(boxes):
            ctx = U[b:b + 1, :, y1:y1 + crop[0], x1:x1 + crop[1]]
            ref = int((torch.softmax(ctx, dim=1).max(dim=1)[0] > thr).sum())
            assert abs(int(got[b, k]) - ref) <= 2, (b, k, int(got[b, k]), ref)   # pixels within float rounding of thr


def test_ms_context_im2col(ops):
    # VFMHead.py:63-67 (context resized to 4x the feature grid) + the k=2, s=2 patches of seg_logits_embed[0] (:38-39)
    H, W, crop, nc = 96, 160, (64, 64), 19
    low0 = _rand(2, nc, H // 8, W // 8, scale=2.0, seed=6)
    crops = [(0, 0, 0, 0), (1, 32, 96, 0), (0, 17, 41, 0)]
    ct = torch.tensor(crops, dtype=torch.int32).cuda()
    ctx_hw = (16, 16)
    got = ops.ms_context_im2col(low0, ct, crop, (H, W), ctx_hw, 80).float().cpu()
    U = F.interpolate(low0.cpu(), size=(H, W), mode="bilinear", align_corners=False)
    for r, (b, y1, x1, _) in enumerate(crops):
        ctx = F.interpolate(U[b:b + 1, :, y1:y1 + crop[0], x1:x1 + crop[1]], size=ctx_hw, mode="bilinear", align_corners=False)
        cols = F.unfold(ctx, kernel_size=2, stride=2)[0].t()          # [oh*ow, nc*4], column = cin*4 + dy*2 + dx
        blk = got[r * 64:(r + 1) * 64]
        assert (blk[:, :nc * 4] - cols).abs().max() < 2e-2 * cols.abs().max()
        assert (blk[:, nc * 4:] == 0).all()


def test_space_to_depth2(ops):
    n, h, w, Cc = 2, 8, 12, 64
    x = _rand(n * h * w, Cc, seed=7, dtype=torch.bfloat16)
    got = ops.space_to_depth2(x, n, h, w).cpu()
    t = x.cpu().view(n, h // 2, 2, w // 2, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(n * (h // 2) * (w // 2), 4 * Cc)
    assert torch.equal(got, t)


@pytest.mark.parametrize("C,groups,act,out_f32,eps", [(64, 32, 2, False, 1e-5), (128, 32, 2, False, 1e-5), (256, 32, 0, False, 1e-5),
                                                     (256, 32, 2, False, 1e-5), (256, 32, 0, True, 1e-6), (64, 8, 1, False, 1e-5)])
def test_groupnorm_act(ops, C, groups, act, out_f32, eps):
    n, P = 3, 96
    x = (_rand(n * P, C, seed=8) * 2 + 0.3).to(torch.bfloat16)
    g, b = _rand(C, seed=9), _rand(C, seed=10)
    got = ops.groupnorm_act(x, g, b, n, groups, eps, act, out_f32).float()
    xr = x.float().view(n, P, C).permute(0, 2, 1)
    ref = F.group_norm(xr, groups, g, b, eps)
    ref = {0: lambda t: t, 1: F.relu, 2: F.gelu}[act](ref).permute(0, 2, 1).reshape(-1, C)
    tol = 1e-4 if out_f32 else 1e-2
    assert ((got - ref).abs() <= tol + tol * ref.abs()).all(), (got - ref).abs().max().item()


def test_geglu(ops):
    M, I = 300, 1024
    x = _rand(M, 2 * I, seed=11, dtype=torch.bfloat16)
    got = ops.geglu(x).float()
    a, gate = x.float().chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    assert ((got - ref).abs() <= 1e-2 + 1e-2 * ref.abs()).all()


def test_cast(ops):
    x = _rand(77, 256, seed=12)
    assert torch.equal(ops.cast_f32_bf16(x), x.to(torch.bfloat16))


@pytest.mark.parametrize("refine", ["none", "some", "all"])
def test_ms_merge_argmax(ops, refine):
    # Ms_VFM_encoder_decoder.py:433-461 with the loop body of :449-459
    H, W, crop, stride, nc = 96, 160, (64, 64), (40, 40), 19
    n_img = 2
    low0 = _rand(n_img, nc, H // 8, W // 8, scale=2.0, seed=13)
    boxes = _boxes(H, W, crop, stride)
    nk = len(boxes)
    bt = torch.tensor(boxes, dtype=torch.int32).cuda()
    g = torch.Generator().manual_seed(14)
    mask = {"none": torch.zeros(n_img, nk, dtype=torch.bool), "all": torch.ones(n_img, nk, dtype=torch.bool),
            "some": torch.rand(n_img, nk, generator=g) > 0.5}[refine]
    ref_index = torch.full((n_img, nk), -1, dtype=torch.int32)
    ref_index[mask] = torch.arange(int(mask.sum()), dtype=torch.int32)
    n_ref = int(mask.sum())
    refined = _rand(max(n_ref, 1), nc, 4, 4, scale=2.0, seed=15)
    labels, logits = ops.ms_merge_argmax(low0, refined, ref_index.cuda(), bt, crop, (H, W), want_logits=True)
    U = F.interpolate(low0.cpu(), size=(H, W), mode="bilinear", align_corners=False)
    preds = torch.zeros(n_img, nc, H, W)
    count = torch.zeros(n_img, 1, H, W)
    for k, (y1, x1) in enumerate(boxes):
        for b in range(n_img):
            ctx = U[b:b + 1, :, y1:y1 + crop[0], x1:x1 + crop[1]]
            if ref_index[b, k] >= 0:
                lg = F.interpolate(refined[int(ref_index[b, k])][None].cpu(), size=crop, mode="bilinear", align_corners=False)
            else:
                lg = ctx
            preds[b:b + 1] += F.pad(lg, (x1, W - x1 - crop[1], y1, H - y1 - crop[0]))
            count[b, :, y1:y1 + crop[0], x1:x1 + crop[1]] += 1
    ref = preds / count
    assert (logits.cpu() - ref).abs().max() < 1e-4
    assert (labels.cpu().long() == logits.cpu().argmax(1)).all()


@pytest.mark.parametrize("H,W,crop,stride,lg0,lgr,n_img", [(1024, 2048, 512, 341, 3, 4, 2), (96, 160, 64, 40, 3, 4, 2), (128, 192, 64, 33, 3, 4, 1),
                                                          (64, 128, 64, 43, 2, 2, 1), (96, 132, 64, 21, 2, 5, 1)])
@pytest.mark.parametrize("refine", ["none", "some", "all"])
def test_ms_merge_class_major_equals_pixel_kernel(ops, monkeypatch, H, W, crop, stride, lg0, lgr, n_img, refine):
    """The class-major tile kernel of the stage-1 merge (ms_merge_class_kernel: context + refined-window footprints staged in shared
    memory, one class at a time, integer geometry for power-of-two upsampling) against the per-pixel gather kernel
    (VFM_MERGE_MODE=1): labels and logits bit for bit; full BASELINE config 3 size, tiles under more than four windows, ragged
    widths, upsampling factors 4 .. 32, no / some / all windows refined."""
    nc = 19
    low0 = _rand(n_img, nc, H >> lg0, W >> lg0, scale=2.0, seed=23)
    boxes = _boxes(H, W, (crop, crop), (stride, stride))
    nk = len(boxes)
    bt = torch.tensor(boxes, dtype=torch.int32).cuda()
    g = torch.Generator().manual_seed(24)
    mask = {"none": torch.zeros(n_img, nk, dtype=torch.bool), "all": torch.ones(n_img, nk, dtype=torch.bool),
            "some": torch.rand(n_img, nk, generator=g) > 0.5}[refine]
    ref_index = torch.full((n_img, nk), -1, dtype=torch.int32)
    n_ref = int(mask.sum())
    ref_index[mask] = torch.randperm(n_ref, generator=g).to(torch.int32)
    refined = _rand(max(n_ref, 1), nc, crop >> lgr, crop >> lgr, scale=2.0, seed=25) if n_ref else None
    ri = ref_index.cuda()
    monkeypatch.setenv("VFM_MERGE_MODE", "1")
    lab_old, log_old = ops.ms_merge_argmax(low0, refined, ri, bt, (crop, crop), (H, W), want_logits=True)
    for mode in ("0", "3"):   # 2 (default) / 3 resident CTAs per SM
        monkeypatch.setenv("VFM_MERGE_MODE", mode)
        lab, log = ops.ms_merge_argmax(low0, refined, ri, bt, (crop, crop), (H, W), want_logits=True)
        lab2, none = ops.ms_merge_argmax(low0, refined, ri, bt, (crop, crop), (H, W))
        assert none is None
        assert torch.equal(log, log_old) and torch.equal(lab, lab_old) and torch.equal(lab2, lab_old), mode
    monkeypatch.delenv("VFM_MERGE_MODE")
    assert torch.equal(lab_old.long(), log_old.argmax(1))
