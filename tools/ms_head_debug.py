"""Debug: per-stage error of the VFMHead launch sequence against the fp32 oracle (tiny config)."""
import sys
from pathlib import Path
import torch, torch.nn.functional as F
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vfmseg_b200
from vfmseg_b200 import synthetic, ops
from vfmseg_b200.vfm_refine import context_tokens
from oracle import torch_ref
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
cfg = synthetic.tiny_ms_config()
sd = synthetic.synthetic_ms_state_dict(cfg, seed=0)
model = vfmseg_b200.MODELS.build(dict(cfg)); model.load_state_dict(sd, strict=False); model = model.cuda().eval()
sd3 = torch_ref.split_ms_state_dict(sd); aux = sd3[2]
bb = cfg["backbone"]["backbone"]
x = torch_ref.preprocess(synthetic.synthetic_images(1, 128, 192, seed=1234), MEAN, STD, True)
kw = dict(depth=bb["depth"], num_heads=bb["num_heads"], out_indices=tuple(bb["out_indices"]), lora_scale=2.0)
def rel(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    print(f"{name:28s} rel rms err {(got-ref).pow(2).mean().sqrt()/ref.pow(2).mean().sqrt():.5f}  max/rms {(got-ref).abs().max()/ref.pow(2).mean().sqrt():.4f}")
with torch.no_grad():
    lr = F.interpolate(x, size=(512, 1024), mode="bilinear", align_corners=False)
    low0 = torch_ref.linear_head_forward(torch_ref.dino_forward(lr, sd3[0], **kw), sd3[1])
    seg = F.interpolate(low0, size=x.shape[2:], mode="bilinear", align_corners=False)
    feats = torch_ref.dino_forward(x[:, :, :64, :64], sd3[0], **kw)
    ctxw = seg[:, :, :64, :64]
    # oracle intermediates
    c16 = F.interpolate(ctxw, size=(16, 16), mode="bilinear", align_corners=False)
    f_pre = F.conv2d(torch.cat(feats, 1), aux["fuse_conv.0.weight"], aux["fuse_conv.0.bias"])
    f = F.gelu(F.group_norm(f_pre, 32, aux["fuse_conv.1.weight"], aux["fuse_conv.1.bias"], 1e-5))
    e1p = F.conv2d(c16, aux["seg_logits_embed.0.weight"], aux["seg_logits_embed.0.bias"], stride=2)
    e1 = F.gelu(F.group_norm(e1p, 32, aux["seg_logits_embed.1.weight"], aux["seg_logits_embed.1.bias"], 1e-5))
    e2p = F.conv2d(e1, aux["seg_logits_embed.3.weight"], aux["seg_logits_embed.3.bias"], stride=2)
    e2 = F.gelu(F.group_norm(e2p, 32, aux["seg_logits_embed.4.weight"], aux["seg_logits_embed.4.bias"], 1e-5))
    e3p = F.conv2d(e2, aux["seg_logits_embed.6.weight"], aux["seg_logits_embed.6.bias"])
    e3 = F.group_norm(e3p, 32, aux["seg_logits_embed.7.weight"], aux["seg_logits_embed.7.bias"], 1e-5)
    dec = torch_ref.transformer_decoder_forward(f, e3, aux, heads=8, depth=2)
    out = F.conv2d(dec, aux["conv_seg.weight"], aux["conv_seg.bias"])
tok = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
head = model.aux_decoder.packed()
taps = torch.cat([t.permute(0, 2, 3, 1) for t in feats], -1).reshape(16, -1).to(torch.bfloat16).cuda().contiguous()
crops = torch.tensor([(0, 0, 0, 0)], dtype=torch.int32).cuda()
low0c = low0.cuda().contiguous()
a1 = ops.ms_context_im2col(low0c, crops, (64, 64), (128, 192), (16, 16), head.kpad)
rel("context im2col", a1[:, :76], F.unfold(c16, 2, stride=2)[0].t())
e = ops.gemm_bias_bf16(a1, head.emb1_w, head.emb1_b); rel("emb1 conv", e, tok(e1p))
e = ops.groupnorm_act(e, *head.emb1_gn, 1, 32, 1e-5, act=2); rel("emb1 gn+gelu", e, tok(e1))
e = ops.space_to_depth2(e, 1, 8, 8); e = ops.gemm_bias_bf16(e, head.emb2_w, head.emb2_b); rel("emb2 conv", e, tok(e2p))
e = ops.groupnorm_act(e, *head.emb2_gn, 1, 32, 1e-5, act=2); rel("emb2 gn+gelu", e, tok(e2))
e = ops.gemm_bias_bf16(e, head.emb3_w, head.emb3_b); rel("emb3 conv", e, tok(e3p))
ctx = ops.groupnorm_act(e, *head.emb3_gn, 1, 32, 1e-5, act=0); rel("emb3 gn (ctx tokens)", ctx, tok(e3))
ff = ops.gemm_bias_bf16(taps, head.fuse_w, head.fuse_b); rel("fuse conv", ff, tok(f_pre))
ff = ops.groupnorm_act(ff, *head.fuse_gn, 1, 32, 1e-5, act=2); rel("fuse gn+gelu", ff, tok(f))
from vfmseg_b200.vfm_refine import vfm_head_forward
o = vfm_head_forward(head, taps, ctx, 1, 4, 4); rel("head out (our ctx)", o, out)
o2 = vfm_head_forward(head, taps, tok(e3).to(torch.bfloat16).cuda().contiguous(), 1, 4, 4); rel("head out (oracle ctx)", o2, out)
