"""CPU: host-side logic — registry/config drop-in surface, state-dict naming (peft layout), slide grid,
weight folding, C-ABI symbol table, metric reduction. No GPU compute is issued."""
import ctypes
import os
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import vfmseg_b200
from oracle import torch_ref
from vfmseg_b200 import _C, engine, synthetic
from vfmseg_b200.dg_metrics import DGIoUMetric, areas_from_confusion, total_area_to_metrics
from vfmseg_b200.registry import BACKBONES, METRICS, MODELS, Registry

ROOT = Path(__file__).resolve().parent.parent


def test_registry_surface_matches_reference_usage():
    assert BACKBONES is MODELS   # mmseg.models.builder.BACKBONES is MODELS
    for name in ("DinoVisionTransformer", "LinearHead", "LoraBackboneEncoderDecoder", "SegDataPreProcessor"):
        assert MODELS.get(name) is not None
    assert METRICS.get("DGIoUMetric") is DGIoUMetric
    r = Registry("x")

    @r.register_module()
    class Foo:
        def __init__(self, a, b=2):
            self.a, self.b = a, b
    f = r.build(dict(type="Foo", a=1))
    assert (f.a, f.b) == (1, 2)
    with pytest.raises(KeyError):
        r.build(dict(type="Nope"))
    with pytest.raises(KeyError):
        r.register_module(name="Foo")(type("Bar", (), {}))


def test_reference_config_builds_and_state_dict_names_match_peft_layout():
    cfg = synthetic.tiny_config()
    model = MODELS.build(dict(cfg))
    sd = synthetic.synthetic_state_dict(cfg, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert unexpected == [] and missing == []
    keys = set(model.state_dict())
    assert "backbone.base_model.model.blocks.0.attn.qkv.base_layer.weight" in keys
    assert "backbone.base_model.model.blocks.0.attn.qkv.lora_A.default.weight" in keys
    assert "backbone.base_model.model.blocks.3.attn.qkv.lora_B.default.weight" in keys
    assert "decode_head.fusion_conv.conv.weight" in keys and "decode_head.fusion_conv.gn.bias" in keys
    assert "decode_head.output_upscaling.1.running_var" in keys and "decode_head.conv_seg.bias" in keys
    assert "decode_head.fusion_conv.conv.bias" not in keys   # mmcv ConvModule: no conv bias when a norm follows
    assert model.align_corners is False and model.num_classes == 19 and model.out_channels == 19
    assert model.test_cfg.mode == "slide" and model.test_cfg.crop_size == [64, 64]
    # the plain-backbone checkpoint route of Lora_encoder_decoder.py:28-36
    m2 = MODELS.build(dict(cfg))
    m2.load_pretrained_backbone(synthetic.backbone_checkpoint_from(sd), ["qkv"])
    a = m2.state_dict()["backbone.base_model.model.blocks.1.attn.qkv.base_layer.weight"]
    assert torch.equal(a, sd["backbone.base_model.model.blocks.1.attn.qkv.base_layer.weight"])
    assert m2.state_dict()["backbone.base_model.model.blocks.1.attn.qkv.lora_B.default.weight"].abs().max() == 0  # peft init


def test_full_size_config_has_reference_shapes():
    cfg = synthetic.model_config()
    model = MODELS.build(dict(cfg))
    n = sum(p.numel() for p in model.backbone.parameters())
    assert abs(n - 304.2e6 - 24 * 32 * (1024 + 3072)) < 1.0e6   # 304.2 M backbone + LoRA r=32
    sdk = model.state_dict()
    assert sdk["backbone.base_model.model.pos_embed"].shape == (1, 1025, 1024)
    assert sdk["decode_head.fusion_conv.conv.weight"].shape == (1024, 4096, 1, 1)
    assert sdk["decode_head.output_upscaling.0.weight"].shape == (1024, 512, 2, 2)
    assert sdk["decode_head.conv_seg.weight"].shape == (19, 256, 1, 1)


def test_no_cpu_fallback():
    cfg = synthetic.tiny_config()
    model = MODELS.build(dict(cfg))
    with pytest.raises(RuntimeError, match="CUDA"):
        model.predict_labels(torch.zeros(1, 3, 64, 64, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="CUDA"):
        model.backbone(torch.zeros(1, 3, 64, 64))


def test_slide_boxes_agree_with_oracle():
    for (H, W, c, s) in [(1024, 2048, 512, 341), (1024, 2048, 512, 320), (80, 112, 64, 43), (512, 512, 512, 341), (600, 700, 512, 341)]:
        ours = engine.slide_boxes(H, W, (c, c), (s, s))
        ref = [(b[0], b[2]) for b in torch_ref.slide_boxes(H, W, (c, c), (s, s))]
        assert ours == ref
    b = engine.slide_boxes(1024, 2048, (512, 512), (341, 341))
    cover = np.zeros((1024, 2048), dtype=np.int32)
    for y, x in b:
        cover[y:y + 512, x:x + 512] += 1
    vals, counts = np.unique(cover, return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {1: 524288, 2: 1048576, 4: 524288}


def test_pos_embed_interpolation_matches_oracle():
    g = torch.Generator().manual_seed(0)
    pe = torch.randn(1, 17, 32, generator=g)
    assert torch.equal(engine._interp_pos_embed(pe, 4, 4), pe[0])
    for gh, gw in [(4, 6), (8, 4), (2, 2)]:
        ours = engine._interp_pos_embed(pe, gh, gw)
        ref = torch_ref.interpolate_pos_encoding(pe, gh * gw, gh * 16, gw * 16, 16)[0]
        assert torch.equal(ours, ref)


def test_abi_header_and_binding_agree_and_library_exports_every_symbol():
    hdr = (ROOT / "include" / "vfmseg_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vfm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_C.SIGNATURES), declared ^ set(_C.SIGNATURES)
    lib = _C.load()           # built by __graft_entry__.build(); dlopen works without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vfm_abi_version() == 2
    assert lib.vfm_launch_count() >= 0
    assert ctypes.sizeof(_C.VfmBlockParams) == 20 * 8
    # argument validation happens before any CUDA call
    assert lib.vfm_layernorm(None, None, None, None, 0, 1024, 1e-6, None) == -1
    assert b"layernorm" in lib.vfm_last_error()
    assert lib.vfm_gemm_cls_nchw(None, 0, None, 0, None, None, 40, 1, 1, 64, None) == -1


def test_struct_layouts_of_header_and_binding_agree(tmp_path):
    """Every parameter struct of include/vfmseg_b200.h, as gcc lays it out, against the ctypes mirror in vfmseg_b200/_C.py:
    size and the offset of every field (the fused drivers vfm_vit_forward / vfm_eva_forward / vfm_sam_forward /
    vfm_linear_head_forward read these blocks on the host)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    structs = ["VfmPixelNorm", "VfmBlockParams", "VfmVitParams", "VfmLinearHeadParams", "VfmEvaBlockParams", "VfmEvaParams",
               "VfmSamBlockParams", "VfmSamParams"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include <stdint.h>', f'#include "{ROOT / "include" / "vfmseg_b200.h"}"', "int main(void) {"]
    for st in structs:
        cls = getattr(_C, st)
        lines.append(f'  printf("{st} size %zu\\n", sizeof({st}));')
        for name, _ in cls._fields_:
            lines.append(f'  printf("{st} {name} %zu\\n", offsetof({st}, {name}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        st, field, val = ln.split()
        cls = getattr(_C, st)
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(val) == want, (st, field, val, want)
        seen += 1
    assert seen == sum(len(getattr(_C, st)._fields_) + 1 for st in structs)


def test_metric_reduction_matches_oracle_and_accepts_reference_records():
    rng = np.random.default_rng(0)
    nc = 19
    m = DGIoUMetric(dataset_keys=["citys", "bdd"], ignore_index=255)
    m.dataset_meta = dict(classes=list(range(nc)))
    recs, ref_records = [], []
    for i in range(5):
        pred = rng.integers(0, nc, (64, 96))
        gt = rng.integers(0, nc + 1, (64, 96))
        gt[gt == nc] = 255
        cm = torch.from_numpy(torch_ref.confusion_matrix_np(pred, gt, nc, 255))
        key = "citys" if i % 2 == 0 else "bdd"
        recs.append([key, cm])
        ai, au, ap, al = torch_ref.intersect_and_union(torch.from_numpy(pred), torch.from_numpy(gt), nc, 255)
        ref_records.append([key, ai, au, ap, al])
        got = [a.numpy() for a in areas_from_confusion(cm, nc)]
        assert all(np.array_equal(g, r.numpy().astype(np.int64)) for g, r in zip(got, (ai, au, ap, al)))
    ours = m.compute_metrics(recs)
    theirs = m.compute_metrics(ref_records)     # the reference's own record layout
    assert ours == theirs
    for key in ("citys", "bdd"):
        tot = [sum(r[i] for r in ref_records if r[0] == key) for i in range(1, 5)]
        want = torch_ref.total_area_to_metrics(*[t.numpy() for t in tot])
        for k in ("mIoU", "mAcc", "aAcc"):
            assert ours[f"{key}_{k}"] == pytest.approx(want[k], abs=1e-6)
    assert ours["mean_mIoU"] == pytest.approx((ours["citys_mIoU"] + ours["bdd_mIoU"]) / 2)
    s = total_area_to_metrics(np.array([0, 5]), np.array([0, 10]), np.array([0, 7]), np.array([0, 8]))
    assert s["mIoU"] == 50.0   # nanmean skips the 0/0 class, like the reference


def test_weight_folding_matches_oracle_semantics():
    """LoRA merge, q-scale fold and BatchNorm fold are constant folds of the reference arithmetic."""
    cfg = synthetic.tiny_config()
    sd = synthetic.synthetic_state_dict(cfg, seed=3)
    bb, hd = torch_ref.split_state_dict(sd)
    C = 256
    x = torch.randn(5, C)
    want = torch_ref._qkv(x, bb, "blocks.0.attn.qkv", 2.0)
    w = bb["blocks.0.attn.qkv.base_layer.weight"] + 2.0 * bb["blocks.0.attn.qkv.lora_B.default.weight"] @ bb["blocks.0.attn.qkv.lora_A.default.weight"]
    got = x @ w.t() + bb["blocks.0.attn.qkv.base_layer.bias"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    # ConvT + eval BN fold as used by PackedLinearHead
    import torch.nn.functional as F
    f = torch.randn(2, C, 4, 4)
    y = F.conv_transpose2d(f, hd["output_upscaling.0.weight"], hd["output_upscaling.0.bias"], stride=2)
    y = F.batch_norm(y, hd["output_upscaling.1.running_mean"], hd["output_upscaling.1.running_var"], hd["output_upscaling.1.weight"],
                     hd["output_upscaling.1.bias"], False, 0.0, 1e-5)
    s = hd["output_upscaling.1.weight"] / torch.sqrt(hd["output_upscaling.1.running_var"] + 1e-5)
    w1 = (hd["output_upscaling.0.weight"] * s.view(1, -1, 1, 1)).permute(2, 3, 1, 0).reshape(4 * (C // 2), C)
    b1 = ((hd["output_upscaling.0.bias"] - hd["output_upscaling.1.running_mean"]) * s + hd["output_upscaling.1.bias"]).repeat(4)
    tok = f.permute(0, 2, 3, 1).reshape(-1, C) @ w1.t() + b1            # [n*h*w, 4*Cout], col = (dy*2+dx)*Cout + co
    got = tok.view(2, 4, 4, 2, 2, C // 2).permute(0, 5, 1, 3, 2, 4).reshape(2, C // 2, 8, 8)
    torch.testing.assert_close(got, y, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ checkpoint plumbing / PNG export
def test_convert_dinov2_matches_reference_converter():
    """vfmseg_b200.convert vs tools/convert_models/convert_dinov2.py (imported from the reference tree when present,
    otherwise against the interpolation it is defined by)."""
    import importlib.util
    import torch
    from vfmseg_b200 import convert
    g = torch.Generator().manual_seed(0)
    w = {"patch_embed.proj.weight": torch.randn(32, 3, 14, 14, generator=g), "pos_embed": torch.randn(1, 1 + 37 * 37, 32, generator=g),
         "cls_token": torch.randn(1, 1, 32, generator=g)}
    out = convert.convert_dinov2_state_dict(w, kernel=16, crop_size=(512, 512))
    assert out["patch_embed.proj.weight"].shape == (32, 3, 16, 16) and out["pos_embed"].shape == (1, 1025, 32)
    assert torch.equal(out["pos_embed"][:, :1], w["pos_embed"][:, :1]) and out["cls_token"] is w["cls_token"]
    ref_path = "/root/reference/tools/convert_models/convert_dinov2.py"
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("ref_convert_dinov2", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = {k: v.clone() for k, v in w.items()}
        mod.interpolate_patch_embed_(ref, kernel_conv=16)
        mod.interpolate_pos_embed_(ref, crop_size=(512, 512), kernel_conv=16)
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    ck = {"state_dict": {"decode_head.conv_seg.bias": torch.zeros(19)}}
    convert.merge_backbone_checkpoint(ck, {"cls_token": w["cls_token"]})
    assert set(ck["state_dict"]) == {"decode_head.conv_seg.bias", "backbone.cls_token"}


def _load_ref_script(path, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_convert_sam_and_eva2_match_reference_converters(tmp_path):
    """convert_sam_state_dict / convert_eva2_state_dict / generate_full_weights vs the reference's scripts (run from the
    reference tree when present: their functions, or the script itself on a temporary checkpoint)."""
    import subprocess
    import torch
    from vfmseg_b200 import convert
    g = torch.Generator().manual_seed(1)
    sam = {f"image_encoder.blocks.{i}.norm1.weight": torch.randn(8, generator=g) for i in range(12)}
    sam.update({"image_encoder.patch_embed.proj.weight": torch.randn(8, 3, 16, 16, generator=g),
                "image_encoder.pos_embed": torch.randn(1, 64, 64, 8, generator=g), "mask_decoder.x": torch.zeros(1)})
    out = convert.convert_sam_state_dict(sam, kernel=16, crop_size=(512, 512))
    assert "x" not in out and "mask_decoder.x" not in out and out["pos_embed"].shape == (1, 32, 32, 8)
    assert out["patch_embed.proj.weight"].shape == (8, 3, 16, 16)
    with pytest.raises(KeyError):
        convert.convert_sam_state_dict({"image_encoder.pos_embed": torch.zeros(1, 4, 4, 8)})
    eva = {"model": {"patch_embed.proj.weight": torch.randn(8, 3, 14, 14, generator=g), "pos_embed": torch.randn(1, 1 + 16 * 16, 8, generator=g),
                     "positional_embedding": torch.randn(1 + 16 * 16, 8, generator=g), "rope.freqs_cos": torch.zeros(4), "blocks.0.attn.rope.freqs_sin": torch.zeros(4),
                     "cls_token": torch.randn(1, 1, 8, generator=g)}}
    eo = convert.convert_eva2_state_dict(eva)
    assert not [k for k in eo if "rope" in k] and eo["pos_embed"].shape == (1, 1025, 8) and eo["positional_embedding"].shape == (1025, 8)
    assert eo["patch_embed.proj.weight"].shape == (8, 3, 16, 16) and torch.equal(eo["pos_embed"][:, :1], eva["model"]["pos_embed"][:, :1])
    bb = {"pos_embed": torch.randn(1, 1 + 37 * 37, 1024, generator=g), "patch_embed.proj.weight": torch.randn(4, 3, 14, 14, generator=g)}
    full = convert.generate_full_weights(bb, {"state_dict": {"decode_head.conv_seg.bias": torch.zeros(19)}})
    assert full["state_dict"]["backbone.pos_embed"].shape == (1, 1025, 1024) and full["state_dict"]["backbone.patch_embed.proj.weight"].shape == (4, 3, 16, 16)
    assert torch.equal(full["state_dict"]["backbone.pos_embed"], convert.convert_dinov2_state_dict(bb)["pos_embed"])
    ref_dir = "/root/reference/tools/convert_models"
    if os.path.exists(ref_dir):
        mod = _load_ref_script(f"{ref_dir}/convert_sam.py", "ref_convert_sam")
        ref = mod.select_component({k: v.clone() for k, v in sam.items()}, "image_encoder.")
        mod.interpolate_patch_embed_(ref, kernel_conv=16)
        mod.interpolate_pos_embed_(ref, crop_size=(512, 512), kernel_conv=16)
        assert set(ref) == set(out)
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
        # the EVA02 converter is a script without a main(): run it on a temporary checkpoint
        src, dst = tmp_path / "eva_in.pt", tmp_path / "eva_out.pt"
        torch.save(eva, src)
        subprocess.run([sys.executable, f"{ref_dir}/convert_eva2_512x512.py", str(src), str(dst)], check=True, capture_output=True)
        ref = torch.load(dst, map_location="cpu")
        assert set(ref) == set(eo)
        for k in ref:
            assert torch.equal(eo[k], ref[k]), k
        gen = _load_ref_script("/root/reference/tools/generate_full_weights.py", "ref_generate_full_weights")
        bpath = tmp_path / "bb.pt"
        torch.save(bb, bpath)
        rb = gen.load_backbone(str(bpath))
        for k in rb:
            assert torch.equal(full["state_dict"]["backbone." + k], rb[k]), k


def test_id2color_and_png_export(tmp_path):
    import numpy as np
    from PIL import Image
    from vfmseg_b200 import dg_metrics
    lab = np.array([[0, 1, 18], [19, 255, 7]])
    rgb = dg_metrics.id2color(lab)
    assert rgb.dtype == np.uint8 and rgb.shape == (2, 3, 3)
    assert tuple(rgb[0, 0]) == (128, 64, 128) and tuple(rgb[0, 2]) == (119, 11, 32) and tuple(rgb[1, 2]) == (220, 220, 0)
    assert tuple(rgb[1, 0]) == (0, 0, 0) and tuple(rgb[1, 1]) == (0, 0, 0)      # outside the palette: black (dg_metrics.py:17-21)
    import torch
    m = dg_metrics.DGIoUMetric(dataset_keys=["citys"], output_dir=str(tmp_path), format_only=True)
    m.dataset_meta = dict(classes=list(range(19)))
    m.process({}, [dict(pred_sem_seg=dict(data=torch.from_numpy(lab)[None]), img_path="/x/y/frankfurt_000000.png", seg_map_path="a/citys/b.png")])
    out = np.asarray(Image.open(tmp_path / "frankfurt_000000.png"))
    assert np.array_equal(out, rgb) and m.results == []


# ------------------------------------------------------------------ SAM ViT host logic (BASELINE config 5)
def test_sam_window_maps_match_reference_partition():
    """PackedSam._window_maps against the arithmetic of window_partition / window_unpartition (sam_vit.py:292-346)."""
    import torch.nn.functional as F
    from vfmseg_b200.sam_engine import PackedSam
    n, gh, gw, ws, C = 2, 20, 16, 14, 8
    holder = type("E", (), {"_maps": {}, "device": "cpu"})()
    part, unpart, n_win = PackedSam._window_maps(holder, n, gh, gw, ws)
    x = torch.arange(n * gh * gw * C, dtype=torch.float32).view(n, gh, gw, C) + 1.0
    ph, pw = (ws - gh % ws) % ws, (ws - gw % ws) % ws
    Hp, Wp = gh + ph, gw + pw
    win = F.pad(x, (0, 0, 0, pw, 0, ph)).view(n, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, C)
    flat = torch.cat([x.view(-1, C), torch.zeros(1, C)])            # index -1 -> the zero row
    assert n_win == win.shape[0] // (ws * ws) == 8
    assert torch.equal(flat[part.long()], win)
    assert torch.equal(win[unpart.long()], x.view(-1, C))


def test_sam_table_terms_folded_into_qkv_match_decomposed_rel_pos():
    """The extra qkv output columns (T . W_q) reproduce add_decomposed_rel_pos (sam_vit.py:391-428): gathering
    G_h[qh - kh + K - 1] / G_w[qw - kw + K - 1] equals the reference's two einsums on the unscaled q."""
    from oracle import torch_ref
    from vfmseg_b200 import synthetic
    from vfmseg_b200.sam_engine import PackedSam, SamSpec
    cfg = synthetic.tiny_sam_config(depth=2, global_attn_indexes=(1,), out_indices=(0, 1))
    sd_full = synthetic.synthetic_sam_state_dict(cfg, seed=0)
    pre = "backbone.model.base_model.model."
    sd = {k[len(pre):]: v for k, v in sd_full.items() if k.startswith(pre)}
    bb, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    C, H = bb["embed_dim"], bb["num_heads"]
    d = C // H
    spec = SamSpec(C, bb["depth"], H, 4 * C, 16, tuple(bb["out_indices"]), bb["img_size"] // 16, bb["window_size"],
                   tuple(bb["global_attn_indexes"]))
    scale = lc["lora_alpha"] / lc["r"]
    pk = PackedSam(sd, spec, scale, "cpu")
    g = torch.Generator().manual_seed(5)
    for i, K in ((0, bb["window_size"]), (1, spec.grid)):            # windowed block, global block (interpolated table)
        x = torch.randn(K * K, C, generator=g)
        w, b = pk.blocks[i]["qkv_w"].float(), pk.blocks[i]["qkv_b"]
        y = x @ w.t() + b
        L = 2 * K - 1
        assert w.shape[0] % 32 == 0 and w.shape[0] >= 3 * C + 2 * H * L
        q = y[:, :C].view(K, K, H, d)
        Gh = y[:, 3 * C:3 * C + H * L].view(K, K, H, L)
        Gw = y[:, 3 * C + H * L:3 * C + 2 * H * L].view(K, K, H, L)
        p = f"blocks.{i}.attn."
        Rh = torch_ref.sam_rel_pos_table(K, K, sd[p + "rel_pos_h"])
        Rw = torch_ref.sam_rel_pos_table(K, K, sd[p + "rel_pos_w"])
        rel_h = torch.einsum("hwnc,hkc->hwnk", q, Rh)
        rel_w = torch.einsum("hwnc,wkc->hwnk", q, Rw)
        idx = torch.arange(K)[:, None] - torch.arange(K)[None, :] + K - 1                 # [q, k]
        got_h = torch.gather(Gh, 3, idx[:, None, None, :].expand(K, K, H, K))
        got_w = torch.gather(Gw, 3, idx[None, :, None, :].expand(K, K, H, K))
        # bf16 weights on one side, fp32 on the other: compare at bf16 resolution of the accumulated products
        assert torch.allclose(got_h, rel_h, rtol=2e-2, atol=2e-2), (got_h - rel_h).abs().max()
        assert torch.allclose(got_w, rel_w, rtol=2e-2, atol=2e-2), (got_w - rel_w).abs().max()


def test_relpos_onehot_matrix():
    """The constant one-hot key matrix of vfm_attention_global_tc: row k has 1.0 at column kh(k) and at column bh + kw(k)
    (bh = k_h rounded up to 16), zero rows pad the key count to a multiple of 64, columns pad to 64-wide atoms."""
    from vfmseg_b200 import ops
    for k_h, k_w in ((32, 32), (64, 64), (20, 20), (24, 40)):
        e = ops.relpos_onehot(k_h, k_w, "cpu").float()
        bh, bw = (k_h + 15) // 16 * 16, (k_w + 15) // 16 * 16
        n = k_h * k_w
        assert e.shape == ((n + 63) // 64 * 64, (bh + bw + 63) // 64 * 64)
        assert torch.all(e[:n].sum(1) == 2) and torch.all(e[n:] == 0)
        rel = torch.randn(bh + bw)
        got = e[:n] @ torch.cat([rel, torch.zeros(e.shape[1] - bh - bw)])
        ref = (rel[:k_h, None] + rel[None, bh:bh + k_w]).reshape(-1)      # rel_h[kh] + rel_w[kw], key = kh * k_w + kw
        assert torch.allclose(got, ref)


def test_gelu_polynomial_in_the_kernel_source():
    """The constants of gelu_erf() (csrc/sm100_ptx.cuh) evaluated in fp32 on the CPU against the exact-erf GELU
    (dino_layers/mlp.py:22 = nn.GELU()): absolute error < 2e-6, relative error under half a bf16 ulp above 1e-5."""
    import math
    import re
    from pathlib import Path

    import numpy as np
    from scipy.special import erf
    src = (Path(__file__).resolve().parent.parent / "vfmseg_b200" / "csrc" / "sm100_ptx.cuh").read_text()
    body = src[src.index("float gelu_erf(float x) {"):]
    body = body[:body.index("\n}\n")]
    clamp = np.float32(float(re.search(r"fminf\(fabsf\(x\), ([0-9.eE+-]+)f\)", body).group(1)))
    first = re.search(r"float q = fmaf\(([0-9.eE+-]+)f, t, ([0-9.eE+-]+)f\);", body)
    rest = re.findall(r"q = fmaf\(q, t, ([0-9.eE+-]+)f\);", body)
    co = [np.float32(float(v)) for v in (first.group(1), first.group(2), *rest)]
    assert len(co) >= 5 and "fmaf(-fabsf(x), e, fmaxf(x, 0.f))" in body
    x = np.linspace(-9, 9, 400001).astype(np.float32)
    t = np.minimum(np.abs(x), clamp)
    q = np.full_like(t, co[0])
    for v in co[1:]:
        q = (q * t + v).astype(np.float32)
    e = np.exp2(q.astype(np.float64)).astype(np.float32)
    got = (np.maximum(x, 0) - np.abs(x) * e).astype(np.float64)
    want = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / math.sqrt(2)))
    err = np.abs(got - want)
    assert err.max() < 2e-6
    m = np.abs(want) > 1e-5
    assert (err[m] / np.abs(want[m])).max() < 2e-3


def test_round1_advice_guards():
    """(1) float images through the preprocessor must not silently skip mean/std; (2) a head whose
    in_index is not the identity must not silently get the taps in out_indices order; (3) in-place weight edits on a
    SUBMODULE must change the version the cached engine is keyed on."""
    import vfmseg_b200
    from vfmseg_b200 import synthetic
    cfg = synthetic.tiny_config()
    model = vfmseg_b200.MODELS.build(dict(cfg))
    with pytest.raises(TypeError):
        model.data_preprocessor(dict(inputs=[torch.zeros(3, 64, 64)]))
    v0 = model._weights_version()
    sub = {k: v + 1 for k, v in model.decode_head.state_dict().items() if v.dtype.is_floating_point}
    model.decode_head.load_state_dict(sub, strict=False)
    assert model._weights_version() != v0
    bad = dict(cfg)
    bad["decode_head"] = dict(cfg["decode_head"], in_index=[1, 0, 2, 3])
    m2 = vfmseg_b200.MODELS.build(bad)
    with pytest.raises(NotImplementedError):
        m2.engine()


def test_fold_layernorm_algebra():
    """ops.fold_layernorm: rstd * (bf16(x) @ wf^T - mean * colsum) + bias_f reproduces Linear(LayerNorm(x)) (what
    EpiTmaBf16LN computes from the statistics EpiTmaResidualStats emits), to bf16-operand accuracy, on the CPU."""
    import torch
    import torch.nn.functional as F
    from vfmseg_b200.ops import fold_layernorm
    g = torch.Generator().manual_seed(7)
    M, C, N = 64, 256, 96
    x = torch.randn(M, C, generator=g) * 1.5 + 0.4
    x[:, 3] += 8.0
    ln_w = 1 + 0.2 * torch.randn(C, generator=g)
    ln_b = 0.1 * torch.randn(C, generator=g)
    w = torch.randn(N, C, generator=g) * C ** -0.5
    b = 0.1 * torch.randn(N, generator=g)
    wf, bf, cs = fold_layernorm(w, b, ln_w, ln_b)
    assert wf.dtype == torch.bfloat16 and bf.dtype == torch.float32 and cs.dtype == torch.float32
    xs = x.view(M, C // 128, 128)
    s, q = xs.sum(-1).sum(-1), (xs * xs).sum(-1).sum(-1)   # slot sums added in slot order, as the epilogue does
    mean = s / C
    rstd = torch.rsqrt(q / C - mean * mean + 1e-6)
    acc = x.to(torch.bfloat16).float() @ wf.float().t()
    got = rstd[:, None] * (acc - mean[:, None] * cs[None, :]) + bf[None, :]
    ref = F.linear(F.layer_norm(x, (C,), ln_w, ln_b, 1e-6), w, b)
    # the yardstick: the unfolded kernel pair rounds LayerNorm(x) and W to bf16
    old = F.layer_norm(x, (C,), ln_w, ln_b, 1e-6).to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t() + b
    e_new = (got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()
    e_old = (old - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()
    assert e_new < 5e-3 and e_new < 1.5 * e_old + 1e-4, (float(e_new), float(e_old))


def test_fused_backbone_driver_parameter_blocks_and_host_side_validation():
    """vfm_eva_forward / vfm_sam_forward read host parameter blocks (VfmEvaParams / VfmSamParams). On the CPU: the packers fill them
    from the reference-named state dict (pointers = the packed tensors, taps ascending, padded hidden width), the library's
    workspace arithmetic agrees with the buffer list in its source, and the argument checks (null pointers, head_dim, workspace
    size) answer before any CUDA call."""
    import ctypes as C
    import vfmseg_b200
    from vfmseg_b200 import synthetic
    from vfmseg_b200.eva_engine import PackedEva
    from vfmseg_b200.sam_engine import PackedSam
    lib = _C.load()
    al = lambda n: (n + 255) // 256 * 256

    # ---- EVA02 (tiny: 256 wide, 4 blocks, 4 x 4 patches, hidden 682 -> 688)
    cfg = synthetic.tiny_eva_config()
    model = vfmseg_b200.MODELS.build(dict(cfg))
    model.load_state_dict(synthetic.synthetic_eva_state_dict(cfg, seed=0), strict=False)
    pk = None
    for m in model.modules():
        if hasattr(m, "packed") and hasattr(m, "pt_hw_seq_len"):
            pk = m.packed(torch.device("cpu"))
    assert isinstance(pk, PackedEva)
    p = pk._params
    assert (p.embed_dim, p.depth, p.heads, p.hidden, p.hidden_pad, p.n_taps, p.grid) == (256, 4, 4, 682, 688, 4, 4)
    assert list(p.tap_blocks)[:4] == [0, 1, 2, 3] and abs(p.ln_eps - 1e-5) < 1e-12
    assert p.patch_w == pk.patch_w.data_ptr() and p.rope_cos == pk.rope_cos.data_ptr() and p.ones == pk.ones.data_ptr()
    for blk, b in zip(pk._c_blocks, pk.blocks):
        assert blk.qkv_w == b["qkv_w"].data_ptr() and blk.w12 == b["w12"].data_ptr() and blk.w3 == b["w3"].data_ptr()
        assert blk.qkv_wf == b["qkv_f"][0].data_ptr() and blk.w12_cs == b["w12_f"][2].data_ptr()
        assert b["w12"].shape == (2 * 688, 256) and b["w3"].shape == (256, 688)
    n, P, Cc, Hp = 3, 16, 256, 688
    M = n * (P + 1)
    want = al(M * Cc * 4) + 2 * al(M * Cc * 2) + al(M * 3 * Cc * 2) + al(max(M * 2 * Hp * 2, n * P * 768 * 2)) + al(M * Hp * 2) + al(M * (Cc // 128 + 1) * 8)
    assert lib.vfm_eva_workspace_bytes(C.byref(p), n) == want
    one = C.c_void_p(8)      # a non-null pointer that is never dereferenced: every check below fails before a CUDA call
    assert lib.vfm_eva_forward(None, one, 1, None, 64, 64, one, n, one, one, want, None) == -1
    assert lib.vfm_eva_forward(C.byref(p), one, 1, None, 64, 64, one, n, one, one, want - 1, None) == -3
    assert b"eva_forward: workspace" in lib.vfm_last_error()
    p.heads = 3
    assert lib.vfm_eva_forward(C.byref(p), one, 1, None, 64, 64, one, n, one, one, want, None) == -1
    assert b"head_dim" in lib.vfm_last_error()
    p.heads = 4

    # ---- SAM ViT (tiny: 640 wide = 8 heads x 80, 16 x 16 tokens, 14 x 14 windows, global blocks 1 and 3)
    cfg = synthetic.tiny_sam_config()
    model = vfmseg_b200.MODELS.build(dict(cfg))
    model.load_state_dict(synthetic.synthetic_sam_state_dict(cfg, seed=0), strict=False)
    pk = None
    for m in model.modules():
        if hasattr(m, "packed") and hasattr(m, "global_attn_indexes"):
            pk = m.packed(torch.device("cpu"))
    assert isinstance(pk, PackedSam)
    n = 2
    p = pk._c_params(n)
    assert (p.embed_dim, p.depth, p.heads, p.head_dim, p.hidden, p.n_taps, p.grid, p.use_rel_pos) == (640, 4, 8, 80, 2560, 4, 16, 1)
    part, unpart, n_win = pk._window_maps(n, 16, 16, 14)
    assert p.win_rows == n_win * 196 == n * 4 * 196 and p.part == part.data_ptr() and p.unpart == unpart.data_ptr()
    assert p.win_buf == pk._window_buffer(n_win * 196, 640).data_ptr() and p.onehot and p.onehot_rows == 256
    assert [blk.window for blk in pk._c_blocks] == [14, 0, 14, 0]
    for blk, b in zip(pk._c_blocks, pk.blocks):
        size = 16 if blk.window == 0 else 14
        assert blk.qkv_n == b["qkv_w"].shape[0] == (3 * 640 + 2 * 8 * (2 * size - 1) + 31) // 32 * 32 and blk.qkv_w == b["qkv_w"].data_ptr()
        assert blk.lin1_wf is None                       # 640 % 256 != 0: no LayerNorm folding at this width
    M = n * 256
    qkv_elems = max(M * pk._c_blocks[1].qkv_n, p.win_rows * pk._c_blocks[0].qkv_n)
    want = al(M * 640 * 4) + 2 * al(M * 640 * 2) + al(p.win_rows * 640 * 2) + al(qkv_elems * 2) + al(max(M * 2560 * 2, M * 768 * 2)) + al(M * (640 // 128 + 1) * 8)
    assert lib.vfm_sam_workspace_bytes(C.byref(p), n) == want
    assert lib.vfm_sam_forward(C.byref(p), one, 1, None, 256, 256, one, n, one, None, want, None) == -1
    assert lib.vfm_sam_forward(C.byref(p), one, 1, None, 256, 256, one, n, one, one, want - 1, None) == -3
    assert b"sam_forward: workspace" in lib.vfm_last_error()
