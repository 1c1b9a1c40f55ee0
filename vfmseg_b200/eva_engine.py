"""Engine side of the EVA02 backbone (BASELINE config 4; rein/models/backbones/eva_02.py:614-853): weight packing
(LoRA merge of q/k/v/proj, q-scale fold, fused q|k|v and w1|w2 operands, hidden width padded 2730 -> 2736 so rows stay
16-byte aligned for TMA) and the per-block launch sequence. Same kernels as the DINOv2 path plus RoPE and SwiGLU+LN.

Block (eva_02.py:486-493, no LayerScale):  x += attn(norm1(x));  x += mlp(norm2(x))
  attn (:331-381, subln=True, xattn=True): q = x Wq^T + q_bias, k = x Wk^T, v = x Wv^T + v_bias; RoPE on the patch tokens
       of q and k; softmax(q k^T / sqrt d) v; proj
  mlp  (:234-241): w3(LayerNorm(silu(w1 x) * w2 x))
nn.LayerNorm eps is 1e-5 in the blocks: EVA2.__init__ does not forward the configured norm_layer to Block (:705-727).
"""
from __future__ import annotations

import ctypes as C_
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from . import _C, ops


@dataclass
class EvaSpec:
    embed_dim: int
    depth: int
    num_heads: int
    hidden: int            # int(embed_dim * mlp_ratio) = 2730 for ViT-L
    patch_size: int
    out_indices: Tuple[int, ...]
    grid: int              # img_size // patch_size (pos_embed / RoPE are fixed to this grid, eva_02.py:690-697,825-826)
    pt_hw_seq_len: int = 16
    ln_eps: float = 1e-5


def rope_tables(head_dim: int, pt_seq_len: int, ft_seq_len: int):
    """VisionRotaryEmbeddingFast.__init__, eva_02.py:119-156 (freqs_for='lang', theta 10000): cos/sin [ft*ft, head_dim]."""
    dim = head_dim // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
    f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)          # [ft, dim]
    f2 = torch.cat((f[:, None, :].expand(ft_seq_len, ft_seq_len, dim), f[None, :, :].expand(ft_seq_len, ft_seq_len, dim)), dim=-1)
    return f2.cos().reshape(-1, 2 * dim).contiguous(), f2.sin().reshape(-1, 2 * dim).contiguous()


def _lora_merged(sd, key: str, scale: float) -> torch.Tensor:
    """peft lora.Linear at inference: W + (alpha/r) B A when the module was wrapped, else the plain weight."""
    if key + ".base_layer.weight" in sd:
        w = sd[key + ".base_layer.weight"].float()
        a, b = key + ".lora_A.default.weight", key + ".lora_B.default.weight"
        if a in sd:
            w = w + scale * (sd[b].float() @ sd[a].float())
        return w
    return sd[key + ".weight"].float()


def _bias(sd, key: str):
    for k in (key + ".base_layer.bias", key + ".bias"):
        if k in sd:
            return sd[k].float()
    return None


def _ln_fold_bits() -> int:
    """VFM_LN_FOLD (same meaning as in vfm_vit_forward): bit 0 = norm1 folded into qkv (statistics from the previous block's
    last GEMM), bit 1 = norm2 folded into the first MLP GEMM (statistics from attn.proj). Default 3 (both)."""
    import os
    return int(os.environ.get("VFM_LN_FOLD", "3"))


class PackedEva:
    def __init__(self, sd: Dict[str, torch.Tensor], spec: EvaSpec, lora_scale: float, device):
        self.spec, self.device = spec, device
        sd = {k: v.detach().cpu() for k, v in sd.items()}      # fold on the host, upload once
        C, H = spec.embed_dim, spec.hidden
        if C // spec.num_heads != 64:
            raise ValueError("vfmseg_b200 attention kernel needs head_dim 64")
        if spec.patch_size != 16:
            raise ValueError("vfmseg_b200 patch gather is built for patch_size 16")
        self.Hp = Hp = (H + 7) // 8 * 8
        keep: List[torch.Tensor] = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        f32, bf = torch.float32, torch.bfloat16
        scale = 64 ** -0.5
        self.patch_w = dev(sd["patch_embed.proj.weight"].reshape(C, -1), bf)
        self.patch_b = dev(sd["patch_embed.proj.bias"], f32)
        self.cls_token = dev(sd["cls_token"].reshape(-1), f32)
        self.pos = dev(sd["pos_embed"][0], f32)                                  # [1 + grid^2, C]
        cos, sin = rope_tables(64, spec.pt_hw_seq_len, spec.grid)
        self.rope_cos, self.rope_sin = dev(cos, f32), dev(sin, f32)
        self.ones = dev(torch.ones(C), f32)
        self.blocks = []
        for i in range(spec.depth):
            p = f"blocks.{i}."
            # eva_02.py:337-339 calls F.linear(x, self.q_proj.weight, ...): with peft that attribute is the BASE weight,
            # so the adapters on q_proj / k_proj / v_proj are inert in the reference's forward pass (scale 0 here);
            # attn.proj is called as a module (:379) and is adapted.
            wq = _lora_merged(sd, p + "attn.q_proj", 0.0) * scale
            wk = _lora_merged(sd, p + "attn.k_proj", 0.0)
            wv = _lora_merged(sd, p + "attn.v_proj", 0.0)
            qb = sd[p + "attn.q_bias"].float() * scale if p + "attn.q_bias" in sd else torch.zeros(C)
            vb = sd[p + "attn.v_bias"].float() if p + "attn.v_bias" in sd else torch.zeros(C)
            w12 = torch.zeros(2 * Hp, C)
            b12 = torch.zeros(2 * Hp)
            w12[:H], w12[Hp:Hp + H] = _lora_merged(sd, p + "mlp.w1", lora_scale), _lora_merged(sd, p + "mlp.w2", lora_scale)
            b12[:H], b12[Hp:Hp + H] = _bias(sd, p + "mlp.w1"), _bias(sd, p + "mlp.w2")
            w3 = torch.zeros(C, Hp)
            w3[:, :H] = _lora_merged(sd, p + "mlp.w3", lora_scale)
            fold = {}
            if C % 256 == 0:   # LayerNorm folded into the Linear behind it (ops.fold_layernorm; vfm_gemm_lnfold_bf16)
                wf, bfold, cs = ops.fold_layernorm(torch.cat([wq, wk, wv], 0), torch.cat([qb, torch.zeros(C), vb]),
                                                   sd[p + "norm1.weight"], sd[p + "norm1.bias"])
                fold["qkv_f"] = (dev(wf, bf), dev(bfold, f32), dev(cs, f32))
                wf, bfold, cs = ops.fold_layernorm(w12, b12, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
                fold["w12_f"] = (dev(wf, bf), dev(bfold, f32), dev(cs, f32))
            self.blocks.append(dict(
                **fold,
                n1=(dev(sd[p + "norm1.weight"], f32), dev(sd[p + "norm1.bias"], f32)),
                n2=(dev(sd[p + "norm2.weight"], f32), dev(sd[p + "norm2.bias"], f32)),
                qkv_w=dev(torch.cat([wq, wk, wv], 0), bf), qkv_b=dev(torch.cat([qb, torch.zeros(C), vb]), f32),
                proj_w=dev(_lora_merged(sd, p + "attn.proj", lora_scale), bf), proj_b=dev(_bias(sd, p + "attn.proj"), f32),
                w12=dev(w12, bf), b12=dev(b12, f32),
                ffn_ln=(dev(sd[p + "mlp.ffn_ln.weight"], f32), dev(sd[p + "mlp.ffn_ln.bias"], f32)),
                w3=dev(w3, bf), b3=dev(_bias(sd, p + "mlp.w3"), f32)))
        self._keep = keep
        self._ws = None
        self._params = self._c_params()

    def _c_params(self) -> _C.VfmEvaParams:
        """The host-side parameter block of vfm_eva_forward (include/vfmseg_b200.h): pointers into the packed weights."""
        s = self.spec
        outs = sorted(s.out_indices)
        if len(outs) > 8:
            raise ValueError("at most 8 feature taps")
        self._c_blocks = (_C.VfmEvaBlockParams * s.depth)()
        for blk, b in zip(self._c_blocks, self.blocks):
            blk.ln1_w, blk.ln1_b = (t.data_ptr() for t in b["n1"])
            blk.ln2_w, blk.ln2_b = (t.data_ptr() for t in b["n2"])
            blk.qkv_w, blk.qkv_b = b["qkv_w"].data_ptr(), b["qkv_b"].data_ptr()
            blk.proj_w, blk.proj_b = b["proj_w"].data_ptr(), b["proj_b"].data_ptr()
            blk.w12, blk.b12 = b["w12"].data_ptr(), b["b12"].data_ptr()
            blk.ffn_ln_w, blk.ffn_ln_b = (t.data_ptr() for t in b["ffn_ln"])
            blk.w3, blk.b3 = b["w3"].data_ptr(), b["b3"].data_ptr()
            if "qkv_f" in b:
                blk.qkv_wf, blk.qkv_bf, blk.qkv_cs = (t.data_ptr() for t in b["qkv_f"])
                blk.w12_wf, blk.w12_bf, blk.w12_cs = (t.data_ptr() for t in b["w12_f"])
        p = _C.VfmEvaParams()
        p.embed_dim, p.depth, p.heads, p.hidden, p.hidden_pad = s.embed_dim, s.depth, s.num_heads, s.hidden, self.Hp
        p.n_taps, p.grid, p.ln_eps = len(outs), s.grid, s.ln_eps
        for i, o in enumerate(outs):
            p.tap_blocks[i] = o
        p.patch_w, p.patch_b = self.patch_w.data_ptr(), self.patch_b.data_ptr()
        p.cls_token, p.pos_embed = self.cls_token.data_ptr(), self.pos.data_ptr()
        p.rope_cos, p.rope_sin, p.ones = self.rope_cos.data_ptr(), self.rope_sin.data_ptr(), self.ones.data_ptr()
        p.blocks = C_.cast(self._c_blocks, C_.POINTER(_C.VfmEvaBlockParams))
        return p

    # SlideEngine.backbone_taps dispatches here
    def forward_taps(self, img: torch.Tensor, crops: torch.Tensor, gh: int, gw: int, pixel_norm) -> torch.Tensor:
        """EVA2.forward_features (eva_02.py:816-849) for the listed windows -> taps bf16 [n*gh*gw, n_taps*C] (token-major,
        cls dropped): the un-normalised residual stream after the blocks in out_indices. One call of the fused C driver
        vfm_eva_forward; VFM_EVA_DRIVER=py issues the same launch sequence operator by operator from Python (the comparison path
        of tests/test_eva_gpu.py: bit-identical taps)."""
        s = self.spec
        if gh != s.grid or gw != s.grid:
            raise _C.VfmError(f"EVA02 adds a fixed {s.grid}x{s.grid} pos_embed / RoPE table without interpolation "
                              f"(eva_02.py:690-697,825-826): windows must be {s.grid * 16}x{s.grid * 16}, got grid {gh}x{gw}")
        n, P, C = crops.shape[0], gh * gw, s.embed_dim
        if os.environ.get("VFM_EVA_DRIVER", "c") != "py":
            is_u8 = img.dtype == torch.uint8
            if is_u8 and pixel_norm is None:
                raise _C.VfmError("uint8 input needs set_pixel_norm() (SegDataPreProcessor mean/std)")
            if not is_u8 and img.dtype != torch.float32:
                raise _C.VfmError(f"input must be uint8 or float32, got {img.dtype}")
            assert img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
            assert crops.dtype == torch.int32 and crops.is_contiguous() and crops.shape[1] == 4
            lib = _C.load()
            need = lib.vfm_eva_workspace_bytes(C_.byref(self._params), n)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            taps = torch.empty(n * P, len(s.out_indices) * C, dtype=torch.bfloat16, device=self.device)
            _C.call("vfm_eva_forward", C_.byref(self._params), img.data_ptr(), int(is_u8),
                    C_.byref(pixel_norm) if (is_u8 and pixel_norm is not None) else None, img.shape[2], img.shape[3],
                    crops.data_ptr(), n, taps.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                    torch.cuda.current_stream().cuda_stream)
            return taps
        T = P + 1
        a = ops.patch_gather(img, crops, gh, gw, pixel_norm if img.dtype == torch.uint8 else None)
        x = ops.gemm_patch_embed(a, self.patch_w, self.patch_b, self.pos, n, P)        # :818-826 (+ pos_embed[1:])
        ops.cls_rows_(x, self.cls_token, self.pos, n, T)                              # cls_token + pos_embed[0]
        outs = sorted(s.out_indices)
        taps = torch.empty(n * P, len(outs) * C, dtype=torch.bfloat16, device=x.device)
        # LayerNorm folding (csrc/gemm_sm100.cuh, EpiTmaResidualStats -> EpiTmaBf16LN): a residual GEMM followed by a plain norm
        # + Linear also emits bf16(x) and the row statistics; norm1 of block 0 and the norm1 passes that write a tap stay kernels
        bits = _ln_fold_bits() if "qkv_f" in self.blocks[0] else 0
        fold1, fold2 = bool(bits & 1), bool(bits & 2)
        pre = None                                                                    # (bf16(x), stats) from the previous block's w3 GEMM
        for i, b in enumerate(self.blocks):
            tap_i = outs.index(i - 1) if (i - 1) in outs else None                    # tap of the previous block's output
            if pre is not None:
                qkv = ops.gemm_lnfold_rope_bf16(*pre, *b["qkv_f"], s.ln_eps, self.rope_cos, self.rope_sin, 2 * C, T)
            else:
                h = ops.layernorm_tap(x, *b["n1"], s.ln_eps, taps if tap_i is not None else None,
                                      (tap_i or 0) * C, T)
                qkv = ops.gemm_bias_rope_bf16(h, b["qkv_w"], b["qkv_b"], self.rope_cos, self.rope_sin, 2 * C, T)   # :337-369
            att = ops.attention_fwd(qkv, n, T, s.num_heads)                           # xops.memory_efficient_attention :376
            if fold2:
                xb, st = ops.gemm_bias_ls_residual_stats_(x, att, b["proj_w"], b["proj_b"], self.ones)
                h12 = ops.gemm_lnfold_bf16(xb, st, *b["w12_f"], s.ln_eps)
            else:
                ops.gemm_bias_ls_residual_(x, att, b["proj_w"], b["proj_b"], self.ones)
                h12 = ops.gemm_bias_bf16(ops.layernorm(x, *b["n2"], s.ln_eps), b["w12"], b["b12"])
            u = ops.swiglu_layernorm(h12, *b["ffn_ln"], s.hidden, s.ln_eps)
            if fold1 and i + 1 < s.depth and i not in outs:
                pre = ops.gemm_bias_ls_residual_stats_(x, u, b["w3"], b["b3"], self.ones)
            else:
                pre = None
                ops.gemm_bias_ls_residual_(x, u, b["w3"], b["b3"], self.ones)
        if (s.depth - 1) in outs:
            ops.layernorm_tap(x, None, None, s.ln_eps, taps, outs.index(s.depth - 1) * C, T, want_out=False)
        return taps
