"""Engine side of the SAM ViT backbone (BASELINE config 5; rein/models/backbones/sam_vit.py:55-147): weight packing (LoRA
merge of qkv, rel-pos tables interpolated and gathered once per block) and the per-block launch sequence.

Block (sam_vit.py:201-217):  x += attn(window_partition(norm1(x)));  x += mlp(norm2(x))
  windowed blocks (window_size > 0): norm1(x) is zero-padded to a multiple of the window AFTER the norm (:306-310) and cut
       into 14 x 14 windows, so padded tokens enter attention as q = k = v = qkv bias; attention per window with the
       decomposed rel-pos bias of the 14 x 14 grid; windows are stitched back and the padding dropped (:335-346). proj is
       per token, so it runs after the un-partition on the real tokens only.
  global blocks: attention over the whole token grid; their rel-pos tables hold 4 * size - 1 entries (:248-254) and are
       linearly interpolated to 2 * size - 1 by get_rel_pos (:372-379).
  mlp (:17-29): lin2(GELU_erf(lin1(x))). No cls token, no LayerScale; pos_embed added without interpolation (:131-132).
"""
from __future__ import annotations

import ctypes as C_
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import _C, ops
from .eva_engine import _ln_fold_bits


@dataclass
class SamSpec:
    embed_dim: int
    depth: int
    num_heads: int
    hidden: int
    patch_size: int
    out_indices: Tuple[int, ...]
    grid: int                          # img_size // patch_size: pos_embed is fixed to this grid
    window_size: int
    global_attn_indexes: Tuple[int, ...]
    use_rel_pos: bool = True
    ln_eps: float = 1e-6


def rel_pos_resized(size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """The (2 size - 1, head_dim) table get_rel_pos indexes with q - k + size - 1 (sam_vit.py:372-388): linear
    interpolation when the stored table has another length (global blocks store 4 size - 1 entries, :248-254)."""
    L = 2 * size - 1
    r = rel_pos.float()
    if r.shape[0] != L:
        r = F.interpolate(r.reshape(1, r.shape[0], -1).permute(0, 2, 1), size=L, mode="linear").reshape(-1, L).permute(1, 0)
    return r.contiguous()


def _lora_merged(sd, key: str, scale: float) -> torch.Tensor:
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


class PackedSam:
    def __init__(self, sd: Dict[str, torch.Tensor], spec: SamSpec, lora_scale: float, device):
        self.spec, self.device = spec, device
        sd = {k: v.detach().cpu() for k, v in sd.items()}      # fold on the host, upload once
        C = spec.embed_dim
        self.head_dim = d = C // spec.num_heads
        if d not in (64, 80):
            raise ValueError(f"vfmseg_b200 rel-pos attention kernel is built for head_dim 64 or 80, got {d}")
        if spec.patch_size != 16:
            raise ValueError("vfmseg_b200 patch gather is built for patch_size 16")
        if C % 128 or spec.hidden % 32:
            raise ValueError("embed_dim must be a multiple of 128 (LayerNorm kernel) and the MLP width of 32")
        keep: List[torch.Tensor] = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        f32, bf = torch.float32, torch.bfloat16
        self.patch_w = dev(sd["patch_embed.proj.weight"].reshape(C, -1), bf)
        self.patch_b = dev(sd["patch_embed.proj.bias"], f32)
        self.pos = dev(sd["pos_embed"].reshape(-1, C), f32)                      # [grid^2, C]
        self.ones = dev(torch.ones(C), f32)
        g, ws = spec.grid, spec.window_size
        self.blocks = []
        for i in range(spec.depth):
            p = f"blocks.{i}."
            size = g if i in spec.global_attn_indexes or ws == 0 else ws
            wqkv, bqkv = _lora_merged(sd, p + "attn.qkv", lora_scale), _bias(sd, p + "attn.qkv")
            if bqkv is None:
                bqkv = torch.zeros(3 * C)
            if spec.use_rel_pos:
                # Table terms as extra output columns of the qkv GEMM. rel_h[q, kh] = q . T_h[qh - kh + size - 1]
                # (add_decomposed_rel_pos :417-421 + get_rel_pos :382-388) and q = W_q x + b_q, so
                # G_h[head][r] = q_head . T_h[r] = (T_h[r] W_q,head) x + T_h[r] . b_q,head is one more linear map of the
                # block input: H * (2 size - 1) rows per axis, appended behind q | k | v; the attention kernel gathers
                # rel_h / rel_w from them. Padded tokens of a window (x = 0) get G = T . b_q, exactly what q = b_q gives.
                wq, bq = wqkv[:C].view(spec.num_heads, d, C), bqkv[:C].view(spec.num_heads, d)
                ext_w, ext_b = [], []
                for nm in ("attn.rel_pos_h", "attn.rel_pos_w"):
                    T = rel_pos_resized(size, sd[p + nm])                             # [L, d]
                    ext_w.append(torch.einsum("rd,hdc->hrc", T, wq).reshape(-1, C))  # [H * L, C]
                    ext_b.append(torch.einsum("rd,hd->hr", T, bq).reshape(-1))
                n_ext = sum(w.shape[0] for w in ext_w)
                pad = (-(3 * C + n_ext)) % 32                                         # GEMM N must be a multiple of 32
                wqkv = torch.cat([wqkv] + ext_w + [torch.zeros(pad, C)], 0)
                bqkv = torch.cat([bqkv] + ext_b + [torch.zeros(pad)], 0)
            blk = dict(
                window=0 if (i in spec.global_attn_indexes or ws == 0) else ws,
                n1=(dev(sd[p + "norm1.weight"], f32), dev(sd[p + "norm1.bias"], f32)),
                n2=(dev(sd[p + "norm2.weight"], f32), dev(sd[p + "norm2.bias"], f32)),
                qkv_w=dev(wqkv, bf), qkv_b=dev(bqkv, f32),
                proj_w=dev(_lora_merged(sd, p + "attn.proj", lora_scale), bf), proj_b=dev(_bias(sd, p + "attn.proj"), f32),
                lin1_w=dev(_lora_merged(sd, p + "mlp.lin1", lora_scale), bf), lin1_b=dev(_bias(sd, p + "mlp.lin1"), f32),
                lin2_w=dev(_lora_merged(sd, p + "mlp.lin2", lora_scale), bf), lin2_b=dev(_bias(sd, p + "mlp.lin2"), f32))
            if C % 256 == 0:   # norm2 -> mlp.lin1 (+ GELU) with the LayerNorm folded into the weights (ops.fold_layernorm)
                wf, bfold, cs = ops.fold_layernorm(_lora_merged(sd, p + "mlp.lin1", lora_scale), _bias(sd, p + "mlp.lin1"),
                                                   sd[p + "norm2.weight"], sd[p + "norm2.bias"])
                blk["lin1_f"] = (dev(wf, bf), dev(bfold, f32), dev(cs, f32))
            self.blocks.append(blk)
        self._keep = keep
        self._maps: Dict[tuple, tuple] = {}
        self._ws = None
        self._c_blocks = (_C.VfmSamBlockParams * spec.depth)()
        for blk, b in zip(self._c_blocks, self.blocks):
            blk.window, blk.qkv_n = b["window"], b["qkv_w"].shape[0]
            blk.ln1_w, blk.ln1_b = (t.data_ptr() for t in b["n1"])
            blk.ln2_w, blk.ln2_b = (t.data_ptr() for t in b["n2"])
            for nm in ("qkv_w", "qkv_b", "proj_w", "proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b"):
                setattr(blk, nm, b[nm].data_ptr())
            if "lin1_f" in b:
                blk.lin1_wf, blk.lin1_bf, blk.lin1_cs = (t.data_ptr() for t in b["lin1_f"])

    def _c_params(self, n: int) -> _C.VfmSamParams:
        """The host-side parameter block of vfm_sam_forward (include/vfmseg_b200.h) for n windows: pointers into the packed
        weights, the window maps / persistent window buffer of this n and the one-hot key matrix of the grid."""
        key = ("cparams", n)
        if key in self._maps:
            return self._maps[key]
        s = self.spec
        outs = sorted(s.out_indices)
        if len(outs) > 8:
            raise ValueError("at most 8 feature taps")
        p = _C.VfmSamParams()
        p.embed_dim, p.depth, p.heads, p.head_dim, p.hidden = s.embed_dim, s.depth, s.num_heads, self.head_dim, s.hidden
        p.n_taps, p.grid, p.use_rel_pos, p.ln_eps = len(outs), s.grid, int(s.use_rel_pos), s.ln_eps
        for i, o in enumerate(outs):
            p.tap_blocks[i] = o
        p.patch_w, p.patch_b = self.patch_w.data_ptr(), self.patch_b.data_ptr()
        p.pos_embed, p.ones = self.pos.data_ptr(), self.ones.data_ptr()
        wsz = {b["window"] for b in self.blocks if b["window"]}
        if len(wsz) > 1:
            raise ValueError("one window size per backbone (sam_vit.py:94-107)")
        if wsz:
            ws = wsz.pop()
            part, unpart, n_win = self._window_maps(n, s.grid, s.grid, ws)
            p.part, p.unpart, p.win_rows = part.data_ptr(), unpart.data_ptr(), n_win * ws * ws
            p.win_buf = self._window_buffer(n_win * ws * ws, s.embed_dim).data_ptr()
        if s.use_rel_pos and self.head_dim == 80 and (s.grid + 15) // 16 * 32 <= 128:
            key1 = ("onehot", s.grid, s.grid)
            if key1 not in self._maps:
                self._maps[key1] = ops.relpos_onehot(s.grid, s.grid, self.device)
            p.onehot, p.onehot_rows = self._maps[key1].data_ptr(), self._maps[key1].shape[0]
        p.blocks = C_.cast(self._c_blocks, C_.POINTER(_C.VfmSamBlockParams))
        self._maps[key] = p
        return p

    def _window_maps(self, n: int, gh: int, gw: int, ws: int):
        """Row maps of window_partition / window_unpartition (sam_vit.py:292-346) for n crops of gh x gw tokens:
        part[i] = token-order row feeding window-order row i (-1: zero padding), unpart[r] = window-order row of token r."""
        key = (n, gh, gw, ws)
        if key not in self._maps:
            Hp, Wp = gh + (ws - gh % ws) % ws, gw + (ws - gw % ws) % ws
            idx = torch.full((n, Hp, Wp), -1, dtype=torch.int64)
            idx[:, :gh, :gw] = torch.arange(n * gh * gw).view(n, gh, gw)
            part = idx.view(n, Hp // ws, ws, Wp // ws, ws).permute(0, 1, 3, 2, 4).reshape(-1)
            unpart = torch.empty(n * gh * gw, dtype=torch.int64)
            real = part >= 0
            unpart[part[real]] = torch.arange(part.numel())[real]
            n_win = n * (Hp // ws) * (Wp // ws)
            self._maps[key] = (part.to(torch.int32).to(self.device), unpart.to(torch.int32).to(self.device), n_win)
        return self._maps[key]

    def _window_buffer(self, rows: int, C: int) -> torch.Tensor:
        """Persistent window-order activation buffer, zero-filled once: the padding rows are never written afterwards."""
        key = ("winbuf", rows, C)
        if key not in self._maps:
            self._maps[key] = torch.zeros(rows, C, dtype=torch.bfloat16, device=self.device)
        return self._maps[key]

    # SlideEngine.backbone_taps dispatches here
    def forward_taps(self, img: torch.Tensor, crops: torch.Tensor, gh: int, gw: int, pixel_norm) -> torch.Tensor:
        """SAMViT.forward (sam_vit.py:123-147) for the listed windows -> taps bf16 [n*gh*gw, n_taps*C] (token-major): the raw
        block outputs at out_indices."""
        s = self.spec
        if gh != s.grid or gw != s.grid:
            raise _C.VfmError(f"SAMViT adds a fixed {s.grid}x{s.grid} pos_embed without interpolation (sam_vit.py:131-132): "
                              f"windows must be {s.grid * 16}x{s.grid * 16}, got grid {gh}x{gw}")
        n, P, C, H, d = crops.shape[0], gh * gw, s.embed_dim, s.num_heads, self.head_dim
        if os.environ.get("VFM_SAM_DRIVER", "c") != "py":
            # one call of the fused C driver; VFM_SAM_DRIVER=py issues the same launch sequence operator by operator from Python
            # (the comparison path of tests/test_sam_gpu.py: bit-identical taps)
            is_u8 = img.dtype == torch.uint8
            if is_u8 and pixel_norm is None:
                raise _C.VfmError("uint8 input needs set_pixel_norm() (SegDataPreProcessor mean/std)")
            if not is_u8 and img.dtype != torch.float32:
                raise _C.VfmError(f"input must be uint8 or float32, got {img.dtype}")
            assert img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
            assert crops.dtype == torch.int32 and crops.is_contiguous() and crops.shape[1] == 4
            prm = self._c_params(n)
            need = _C.load().vfm_sam_workspace_bytes(C_.byref(prm), n)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            taps = torch.empty(n * P, len(s.out_indices) * C, dtype=torch.bfloat16, device=self.device)
            _C.call("vfm_sam_forward", C_.byref(prm), img.data_ptr(), int(is_u8),
                    C_.byref(pixel_norm) if (is_u8 and pixel_norm is not None) else None, img.shape[2], img.shape[3],
                    crops.data_ptr(), n, taps.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                    torch.cuda.current_stream().cuda_stream)
            return taps
        scale = d ** -0.5
        g0 = 3 * C if s.use_rel_pos else -1          # first table-term column of the extended qkv rows
        a = ops.patch_gather(img, crops, gh, gw, pixel_norm if img.dtype == torch.uint8 else None)
        x = ops.gemm_patch_embed_nocls(a, self.patch_w, self.patch_b, self.pos, n, P)
        outs = sorted(s.out_indices)
        taps = torch.empty(n * P, len(outs) * C, dtype=torch.bfloat16, device=x.device)
        for i, b in enumerate(self.blocks):
            tap_i = outs.index(i - 1) if (i - 1) in outs else None                    # tap of the previous block's output
            ws = b["window"]
            tap_args = (taps if tap_i is not None else None, (tap_i or 0) * C)
            if ws:
                # window_partition folded into the LayerNorm store (rows go straight to window order; the padding rows of
                # the persistent buffer stay zero) and window_unpartition into the attention kernel's store
                part, unpart, n_win = self._window_maps(n, gh, gw, ws)
                hw = self._window_buffer(n_win * ws * ws, C)
                ops.layernorm_tap_nocls_into(x, *b["n1"], s.ln_eps, hw, unpart, *tap_args)
                qkv = ops.gemm_bias_bf16(hw, b["qkv_w"], b["qkv_b"])                               # q | k | v | G_h | G_w
                if d == 80 and ws * ws <= 208 and ws <= 16:      # whole window in one tcgen05 score tile
                    att = ops.attention_window_tc_into(qkv, n_win, ws * ws, H, d, ws, ws, scale, g0, part, n * P)
                else:
                    att = ops.rows_gather(ops.attention_relpos_terms(qkv, n_win, ws * ws, H, d, ws, ws, scale, g0), unpart)
            else:
                h = ops.layernorm_tap_nocls(x, *b["n1"], s.ln_eps, *tap_args)
                qkv = ops.gemm_bias_bf16(h, b["qkv_w"], b["qkv_b"])
                bias_cols = (gh + 15) // 16 * 16 + (gw + 15) // 16 * 16
                if d == 80 and g0 >= 0 and bias_cols <= 128:           # key-tile loop on tcgen05
                    key = ("onehot", gh, gw)
                    if key not in self._maps:
                        self._maps[key] = ops.relpos_onehot(gh, gw, self.device)
                    att = ops.attention_global_tc(qkv, self._maps[key], n, P, H, d, gh, gw, scale, g0)
                else:
                    att = ops.attention_relpos_terms(qkv, n, P, H, d, gh, gw, scale, g0)
            if "lin1_f" in b and (_ln_fold_bits() & 2):
                # the proj GEMM also emits bf16(x) and the row statistics; lin1 applies norm2 in its epilogue (no LayerNorm pass)
                xb, st = ops.gemm_bias_ls_residual_stats_(x, att, b["proj_w"], b["proj_b"], self.ones)
                hid = ops.gemm_lnfold_bf16(xb, st, *b["lin1_f"], s.ln_eps, gelu=True)
            else:
                ops.gemm_bias_ls_residual_(x, att, b["proj_w"], b["proj_b"], self.ones)
                hid = ops.gemm_bias_gelu_bf16(ops.layernorm(x, *b["n2"], s.ln_eps), b["lin1_w"], b["lin1_b"])
            ops.gemm_bias_ls_residual_(x, hid, b["lin2_w"], b["lin2_b"], self.ones)
        if (s.depth - 1) in outs:
            ops.layernorm_tap_nocls(x, None, None, s.ln_eps, taps, outs.index(s.depth - 1) * C, want_out=False)
        return taps
