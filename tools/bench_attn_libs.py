"""Library comparator for the attention core at the config-2 shape: [windows, 1025 tokens, 16 heads, 64] bf16.

Times torch SDPA (cuDNN / flash / efficient backends, each forced) and flash_attn_func on the same box as
the hand-written kernels ("hand-written only counts if it wins"). CUDA events, warm-up,
L2 flushed between iterations, inputs in the layout each library prefers (no transposes inside the timed
region). Prints one JSON line per library plus the repo's own kernel through the C ABI.
"""
import argparse
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=36)
    ap.add_argument("--tokens", type=int, default=1025)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    B, S, H, D = args.windows, args.tokens, args.heads, 64
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B, S, 3, H, D, device=dev, dtype=torch.bfloat16, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flops = 4.0 * B * H * S * S * D
    out = []

    def report(name, fn):
        try:
            med, best = timeit(fn, args.iters, flush)
            rec = {"kernel": name, "shape": [B, S, H, D], "ms_median": round(med, 4), "ms_best": round(best, 4),
                   "tflops": round(flops / med / 1e9, 1)}
        except Exception as e:  # a backend that refuses the shape is a result too
            rec = {"kernel": name, "error": str(e).splitlines()[0][:200]}
        print(json.dumps(rec), flush=True)
        out.append(rec)

    # torch SDPA wants [B, H, S, D]; give it contiguous tensors so no copy is timed
    q, k, v = (qkv[:, :, i].transpose(1, 2).contiguous() for i in range(3))
    from torch.nn.attention import SDPBackend, sdpa_kernel
    for name, be in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION),
                     ("sdpa_efficient", SDPBackend.EFFICIENT_ATTENTION)):
        def fn(be=be):
            with sdpa_kernel(be):
                return F.scaled_dot_product_attention(q, k, v)
        report(name, fn)
    try:
        from flash_attn import flash_attn_func, flash_attn_qkvpacked_func
        qf, kf, vf = (qkv[:, :, i].contiguous() for i in range(3))
        report("flash_attn_func", lambda: flash_attn_func(qf, kf, vf))
        report("flash_attn_qkvpacked", lambda: flash_attn_qkvpacked_func(qkv))
    except Exception as e:
        print(json.dumps({"kernel": "flash_attn", "error": str(e)[:200]}))
    try:
        import flashinfer
        qf, kf, vf = (qkv[:, :, i].reshape(B * S, H, D).contiguous() for i in range(3))
        # ragged batch prefill, every sequence the same length
        ws = torch.empty(128 << 20, dtype=torch.uint8, device=dev)
        indptr = torch.arange(0, (B + 1) * S, S, dtype=torch.int32, device=dev)
        for backend in ("auto", "cutlass"):
            try:
                wr = flashinfer.BatchPrefillWithRaggedKVCacheWrapper(ws, "NHD", backend=backend)
                wr.plan(indptr, indptr, H, H, D, causal=False, q_data_type=torch.bfloat16)
                report(f"flashinfer_ragged_{backend}", lambda: wr.run(qf, kf, vf))
            except Exception as e:
                print(json.dumps({"kernel": f"flashinfer_ragged_{backend}", "error": str(e).splitlines()[0][:200]}))
    except Exception as e:
        print(json.dumps({"kernel": "flashinfer", "error": str(e)[:200]}))

    # the repo's own kernel, through the C ABI, on the packed qkv GEMM output layout [B*S, 3*H*D]
    from vfmseg_b200 import ops
    packed = qkv.reshape(B * S, 3 * H * D).contiguous()
    for mode in (1, 5):
        report(f"vfm_attention mode {mode}", lambda mode=mode: ops.attention_fwd(packed, B, S, H, mode=mode))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "attn_libs.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
