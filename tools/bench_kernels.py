"""Per-kernel timing at the config-2 shapes (one 1024x2048 image = 18 crops, M = 18450 tokens).

CUDA events on the current stream, warm-up, L2 flushed between iterations. Prints one JSON line per
kernel with achieved TFLOP/s or GB/s against MEASURED_PEAKS.json. Development tool; bench.py is the
contract benchmark.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import ops  # noqa: E402


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0)
    return 1590.0, 6650.0


ONLY = ""
WARMUP = 3


def timeit(fn, iters=10, warmup=None, flush=None, label=""):
    if ONLY and ONLY not in label:
        return None, None
    warmup = WARMUP if warmup is None else warmup
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
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
    ap.add_argument("--crops", type=int, default=18)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="substring filter on the kernel label (profiling aid)")
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    global ONLY, WARMUP
    ONLY, WARMUP = args.only, args.warmup
    tf_peak, gb_peak = peaks()
    dev = "cuda"
    T, C, H, heads = 1025, 1024, 4096, 16
    M = args.crops * T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
    f32 = lambda *s: torch.randn(*s, device=dev)
    res = []

    def rec(name, ms, best, flops=None, bytes_=None):
        if ms is None:
            return
        d = {"kernel": name, "ms_median": round(ms, 4), "ms_best": round(best, 4)}
        if flops:
            d["tflops"] = round(flops / ms / 1e9, 1)
            d["frac_of_measured_bf16_peak"] = round(flops / ms / 1e9 / tf_peak, 3)
        if bytes_:
            d["gbs"] = round(bytes_ / ms / 1e6, 1)
            d["frac_of_measured_hbm_peak"] = round(bytes_ / ms / 1e6 / gb_peak, 3)
        print(json.dumps(d), flush=True)
        res.append(d)

    a = bf(M, C); a4 = bf(M, H)
    wqkv = bf(3 * C, C); bq = f32(3 * C)
    wproj = bf(C, C); bp = f32(C); g = f32(C)
    w1 = bf(H, C); b1 = f32(H); w2 = bf(C, H)
    x = f32(M, C)
    ms, best = timeit(lambda: ops.gemm_bias_bf16(a, wqkv, bq), args.iters, flush=flush, label="gemm qkv  [M,1024]x[3072,1024]")
    rec("gemm qkv  [M,1024]x[3072,1024]", ms, best, 2.0 * M * 3 * C * C)
    ms, best = timeit(lambda: ops.gemm_bias_ls_residual_(x, a, wproj, bp, g), args.iters, flush=flush, label="gemm proj [M,1024]x[1024,1024] +residual")
    rec("gemm proj [M,1024]x[1024,1024] +residual", ms, best, 2.0 * M * C * C)
    ms, best = timeit(lambda: ops.gemm_bias_gelu_bf16(a, w1, b1), args.iters, flush=flush, label="gemm fc1  [M,1024]x[4096,1024] +gelu")
    rec("gemm fc1  [M,1024]x[4096,1024] +gelu", ms, best, 2.0 * M * H * C)
    ms, best = timeit(lambda: ops.gemm_bias_ls_residual_(x, a4, w2, bp, g), args.iters, flush=flush, label="gemm fc2  [M,4096]x[1024,4096] +residual")
    rec("gemm fc2  [M,4096]x[1024,4096] +residual", ms, best, 2.0 * M * H * C)
    # folded-LayerNorm schedule: residual GEMMs that also emit bf16(x) + row statistics, and the GEMMs that consume them
    ms, best = timeit(lambda: ops.gemm_bias_ls_residual_stats_(x, a, wproj, bp, g), args.iters, flush=flush, label="gemm proj +residual +stats")
    rec("gemm proj +residual +stats", ms, best, 2.0 * M * C * C)
    ms, best = timeit(lambda: ops.gemm_bias_ls_residual_stats_(x, a4, w2, bp, g), args.iters, flush=flush, label="gemm fc2  +residual +stats")
    rec("gemm fc2  +residual +stats", ms, best, 2.0 * M * H * C)
    if not ONLY or "lnfold" in ONLY or "stats" in ONLY:
        xb, st = ops.gemm_bias_ls_residual_stats_(x, a, wproj, bp, g)
        cq, c1 = f32(3 * C), f32(H)
        ms, best = timeit(lambda: ops.gemm_lnfold_bf16(xb, st, wqkv, bq, cq, 1e-6), args.iters, flush=flush, label="gemm qkv  lnfold")
        rec("gemm qkv  lnfold", ms, best, 2.0 * M * 3 * C * C)
        ms, best = timeit(lambda: ops.gemm_lnfold_bf16(xb, st, w1, b1, c1, 1e-6, gelu=True), args.iters, flush=flush, label="gemm fc1  lnfold +gelu")
        rec("gemm fc1  lnfold +gelu", ms, best, 2.0 * M * H * C)
    qkv = bf(M, 3 * C)
    ms, best = timeit(lambda: ops.attention_fwd(qkv, args.crops, T, heads), args.iters, flush=flush, label="attention 16 heads S=1025 d=64")
    rec("attention 16 heads S=1025 d=64", ms, best, 4.0 * args.crops * heads * T * T * 64)
    lw, lb = f32(C), f32(C)
    ms, best = timeit(lambda: ops.layernorm(x, lw, lb, 1e-6), args.iters, flush=flush, label="layernorm [M,1024] f32->bf16")
    rec("layernorm [M,1024] f32->bf16", ms, best, None, M * C * 6.0)
    # tail
    boxes = []
    for y1 in (0, 341, 512):
        for x1 in (0, 341, 682, 1023, 1364, 1536):
            boxes.append((y1, x1))
    bt = torch.tensor(boxes, dtype=torch.int32, device=dev)
    low = f32(18, 19, 128, 128)
    ms, best = timeit(lambda: ops.slide_merge_argmax(low, bt, 1, (512, 512), (1024, 2048)), args.iters, flush=flush, label="slide_merge_argmax 1024x2048")
    rec("slide_merge_argmax 1024x2048", ms, best, None, low.numel() * 4.0 + 1024 * 2048)
    pred = torch.randint(0, 19, (1024 * 2048,), device=dev, dtype=torch.uint8)
    lab = torch.randint(0, 19, (1024 * 2048,), device=dev, dtype=torch.uint8)
    cm = torch.zeros(20, 19, dtype=torch.int64, device=dev)
    ms, best = timeit(lambda: ops.confusion_matrix_(cm, pred, lab, 19), args.iters, flush=flush, label="confusion_matrix 2M px (random labels)")
    rec("confusion_matrix 2M px (random labels)", ms, best, None, 2.0 * 1024 * 2048)
    la, lb_ = f32(1, 19, 1024, 2048), f32(1, 19, 1024, 2048)
    ms, best = timeit(lambda: ops.tta_flip_mean_argmax(la, lb_, want_logits=True), args.iters, flush=flush, label="tta_flip_mean_argmax 1024x2048 (+logits)")
    rec("tta_flip_mean_argmax 1024x2048 (+logits)", ms, best, None, 3.0 * la.numel() * 4 + 1024 * 2048)
    ms, best = timeit(lambda: ops.tta_flip_mean_argmax(la, lb_), args.iters, flush=flush, label="tta_flip_mean_argmax 1024x2048 (labels)")
    rec("tta_flip_mean_argmax 1024x2048 (labels)", ms, best, None, 2.0 * la.numel() * 4 + 1024 * 2048)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bench_kernels.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
