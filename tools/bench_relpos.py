"""Kernel-only timing of vfm_attention_relpos_ex on the two SAM ViT-H shapes of a 18-crop pass (config 5 at crop 512):
windowed (162 windows x 196 tokens, 16 heads x 80, table terms behind q|k|v) and global (18 x 1024 tokens). Dev tool."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vfmseg_b200 import ops
H, d = 16, 80
C = H * d
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(name, n_seq, k):
    S = k * k
    ld = 3 * C + 2 * H * (2 * k - 1)
    ld += (-ld) % 32
    qkv = (torch.randn(n_seq * S, ld, device="cuda") * 0.7).to(torch.bfloat16)
    for with_bias in (True, False):
        f = (lambda: ops.attention_relpos_terms(qkv, n_seq, S, H, d, k, k, d ** -0.5, 3 * C)) if with_bias else \
            (lambda: ops.attention_relpos_terms(qkv, n_seq, S, H, d, k, k, d ** -0.5, -1))
        for _ in range(3): f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ms = sorted(ts)[5]
        flop = 4 * n_seq * H * S * S * d
        print(f"{name} bias={with_bias}: {ms*1e3:.1f} us, {flop/ms/1e9:.1f} TFLOP/s (algorithmic), qkv buffer {qkv.numel()*2/1e6:.0f} MB")
run("windowed 162 x 196", 162, 14)
def run_tc(n_seq, k):
    S = k * k
    ld = 3 * C + 2 * H * (2 * k - 1)
    ld += (-ld) % 32
    qkv = (torch.randn(n_seq * S, ld, device="cuda") * 0.7).to(torch.bfloat16)
    for g0 in (3 * C, -1):
        f = lambda: ops.attention_window_tc(qkv, n_seq, S, H, d, k, k, d ** -0.5, g0)
        for _ in range(3): f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ms = sorted(ts)[5]
        print(f"windowed tcgen05 bias={g0 >= 0}: {ms*1e3:.1f} us, {4 * n_seq * H * S * S * d/ms/1e9:.1f} TFLOP/s (algorithmic)")
run_tc(162, 14)
run("global 18 x 1024", 18, 32)

def run_glob(n_seq, k):
    S = k * k
    ld = 3 * C + 2 * H * (2 * k - 1)
    ld += (-ld) % 32
    qkv = (torch.randn(n_seq * S, ld, device="cuda") * 0.7).to(torch.bfloat16)
    e = ops.relpos_onehot(k, k, "cuda")
    f = lambda: ops.attention_global_tc(qkv, e, n_seq, S, H, d, k, k, d ** -0.5, 3 * C)
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        s, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); s.record(); f(); e2.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e2))
    ms = sorted(ts)[5]
    print(f"global tcgen05 {n_seq} x {S}: {ms*1e3:.1f} us, {4 * n_seq * H * S * S * d/ms/1e9:.1f} TFLOP/s (algorithmic)")
run_glob(18, 32)
run_glob(6, 64)
run("global 6 x 4096", 6, 64)
