import sys; sys.path.insert(0, "/root/repo")
import torch
from vfmseg_b200 import ops
from vfmseg_b200.engine import slide_boxes
bx = torch.tensor(slide_boxes(1024, 2048, (512, 512), (341, 341)), dtype=torch.int32).cuda()
low = torch.randn(36, 19, 128, 128, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for want in (False, True):
    f = lambda: ops.slide_merge_argmax(low, bx, 2, (512, 512), (1024, 2048), want_logits=want)
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print("want_logits", want, "median us", sorted(ts)[5] * 1e3)
