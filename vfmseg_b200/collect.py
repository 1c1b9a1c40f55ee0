"""Multi-GPU collection of per-image results in dataset order — the B200 counterpart of mmengine
`collect_results` (used by `IoUMetric.evaluate(size)` and by `tools/test.py --out`, /root/reference/tools/test.py:137-139;
the PNG dump of rein/dg_metrics.py:60-72 consumes the same maps).

mmengine's `DefaultSampler(round_up=True)` pads the dataset to a multiple of the world size by wrapping around and
deals index g to rank g % W as its (g // W)-th sample; `collect_results` zips the ranks' lists back together and
truncates to `size`, which drops exactly the wrapped duplicates. Here that is two functions:

* `num_real_samples(size, rank, world)` — how many of a rank's results are real (not padding duplicates);
* `gather_label_maps(maps, size)` — uint8 label maps (2 MiB per 1024x2048 image) gathered to rank 0 with one
  `torch.distributed.gather` per chunk (NCCL over NVLink on GPUs, gloo on CPU), ragged shapes padded to the chunk maximum.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def num_real_samples(size: Optional[int], rank: int, world: int, n_local: int) -> int:
    """Results of `rank` that survive mmengine's `collect_results(..., size)`: its j-th result is dataset index
    rank + j * world and is kept iff that index < size."""
    if size is None:
        return n_local
    return max(0, min(n_local, -(-(size - rank) // world)))


def gather_label_maps(maps: Sequence[torch.Tensor], size: Optional[int] = None, chunk: int = 8,
                      dst: int = 0) -> Optional[List[torch.Tensor]]:
    """maps: this rank's uint8 [H, W] label maps in the order it processed them (every rank holds the same number,
    as the round-up sampler guarantees). Returns on rank `dst` the list of all maps in dataset order truncated to
    `size` (CPU tensors), None on the other ranks. Single process: the maps themselves (on the CPU)."""
    maps = [m.squeeze() for m in maps]
    for m in maps:
        if m.dtype != torch.uint8 or m.dim() != 2:
            raise ValueError("gather_label_maps expects uint8 [H, W] label maps")
    if not _dist_on():
        out = [m.cpu() for m in maps]
        return out if size is None else out[:size]
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(maps)
    dev = maps[0].device if maps else torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    if dist.get_backend() == "gloo":
        dev = torch.device("cpu")
    # every rank must hold the same number of maps (DefaultSampler round_up); shapes may differ (mixed datasets)
    meta = torch.tensor([n, -n], dtype=torch.int64, device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX)
    if int(meta[0]) != -int(meta[1]):
        raise RuntimeError(f"gather_label_maps: ranks hold different numbers of maps ({-int(meta[1])}..{int(meta[0])})")
    shapes = torch.tensor([[m.shape[0], m.shape[1]] for m in maps], dtype=torch.int64, device=dev).view(n, 2)
    all_shapes = [torch.empty_like(shapes) for _ in range(world)]
    dist.all_gather(all_shapes, shapes)
    all_shapes = torch.stack(all_shapes).cpu()            # [world, n, 2]
    out: List[Optional[torch.Tensor]] = [None] * (n * world) if rank == dst else []
    for j0 in range(0, n, chunk):
        j1 = min(n, j0 + chunk)
        hmax = int(all_shapes[:, j0:j1, 0].max())
        wmax = int(all_shapes[:, j0:j1, 1].max())
        buf = torch.zeros(j1 - j0, hmax, wmax, dtype=torch.uint8, device=dev)
        for j in range(j0, j1):
            m = maps[j].to(dev)
            buf[j - j0, :m.shape[0], :m.shape[1]] = m
        recv = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, recv, dst=dst)
        if rank == dst:
            for r in range(world):
                host = recv[r].cpu()
                for j in range(j0, j1):
                    h, w = int(all_shapes[r, j, 0]), int(all_shapes[r, j, 1])
                    out[j * world + r] = host[j - j0, :h, :w].clone()
    if rank != dst:
        return None
    return out if size is None else out[:size]
