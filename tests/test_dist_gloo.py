"""CPU, world_size 2 over gloo: the multi-GPU reduction of the path — per-rank confusion matrices summed
with one all-reduce (DGIoUMetric.evaluate) — equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vfmseg_b200.dg_metrics import DGIoUMetric
    nc = 19
    m = DGIoUMetric(dataset_keys=["citys", "bdd"], ignore_index=255)
    m.dataset_meta = dict(classes=list(range(nc)))
    rng = np.random.default_rng(100)
    for i in range(6):                       # images round-robin over ranks, like mmengine DefaultSampler
        pred = rng.integers(0, nc, (32, 48))
        gt = rng.integers(0, nc + 1, (32, 48))
        gt[gt == nc] = 255
        if i % world != rank:
            continue
        cm = torch.from_numpy(torch_ref.confusion_matrix_np(pred, gt, nc, 255))
        m.results.append(["citys" if i < 4 else "bdd", cm])
    out = m.evaluate()
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_confusion_allreduce_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    from vfmseg_b200.dg_metrics import DGIoUMetric
    nc = 19
    m = DGIoUMetric(dataset_keys=["citys", "bdd"], ignore_index=255)
    m.dataset_meta = dict(classes=list(range(nc)))
    rng = np.random.default_rng(100)
    for i in range(6):
        pred = rng.integers(0, nc, (32, 48))
        gt = rng.integers(0, nc + 1, (32, 48))
        gt[gt == nc] = 255
        m.results.append(["citys" if i < 4 else "bdd", torch.from_numpy(torch_ref.confusion_matrix_np(pred, gt, nc, 255))])
    want = m.evaluate()
    assert outs[0] == outs[1] == want
    assert set(want) >= {"citys_mIoU", "bdd_mIoU", "mean_mIoU", "mean_aAcc"}
