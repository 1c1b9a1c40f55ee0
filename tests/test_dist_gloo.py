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


# ---------------------------------------------------------------- len(dataset) % world_size != 0 + prediction-map gather
def _images(n, nc=19):
    """n synthetic (pred, gt) pairs; image 3 is smaller (a second dataset with another resolution)."""
    rng = np.random.default_rng(7)
    out = []
    for i in range(n):
        shape = (24, 40) if i == 3 else (32, 48)
        pred = rng.integers(0, nc, shape).astype(np.uint8)
        gt = rng.integers(0, nc + 1, shape)
        gt[gt == nc] = 255
        out.append((pred, gt.astype(np.uint8)))
    return out


def _worker_uneven(rank, world, port, q, n_images):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vfmseg_b200.dg_metrics import DGIoUMetric
    nc = 19
    m = DGIoUMetric(dataset_keys=["citys", "bdd"], ignore_index=255, keep_predictions=True)
    m.dataset_meta = dict(classes=list(range(nc)))
    imgs = _images(n_images)
    # mmengine DefaultSampler(round_up=True): indices wrap around to a multiple of the world size, rank r takes r::world
    total = -(-n_images // world) * world
    padded = (list(range(n_images)) * (total // n_images + 1))[:total]
    for g in padded[rank::world]:
        pred, gt = imgs[g]
        m.results.append(["citys" if g < 3 else "bdd", torch.from_numpy(torch_ref.confusion_matrix_np(pred, gt, nc, 255))])
        m._pred_maps.append(torch.from_numpy(pred))
    out = m.evaluate(size=n_images)
    preds = None if m.predictions is None else [p.numpy() for p in m.predictions]
    q.put((rank, out, preds))
    dist.destroy_process_group()


def test_uneven_split_drops_padding_duplicates_and_gathers_maps_in_dataset_order():
    """5 images on 2 ranks: rank 1's third sample is the wrapped-around image 0 (it used to be counted twice).
    Metrics must equal the single-process ones; the gathered label maps must be images 0..4 in order, ragged shapes kept."""
    world, n_images = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_uneven, args=(r, world, port, q, n_images)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (o, pr) for r, o, pr in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from vfmseg_b200.dg_metrics import DGIoUMetric
    nc = 19
    m = DGIoUMetric(dataset_keys=["citys", "bdd"], ignore_index=255)
    m.dataset_meta = dict(classes=list(range(nc)))
    imgs = _images(n_images)
    for g, (pred, gt) in enumerate(imgs):
        m.results.append(["citys" if g < 3 else "bdd", torch.from_numpy(torch_ref.confusion_matrix_np(pred, gt, nc, 255))])
    want = m.evaluate(size=n_images)
    assert got[0][0] == got[1][0] == want
    assert got[1][1] is None and len(got[0][1]) == n_images
    for g in range(n_images):
        assert got[0][1][g].shape == imgs[g][0].shape and (got[0][1][g] == imgs[g][0]).all()


def test_num_real_samples():
    from vfmseg_b200.collect import num_real_samples
    # 500 Cityscapes val images on 8 GPUs: 63 per rank, ranks 4..7 carry one duplicate each
    assert [num_real_samples(500, r, 8, 63) for r in range(8)] == [63, 63, 63, 63, 62, 62, 62, 62]
    assert num_real_samples(None, 3, 8, 63) == 63 and num_real_samples(2, 5, 8, 1) == 0
