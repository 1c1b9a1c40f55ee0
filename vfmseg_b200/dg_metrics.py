"""`DGIoUMetric` — registered metric mirroring rein/dg_metrics.py:24-102 (a subclass of mmseg
IoUMetric there), with the three float32 `torch.histc` passes + four device->host syncs per image
replaced by one integer confusion-matrix kernel and no sync until `compute_metrics`.

Per image the reference stores [dataset_key, area_intersect, area_union, area_pred, area_label]
(float32[19] each, on the CPU). Here the per-image record is [dataset_key, cm] with cm an int64
[(nc+1), nc] DEVICE tensor (cm[label, pred]; row nc collects labels outside [0, nc) that are not
ignored, which the reference still counts in area_pred). The four areas are diag / column sums /
row sums of cm — identical integers to the reference's histograms.
Multi-GPU: `evaluate()` all-reduces the per-key int64 matrices over NCCL (the only collective of
the whole path) instead of mmengine's pickled-object all-gather.
"""
from __future__ import annotations

import os

from collections import OrderedDict, defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .registry import METRICS


def areas_from_confusion(cm: torch.Tensor, num_classes: int):
    """int64 [(nc+1), nc] -> (intersect, union, pred, label) int64 [nc] (mmseg intersect_and_union)."""
    inter = torch.diagonal(cm[:num_classes])
    pred = cm.sum(0)
    label = cm[:num_classes].sum(1)
    return inter, pred + label - inter, pred, label


def total_area_to_metrics(inter, union, pred, label) -> "OrderedDict[str, float]":
    """mmseg IoUMetric.compute_metrics for iou_metrics=['mIoU'], in float32 like the reference:
    aAcc = sum(I)/sum(L), IoU = I/U, Acc = I/L, np.round(np.nanmean(.) * 100, 2)."""
    f = lambda t: torch.as_tensor(t).detach().to("cpu", torch.float32)
    inter, union, label = f(inter), f(union), f(label)
    ret = OrderedDict(aAcc=(inter.sum() / label.sum()).numpy(), IoU=(inter / union).numpy(), Acc=(inter / label).numpy())
    out = OrderedDict()
    for k, v in ret.items():
        out[k if k == "aAcc" else "m" + k] = float(np.round(np.nanmean(v) * 100, 2))
    return out


# mmseg CityscapesDataset.METAINFO['palette'] [3P], the 19 train-id colours
CITYSCAPES_PALETTE = [[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
                      [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [70, 130, 180], [220, 20, 60], [255, 0, 0],
                      [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]]


def id2color(label_pred: np.ndarray) -> np.ndarray:
    """rein/dg_metrics.py:14-22: label map [H,W] -> RGB uint8 [H,W,3]; ids outside the palette stay black."""
    lut = np.zeros((int(max(label_pred.max(), 0)) + 2, 3), dtype=np.uint8)
    n = min(len(CITYSCAPES_PALETTE), lut.shape[0])
    lut[:n] = np.asarray(CITYSCAPES_PALETTE[:n], dtype=np.uint8)
    return lut[np.clip(label_pred, 0, lut.shape[0] - 1)]


@METRICS.register_module()
class DGIoUMetric:
    default_prefix = None

    def __init__(self, dataset_keys=[], mean_used_keys=[], ignore_index: int = 255, iou_metrics=("mIoU",), nan_to_num=None,
                 beta: int = 1, collect_device: str = "cpu", output_dir: Optional[str] = None, format_only: bool = False,
                 prefix: Optional[str] = None, keep_predictions: bool = False, **kwargs):
        if list(iou_metrics) != ["mIoU"]:
            raise NotImplementedError("only iou_metrics=['mIoU'] (what every reference config uses) is implemented")
        self.output_dir = output_dir
        self.format_only = format_only
        if output_dir is not None:
            os.makedirs(output_dir, exist_ok=True)
        self.dataset_keys = list(dataset_keys)
        self.mean_used_keys = list(mean_used_keys) if mean_used_keys else list(dataset_keys)
        self.ignore_index = ignore_index
        self.results: List[list] = []
        self._dataset_meta: Optional[dict] = None
        # keep_predictions: hold the uint8 label maps and gather them to rank 0 in dataset order in evaluate()
        # (mmengine collect_results behind tools/test.py --out); `predictions` is filled on rank 0 only
        self.keep_predictions = keep_predictions
        self._pred_maps: List[torch.Tensor] = []
        self.predictions: Optional[List[torch.Tensor]] = None

    @property
    def dataset_meta(self):
        return self._dataset_meta

    @dataset_meta.setter
    def dataset_meta(self, m):
        self._dataset_meta = m

    # ------------------------------------------------------------------ per batch
    def process(self, data_batch: dict, data_samples: Sequence[dict]) -> None:
        """rein/dg_metrics.py:34-58. `data_samples[i]['pred_sem_seg']['data']` is the label map (uint8 or
        int64, on the GPU), `['gt_sem_seg']['data']` the ground truth."""
        num_classes = len(self.dataset_meta["classes"])
        for data_sample in data_samples:
            pred = data_sample["pred_sem_seg"]["data"].squeeze()
            if self.output_dir is not None:      # dg_metrics.py:60-72: colourised PNG of the prediction
                self._save_png(pred, data_sample)
            if self.keep_predictions:
                self._pred_maps.append(pred.to(torch.uint8))
            if self.format_only:                 # test sets without ground truth (:46-47)
                continue
            label = data_sample["gt_sem_seg"]["data"].squeeze().to(pred.device)
            if pred.dtype != torch.uint8:
                pred = pred.to(torch.uint8)
            if label.dtype != torch.uint8:
                # ignore_index 255 and class ids < 255 survive the narrowing; anything else is not a valid label map
                label = label.to(torch.uint8)
            cm = torch.zeros(num_classes + 1, num_classes, dtype=torch.int64, device=pred.device)
            ops.confusion_matrix_(cm, pred.contiguous().view(-1), label.contiguous().view(-1), num_classes, self.ignore_index)
            dataset_key = "unknown"
            for key in self.dataset_keys:
                if key in data_samples[0]["seg_map_path"]:   # [0], as in the reference (:55)
                    dataset_key = key
                    break
            self.results.append([dataset_key, cm])

    def _save_png(self, pred: torch.Tensor, data_sample: dict) -> None:
        from PIL import Image
        basename = os.path.splitext(os.path.basename(data_sample["img_path"]))[0]
        mask = pred.detach().cpu().numpy().astype(np.int64)
        if data_sample.get("reduce_zero_label", False):
            mask = mask + 1
        Image.fromarray(id2color(mask)).save(os.path.abspath(os.path.join(self.output_dir, f"{basename}.png")))

    # ------------------------------------------------------------------ reduction
    @staticmethod
    def _areas_of(entry, num_classes):
        """Accepts our [key, cm] records and the reference's [key, I, U, P, L] records."""
        if len(entry) == 2:
            return [a for a in areas_from_confusion(entry[1], num_classes)]
        return [torch.as_tensor(a) for a in entry[1:]]

    def compute_metrics(self, results: list) -> Dict[str, float]:
        """rein/dg_metrics.py:74-102: group by dataset key, sum, IoU/Acc per key, mean over mean_used_keys."""
        num_classes = len(self.dataset_meta["classes"]) if self.dataset_meta else None
        grouped = defaultdict(list)
        for r in results:
            grouped[r[0]].append(r)
        metrics: Dict[str, float] = {}
        type2mean = defaultdict(list)
        for key, rs in grouped.items():
            nc = num_classes or rs[0][1].shape[-1]
            if all(len(r) == 2 for r in rs):
                cm = torch.stack([r[1] for r in rs]).sum(0)     # int64, exact
                areas = areas_from_confusion(cm, nc)
            else:
                per = [self._areas_of(r, nc) for r in rs]
                areas = [sum(p[i].to("cpu") for p in per) for i in range(4)]
            for k, v in total_area_to_metrics(*areas).items():
                metrics[f"{key}_{k}"] = v
                if key in self.mean_used_keys:
                    type2mean[k].append(v)
        for k, v in type2mean.items():
            metrics[f"mean_{k}"] = sum(v) / len(v)
        return metrics

    def evaluate(self, size: Optional[int] = None) -> Dict[str, float]:
        """Evaluator entry: all-reduce the per-key confusion matrices across ranks (int64 sum over NCCL),
        compute on every rank, clear. Replaces mmengine collect_results + broadcast_object_list.

        `size` = len(dataset), as mmengine passes it: the round-up sampler hands the last ranks wrapped-around
        duplicates when len(dataset) % world_size != 0 and `collect_results(..., size)` drops them; here every rank
        drops its own trailing duplicates before the all-reduce (collect.num_real_samples), so the multi-GPU metrics
        equal the single-GPU ones."""
        import torch.distributed as dist
        from .collect import gather_label_maps, num_real_samples
        results = self.results
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.keep_predictions:
            self.predictions = gather_label_maps(self._pred_maps, size)
            self._pred_maps = []
        if multi and size is not None and not self.format_only:
            results = results[:num_real_samples(size, dist.get_rank(), dist.get_world_size(), len(results))]
        if multi:
            nc = len(self.dataset_meta["classes"])
            keys = list(self.dataset_keys) + ["unknown"]
            dev = results[0][1].device if results else torch.device("cuda", torch.cuda.current_device())
            if dist.get_backend() == "gloo":
                dev = torch.device("cpu")
            stacked = torch.zeros(len(keys), nc + 1, nc, dtype=torch.int64, device=dev)
            for r in results:
                stacked[keys.index(r[0])] += r[1].to(dev)
            dist.all_reduce(stacked, op=dist.ReduceOp.SUM)
            results = [[k, stacked[i]] for i, k in enumerate(keys) if int(stacked[i].sum()) > 0]
        out = self.compute_metrics(results)
        self.results = []
        return out
