"""Double-buffered host <-> device pipeline around the segmentor's public call.

`predict_labels` wants its images in HBM and leaves the label maps there. A test loop that feeds pinned host images and
consumes host label maps (mmengine's test loop behind tools/test.py: DataLoader -> data_preprocessor -> predict -> metric)
pays the H2D copy of step i+1's input (6.3 MB per 1024x2048 uint8 image) and the D2H copy of step i's label map (2.1 MB)
serially around the compute unless they are moved to side streams. `HostPipeline` does that with two device input slots and
two pinned host output slots: while step i computes on the current stream, the copy-in stream uploads step i+1 and the
copy-out stream downloads step i-1's labels and confusion matrix. Results come back one step late (`submit` returns the
previous step's host tensors, `drain` the last one), which is how a throughput loop consumes them anyway.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class HostPipeline:
    def __init__(self, model, images_like: torch.Tensor, gt_like: Optional[torch.Tensor] = None, num_classes: int = 19,
                 ignore_index: int = 255):
        dev = next(model.parameters()).device
        self.model, self.dev = model, dev
        self.eng = model.engine()
        self.ignore_index = ignore_index
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        B, _, H, W = images_like.shape
        self.in_dev = [torch.empty(images_like.shape, dtype=images_like.dtype, device=dev) for _ in range(2)]
        self.gt_dev = [torch.empty(B, H, W, dtype=torch.uint8, device=dev) for _ in range(2)] if gt_like is not None else None
        self.lab_host = [torch.empty(B, H, W, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.cm_host = [torch.empty(num_classes + 1, num_classes, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.cm = torch.zeros(num_classes + 1, num_classes, dtype=torch.int64, device=dev)
        self.in_ready = [torch.cuda.Event() for _ in range(2)]
        self.slot_free = [torch.cuda.Event() for _ in range(2)]     # compute that read input slot s has finished
        self.out_done = [torch.cuda.Event() for _ in range(2)]
        self.n = 0            # steps submitted
        self._staged = -1     # highest step whose upload has been issued

    def _upload(self, step: int, images: torch.Tensor, gt: Optional[torch.Tensor]):
        s = step % 2
        with torch.cuda.stream(self.copy_in):
            if step >= 2:
                self.copy_in.wait_event(self.slot_free[s])
            self.in_dev[s].copy_(images, non_blocking=True)
            if gt is not None and self.gt_dev is not None:
                self.gt_dev[s].copy_(gt, non_blocking=True)
            self.in_ready[s].record(self.copy_in)
        self._staged = step

    def submit(self, images: torch.Tensor, gt: Optional[torch.Tensor] = None, next_images: Optional[torch.Tensor] = None,
               next_gt: Optional[torch.Tensor] = None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        """Run one step on pinned host `images` (and `gt`); `next_images` lets the upload of the following step overlap this
        step's compute. Returns the PREVIOUS step's (labels, confusion matrix) host tensors, or None on the first call."""
        i, s = self.n, self.n % 2
        cur = torch.cuda.current_stream(self.dev)
        if self._staged < i:
            self._upload(i, images, gt)
        cur.wait_event(self.in_ready[s])
        if i >= 2:
            self.out_done[s].synchronize()                # host slot s is about to be rewritten: its consumer had a full step
        labels, _ = self.model.predict_labels(self.in_dev[s])
        if gt is not None:
            self.eng.confusion(self.cm, labels, self.gt_dev[s], self.ignore_index)
        self.slot_free[s].record(cur)
        if next_images is not None:
            self._upload(i + 1, next_images, next_gt)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.slot_free[s])
            labels.record_stream(self.copy_out)
            self.lab_host[s].copy_(labels, non_blocking=True)
            self.cm_host[s].copy_(self.cm, non_blocking=True)
            self.out_done[s].record(self.copy_out)
        self.n += 1
        if i == 0:
            return None
        p = (i - 1) % 2
        self.out_done[p].synchronize()
        return self.lab_host[p], self.cm_host[p]

    def drain(self) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        if self.n == 0:
            return None
        p = (self.n - 1) % 2
        self.out_done[p].synchronize()
        return self.lab_host[p], self.cm_host[p]
