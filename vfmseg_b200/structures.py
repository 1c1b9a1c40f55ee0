"""Minimal stand-ins for mmseg.structures.SegDataSample / mmengine PixelData: just enough surface
for `sample.pred_sem_seg.data`, `sample.seg_logits.data`, `sample.gt_sem_seg.data`, `.metainfo`
and dict-style access used by rein/dg_metrics.py:47-55."""
from __future__ import annotations


class PixelData:
    def __init__(self, data=None):
        self.data = data

    def __getitem__(self, k):
        if k == "data":
            return self.data
        raise KeyError(k)


class SegDataSample:
    def __init__(self, metainfo=None, **fields):
        self.metainfo = dict(metainfo or {})
        for k, v in fields.items():
            setattr(self, k, v)

    def set_data(self, d):
        for k, v in d.items():
            setattr(self, k, v)

    def set_metainfo(self, d):
        self.metainfo.update(d)

    # evaluator-side dict view (mmengine calls .to_dict() before Metric.process)
    def to_dict(self):
        out = dict(self.metainfo)
        for k, v in self.__dict__.items():
            if isinstance(v, PixelData):
                out[k] = {"data": v.data}
        return out

    def __getitem__(self, k):
        if k in self.metainfo:
            return self.metainfo[k]
        v = getattr(self, k)
        return v

    def get(self, k, default=None):
        try:
            return self[k]
        except (KeyError, AttributeError):
            return default
