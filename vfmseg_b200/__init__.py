"""vfmseg_b200 — B200-native (sm_100a) slide-inference hot path of tpy001/VFMSeg.

Importing the package registers the drop-in classes (same `type=` names as the reference's
`rein` plugin) into `vfmseg_b200.registry.MODELS / METRICS`.
"""
__version__ = "0.1.0"

from .registry import BACKBONES, METRICS, MODELS, register_into_mmseg  # noqa: F401
from . import models  # noqa: F401  (registers backbone / head / segmentor)
from .dg_metrics import DGIoUMetric  # noqa: F401

register_into_mmseg()
