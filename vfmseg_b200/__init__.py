"""vfmseg_b200 — B200-native (sm_100a) slide-inference hot path of tpy001/VFMSeg."""
__version__ = "0.1.0"
