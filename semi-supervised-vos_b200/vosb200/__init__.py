"""vosb200: B200-native label-propagation engine (host side of libvosprop.so)."""
from ._capi import KERNEL_SIMT, KERNEL_TC, KERNEL_TC_DENSE, VosPropError  # noqa: F401
from .engine import PropagationEngine, plan_refs, sample_frames  # noqa: F401
