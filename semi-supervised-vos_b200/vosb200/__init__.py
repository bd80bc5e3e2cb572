"""vosb200: B200-native label-propagation engine (host side of libvosprop.so)."""
from ._capi import (KERNEL_SIMT, KERNEL_TC, KERNEL_TC_DENSE, PREC_BF16, PREC_F16, PREC_SPLIT3,  # noqa: F401
                    VosPropError)
from .engine import PropagationEngine, normalize_frames, plan_refs, precision_for, sample_frames  # noqa: F401
