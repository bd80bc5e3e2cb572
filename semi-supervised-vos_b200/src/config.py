"""Global constants, same names as the reference's src/config.py:10-14 (a mutable class used as a
namespace; `inference_command_impl` overrides DEVICE from --device)."""
import os

import torch


class Config:
    # the propagation engine is CUDA-only; 'cpu' here only lets the host-side tests import the package
    DEVICE = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
    SCALE = 1.0 / 8.0                       # feature stride 8
    CONTINUOUS_FRAME = 4                    # the 3 most recent frames are always references (+1)
    CPU_COUNT = max(os.cpu_count() or 1, 1)  # evaluation's process pool
