"""Global constants, same names as the reference's src/config.py:10-14 (a mutable class used as a
namespace; `inference_command_impl` overrides DEVICE from --device)."""
import multiprocessing

import torch


class Config(object):
    DEVICE = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    SCALE = 0.125            # feature stride 8
    CONTINUOUS_FRAME = 4     # 3 most recent frames are always references (+1)
    CPU_COUNT = max(multiprocessing.cpu_count(), 1)
