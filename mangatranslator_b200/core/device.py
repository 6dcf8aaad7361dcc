"""Device helpers with the reference's names (core/device.py:7-225).  This build targets B200 only: the device is
CUDA and the library refuses to run without it (no CPU/MPS/XPU path)."""
import gc
from typing import Optional

import torch


def get_best_device() -> torch.device:
    if not torch.cuda.is_available():
        # the reference would fall back to CPU here (core/device.py:31); this build has no CPU path
        return torch.device("cpu")
    return torch.device("cuda", torch.cuda.current_device())


def get_best_dtype(device: Optional[torch.device] = None) -> torch.dtype:
    device = device or get_best_device()
    return torch.bfloat16 if device.type == "cuda" else torch.float32


def empty_cache(device: Optional[torch.device] = None) -> None:
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def synchronize(device: Optional[torch.device] = None) -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
