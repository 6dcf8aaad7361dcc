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


def get_device_info(device: Optional[torch.device] = None) -> dict:
    """The reference's report format (core/device.py:116-172): device name + allocated / reserved GB as 2-decimal strings
    on CUDA, {"device": "cpu", "memory": "N/A"} otherwise."""
    device = device or get_best_device()
    if device.type == "cuda" and torch.cuda.is_available():
        return {"device": torch.cuda.get_device_name(0),
                "allocated_gb": f"{torch.cuda.memory_allocated() / 1024 ** 3:.2f}",
                "reserved_gb": f"{torch.cuda.memory_reserved() / 1024 ** 3:.2f}"}
    return {"device": "cpu", "memory": "N/A"}


def is_gpu_available() -> bool:
    return torch.cuda.is_available()
