"""CUDA-graph capture of launch-bound static kernel sequences (one page of RCAN is ~800 launches, the YOLO graph ~150,
the SAM encoder ~120).  The kernels are launched through the C ABI on torch's current stream, so they are captured by
`torch.cuda.graph` like any other stream work; replay costs one launch."""
from __future__ import annotations

import os
from typing import Callable

import torch

ENABLED = os.environ.get("MTB200_CUDA_GRAPHS", "1") != "0"
replayed_launches = 0     # kernels of this library launched through graph replays (bench.py's gpu_launches)


class CapturedGraph:
    def __init__(self, fn: Callable[[], None], warmup: int = 1):
        """`fn` must only read/write pre-allocated (static) device buffers."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from ._lib import launch_count
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            fn()
        self.n_kernels = launch_count() - n0          # C-ABI launches recorded into the graph

    def replay(self) -> None:
        global replayed_launches
        self.graph.replay()
        replayed_launches += self.n_kernels
