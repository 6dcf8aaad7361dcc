"""Batch coordination — the reference's request-budget helpers (core/batch_coordinator.py:18-164, same names and
behaviour) plus what this build ADDS under the same file name: sharding whole pages across the GPUs of a node.

Pages are independent on the hot path, so there is no data-path collective: rank r of R takes the pages
i = r (mod R) of the naturally sorted list (the order core/pipeline.py:2548-2557 establishes), every rank holds a full
weight replica, and torch.distributed (NCCL over NVLink on the GPU boxes, gloo in the CPU tests) is used only for the
barrier around the timed region and the gather of per-rank results / failure lists to rank 0.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from contextlib import contextmanager
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple, TypeVar

import numpy as np
from PIL import Image

from mangatranslator_b200.utils.exceptions import CancellationError

T = TypeVar("T")
R = TypeVar("R")
BBox = Tuple[int, int, int, int]


class BatchRequestCoordinator:
    """Shared request budget (semaphore) with same-thread re-entry and ordered fan-out."""

    def __init__(self, max_requests: int, cancellation_manager=None):
        self.max_requests = max(1, int(max_requests or 1))
        self._sem = threading.BoundedSemaphore(self.max_requests)
        self._cancel = cancellation_manager
        self._local = threading.local()

    def _check_cancelled(self) -> None:
        if self._cancel is not None and self._cancel.is_cancelled():
            raise CancellationError("Batch process cancelled by user.")

    def in_slot(self) -> bool:
        return getattr(self._local, "depth", 0) > 0

    @contextmanager
    def slot(self):
        if self.in_slot():
            yield
            return
        self._check_cancelled()
        self._sem.acquire()
        self._local.depth = 1
        try:
            self._check_cancelled()
            yield
        finally:
            self._local.depth = 0
            self._sem.release()

    def run(self, fn: Callable[..., R], *args, **kwargs) -> R:
        with self.slot():
            return fn(*args, **kwargs)

    def map_ordered(self, jobs: Sequence[Callable[[], R]]) -> List[R]:
        if not jobs:
            return []
        if len(jobs) == 1:
            return [self.run(jobs[0])]
        with ThreadPoolExecutor(max_workers=min(len(jobs), self.max_requests)) as ex:
            futures = [ex.submit(self.run, j) for j in jobs]
            return [f.result() for f in futures]


def bboxes_overlap(first: BBox, second: BBox) -> bool:
    return not (first[2] <= second[0] or second[2] <= first[0] or first[3] <= second[1] or second[3] <= first[1])


def expanded_mask_bbox(mask: np.ndarray, image_size: Tuple[int, int], padding_ratio: float = 0.5, max_padding: int = 160,
                       min_padding: int = 64, extra_padding: int = 16) -> Optional[BBox]:
    m = np.asarray(mask)
    if m.ndim == 3:
        m = m[..., 0]
    ys, xs = np.where(m.astype(bool))
    if ys.size == 0:
        return None
    w, h = image_size
    x1, x2, y1, y2 = int(xs.min()), int(xs.max()) + 1, int(ys.min()), int(ys.max()) + 1
    pad = max(min_padding, int(min(max(x2 - x1, y2 - y1) * padding_ratio, max_padding))) + extra_padding
    return (max(0, x1 - pad), max(0, y1 - pad), min(w, x2 + pad), min(h, y2 + pad))


def partition_non_overlapping_waves(items: Iterable[T], get_bbox: Callable[[T], Optional[BBox]]) -> List[List[T]]:
    waves: List[List[T]] = []
    cur: List[T] = []
    boxes: List[BBox] = []
    for it in items:
        bb = get_bbox(it)
        if bb is None:
            if cur:
                waves.append(cur)
                cur, boxes = [], []
            waves.append([it])
            continue
        if boxes and any(bboxes_overlap(bb, o) for o in boxes):
            waves.append(cur)
            cur, boxes = [], []
        cur.append(it)
        boxes.append(bb)
    if cur:
        waves.append(cur)
    return waves


def paste_image_region(target: Image.Image, source: Image.Image, bbox: BBox) -> Image.Image:
    out = target.copy()
    out.paste(source.crop(bbox), (bbox[0], bbox[1]))
    return out


# ---- NEW in this build: page -> GPU sharding -----------------------------------------------------------------------
class PageShardCoordinator:
    """One process per GPU (torchrun).  `shard(items)` gives this rank's pages; `gather(obj)` collects one python
    object per rank on rank 0; `barrier()` brackets timed regions.  Works without torch.distributed (world size 1)."""

    def __init__(self, backend: Optional[str] = None):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self._dist = None
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29511")
                if be == "nccl":
                    torch.cuda.set_device(self.local_rank)
                    dist.init_process_group(be, rank=self.rank, world_size=self.world,
                                            device_id=torch.device("cuda", self.local_rank))
                else:
                    dist.init_process_group(be, rank=self.rank, world_size=self.world)
            self._dist = dist
        elif torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)

    def shard(self, items: Sequence[T]) -> List[T]:
        return [it for i, it in enumerate(items) if i % self.world == self.rank]

    def owner_of(self, index: int) -> int:
        return index % self.world

    def barrier(self) -> None:
        if self._dist is not None:
            self._dist.barrier()

    def gather(self, obj: Any) -> Optional[List[Any]]:
        if self._dist is None:
            return [obj]
        out: List[Any] = [None] * self.world if self.rank == 0 else None
        self._dist.gather_object(obj, out, dst=0)
        return out

    def broadcast(self, obj: Any, src: int = 0) -> Any:
        """The same python object on every rank (e.g. a default output directory named after rank 0's clock)."""
        if self._dist is None:
            return obj
        box = [obj if self.rank == src else None]
        self._dist.broadcast_object_list(box, src=src)
        return box[0]

    def all_reduce_max(self, value: float) -> float:
        if self._dist is None:
            return value
        import torch
        dev = torch.device("cuda", self.local_rank) if self._dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def close(self) -> None:
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
            self._dist = None
