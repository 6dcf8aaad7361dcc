"""Host side of the safe-text-box kernel (csrc/safebox_core.cuh): ctypes mirrors of the C-ABI structs, job planning and
result decoding.  Reference: core/image/image_utils.py:173-348 `calculate_centroid_expansion_box`.

Nothing is computed here: `safe_boxes_device` allocates the job table + workspace on the device and launches
`mtb_safe_boxes`; the masks stay where the segment / clean stages left them.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib, stream_ptr

ST_OK, ST_EMPTY_MASK, ST_NO_SAFE_AREA, ST_BAD_DIMS, ST_OUT_OF_BOUNDS, ST_WORKSPACE = range(6)
MOVED_POLE, MOVED_NEAREST = 1, 2
EMPTY_MESSAGE = "Invalid or empty mask provided"          # image_utils.py:204-205
FAILED_MESSAGE = "Safe area calculation failed"           # :348


class SafeBoxJob(C.Structure):
    _fields_ = [("mask", C.c_void_p), ("pitch", C.c_longlong), ("H", C.c_int), ("W", C.c_int), ("t2", C.c_uint),
                ("cap", C.c_int), ("g", C.c_void_p), ("safe", C.c_void_p)]


class SafeBoxResult(C.Structure):
    _fields_ = [("status", C.c_int), ("box", C.c_int * 4), ("moved", C.c_int), ("max_d2", C.c_int),
                ("anchor", C.c_int * 2), ("mask_bbox", C.c_int * 4), ("reserved", C.c_int),
                ("cx", C.c_double), ("cy", C.c_double)]


RESULT_DTYPE = np.dtype([("status", "<i4"), ("box", "<i4", (4,)), ("moved", "<i4"), ("max_d2", "<i4"),
                         ("anchor", "<i4", (2,)), ("mask_bbox", "<i4", (4,)), ("reserved", "<i4"),
                         ("cx", "<f8"), ("cy", "<f8")])
assert RESULT_DTYPE.itemsize == C.sizeof(SafeBoxResult)


def threshold_sq(padding_pixels: float) -> int:
    """`distance_map >= padding_pixels` (:218) on integer squared distances."""
    f = lib().mtb_safebox_threshold_sq
    f.argtypes, f.restype = [C.c_double], C.c_uint
    return int(f(float(padding_pixels)))


def window_cap(h: int, w: int, bbox: Optional[Sequence[int]] = None) -> int:
    """Workspace pixels for one mask: the whole framed image, or a known (x0, y0, x1, y1) upper bound of the mask's
    nonzero pixels (exclusive x1 / y1, e.g. the detection bbox the mask was clipped to) plus the ring."""
    if bbox is None:
        return (h + 2) * (w + 2)
    x0, y0, x1, y1 = (int(v) for v in bbox)
    return (min(max(x1 - x0, 1), w) + 2) * (min(max(y1 - y0, 1), h) + 2)


def safe_boxes_device(masks: Sequence[torch.Tensor], padding_pixels: float = 4.0,
                      bboxes: Optional[Sequence[Optional[Sequence[int]]]] = None) -> np.ndarray:
    """One launch for all `masks` (device uint8 HxW, rows contiguous, any common or differing sizes).  Returns the result
    records (numpy structured array, RESULT_DTYPE) — a few dozen bytes per bubble are all that crosses PCIe.
    `bboxes[i]` optionally bounds mask i's nonzero pixels so the workspace is sized to the bubble, not the page."""
    n = len(masks)
    if n == 0:
        return np.zeros(0, RESULT_DTYPE)
    dev = masks[0].device
    if dev.type != "cuda":
        raise RuntimeError("safe_boxes_device: masks must live on the CUDA device (no CPU fallback)")
    t2 = threshold_sq(padding_pixels)
    caps = []
    for i, m in enumerate(masks):
        if m.dtype != torch.uint8 or m.dim() != 2 or m.stride(1) != 1 or m.device != dev:
            raise ValueError("safe_boxes_device: masks must be uint8 HxW device tensors with contiguous rows")
        h, w = m.shape
        if h > 32766 or w > 32766:
            raise ValueError("safe_boxes_device: mask side above 32766 px")
        caps.append(window_cap(h, w, None if (bboxes is None or t2 == 0) else bboxes[i]))
    offs = np.concatenate([[0], np.cumsum([(c + 7) // 8 * 8 for c in caps])])
    g = torch.empty(int(offs[-1]), dtype=torch.int16, device=dev)
    safe = torch.empty(int(offs[-1]), dtype=torch.uint8, device=dev)
    jobs = (SafeBoxJob * n)()
    for i, m in enumerate(masks):
        j = jobs[i]
        j.mask, j.pitch, j.H, j.W = m.data_ptr(), m.stride(0), m.shape[0], m.shape[1]
        j.t2, j.cap = t2, caps[i]
        j.g, j.safe = g.data_ptr() + 2 * int(offs[i]), safe.data_ptr() + int(offs[i])
    jobs_dev = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(dev)
    res_dev = torch.empty(n * C.sizeof(SafeBoxResult), dtype=torch.uint8, device=dev)
    check(lib().mtb_safe_boxes(C.c_void_p(jobs_dev.data_ptr()), C.c_void_p(res_dev.data_ptr()), C.c_int(n),
                               C.c_void_p(stream_ptr())), "mtb_safe_boxes")
    out = res_dev.cpu().numpy().view(RESULT_DTYPE).copy()
    del g, safe, jobs_dev                              # kept alive until the D2H above has synchronised
    return out


def decode(rec) -> Tuple[Tuple[int, int, int, int], Tuple[float, float]]:
    """Result record -> the reference's return value, or the reference's error message as ValueError.args[0]."""
    st = int(rec["status"])
    if st == ST_OK:
        b = rec["box"]
        return (int(b[0]), int(b[1]), int(b[2]), int(b[3])), (float(rec["cx"]), float(rec["cy"]))
    if st == ST_EMPTY_MASK:
        raise ValueError(EMPTY_MESSAGE)
    if st == ST_WORKSPACE:
        raise RuntimeError("safe box: the mask exceeds the bounding box it was planned with")
    raise ValueError(FAILED_MESSAGE)
