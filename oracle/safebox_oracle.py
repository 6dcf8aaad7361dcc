"""TEST INFRASTRUCTURE ONLY — CPU oracle of the renderer-side "safe text box" of a cleaned bubble mask
(reference: core/image/image_utils.py:173-348 `calculate_centroid_expansion_box`).

Restated on INTEGER squared distances, which is the formulation the CUDA kernel (csrc/safebox_core.cuh) implements:

  * `cv2.distanceTransform(padded, DIST_L2, DIST_MASK_PRECISE)` (:214-216) of OpenCV's own code is the exact Euclidean
    transform: float32 sqrt of the exact integer squared distance to the nearest zero pixel, the image being framed by a
    one-pixel ring of zeros (:210-213).  Checked in tests/test_safebox.py against cv2 with IPP disabled; the wheel's IPP
    build of that call is off by one float32 ulp on some pixels (not correctly rounded sqrt, alignment dependent), which
    only matters on exact ties: 1 of 3000 random masks, where two pixels share the maximal distance.
  * `dist >= padding_pixels` (:218) compares float32 with a Python float: NumPy 2 casts the scalar to float32.
  * `cv2.moments` of the 0/255 safe mask (:228-234): integer sums, m10/m00 = (255*Sx)/(255*S) in double.
  * `cv2.minMaxLoc` (:237): first maximum in raster order.
  * `dist_at_centroid < max_val * 0.70` (:247): float32 scalar against a double product cast to float32.
  * nearest safe pixel (:272-281): float64 sqrt((y-cy)^2 + (x-cx)^2), np.argmin -> first in raster order.
  * ray casts (:283-293), the -1 rule (:297-304), Python round (half to even) of the box corner (:314-318), bounds (:322-327).

Pinned by tests/test_safebox.py against the UNMODIFIED reference function (live, build container) and by the golden
vectors tests/golden/safebox_golden.json (oracle/gen_golden_safebox.py).
"""
from __future__ import annotations

import numpy as np


class SafeBoxError(Exception):
    """Stands for the reference's ImageProcessingError; `.args[0]` is the reference's message."""


EMPTY = "Invalid or empty mask provided"            # image_utils.py:204-205
FAILED = "Safe area calculation failed"             # :348 (every failure inside the try block ends here)


def squared_edt(mask: np.ndarray) -> np.ndarray:
    """Exact squared distance of every pixel to the nearest zero pixel of `mask` framed by a ring of zeros."""
    from scipy import ndimage
    h, w = mask.shape
    framed = np.zeros((h + 2, w + 2), bool)
    framed[1:-1, 1:-1] = mask != 0
    iy, ix = ndimage.distance_transform_edt(framed, return_distances=False, return_indices=True)
    yy, xx = np.indices(framed.shape)
    d2 = (yy - iy).astype(np.int64) ** 2 + (xx - ix).astype(np.int64) ** 2
    return d2[1:-1, 1:-1]


def threshold_sq(padding_pixels: float) -> int:
    """Smallest integer n with float32(sqrt(float32(n))) >= float32(padding): `dist >= padding` on integers."""
    p = np.float32(padding_pixels)
    if not p > 0:
        return 0
    n = int(np.ceil(float(p) * float(p)))
    while n > 0 and np.sqrt(np.float32(n - 1)) >= p:
        n -= 1
    while np.sqrt(np.float32(n)) < p:
        n += 1
    return n


def _first_zero_offset(line: np.ndarray) -> int | None:
    z = np.flatnonzero(line == 0)
    return int(z[0]) if z.size else None


def safe_box(mask: np.ndarray, padding_pixels: float = 4.0, trace: dict | None = None):
    """-> ((x, y, w, h), (cx, cy)) like the reference, or raises SafeBoxError with the reference's message.
    `trace` (optional) receives which anchor rule fired (`moved`: 1 = pole of inaccessibility, 2 = nearest safe pixel),
    the anchor pixel and the maximal squared distance — what the kernel reports next to the box."""
    trace = {} if trace is None else trace
    trace.update(moved=0, anchor=None, max_d2=0)
    if mask is None or not np.any(mask):
        raise SafeBoxError(EMPTY)
    h, w = mask.shape
    d2 = squared_edt(mask)
    safe = d2 >= threshold_sq(padding_pixels)
    n = int(safe.sum())
    if n == 0:
        raise SafeBoxError(FAILED)
    ys, xs = np.nonzero(safe)
    cx = float(255 * int(xs.sum())) / float(255 * n)
    cy = float(255 * int(ys.sum())) / float(255 * n)
    flat = int(np.argmax(d2))                                   # first maximum, raster order
    my, mx = divmod(flat, w)
    max_val = float(np.sqrt(np.float32(d2[my, mx])))
    trace["max_d2"] = int(d2[my, mx])
    qx = min(max(int(round(cx)), 0), w - 1)
    qy = min(max(int(round(cy)), 0), h - 1)
    if np.sqrt(np.float32(d2[qy, qx])) < np.float32(max_val * 0.70):
        cx, cy = float(mx), float(my)                           # pole of inaccessibility
        trace["moved"] |= 1
    px, py = int(round(cx)), int(round(cy))
    if not (0 <= px < w and 0 <= py < h and safe[py, px]):
        dist = np.sqrt((ys - cy) ** 2 + (xs - cx) ** 2)         # float64, separate IEEE ops
        k = int(np.argmin(dist))
        py, px = int(ys[k]), int(xs[k])
        cx, cy = float(px), float(py)
        trace["moved"] |= 2
    trace["anchor"] = (px, py)
    # distance to the nearest unsafe pixel of the anchor's row / column (or to the image edge if there is none)
    o = _first_zero_offset(safe[py, :px][::-1])
    left = px if o is None else o + 1
    o = _first_zero_offset(safe[py, px:])
    right = w - px if o is None else o
    o = _first_zero_offset(safe[:py, px][::-1])
    up = py if o is None else o + 1
    o = _first_zero_offset(safe[py:, px])
    down = h - py if o is None else o
    half_w, half_h = min(left, right), min(up, down)
    half_w = half_w - 1 if half_w > 1 else half_w
    half_h = half_h - 1 if half_h > 1 else half_h
    bw, bh = 2 * max(0, half_w), 2 * max(0, half_h)
    if bw <= 0 or bh <= 0:
        raise SafeBoxError(FAILED)
    bx, by = int(round(cx - bw / 2.0)), int(round(cy - bh / 2.0))
    if bx >= 0 and by >= 0 and bx + bw <= w and by + bh <= h:
        return (bx, by, bw, bh), (cx, cy)
    raise SafeBoxError(FAILED)
