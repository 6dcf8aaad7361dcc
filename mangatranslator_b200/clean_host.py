"""Host-side parameter scaling and job planning for the bubble-cleaning kernels.

Pure-integer/float host logic only (no pixels are touched here): it reproduces the reference's parameter scaling
(core/image/cleaning.py:629-648, core/scaling.py) and turns cv2's structuring elements and 5x5 chamfer metric into
the row half-width tables the CUDA kernel consumes (mtb_clean_params in include/mtb200.h).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Sequence, Tuple

import numpy as np

MAX_SE = 63
MAX_BALL = 65
MAX_NEIGHBORS = 16           # MTB_CLEAN_MAX_NEIGHBORS: >= MTB_SPLIT_MAX_CHILDREN - 1, so no neighbour of a split group is dropped
N_PLANES = 12
PLANE_FINAL = 9

# reference constants (core/image/cleaning.py:26-39)
MIN_CONTOUR_AREA = 50
DILATION_KERNEL_SIZE = (7, 7)
EROSION_KERNEL_SIZE = (5, 5)
JUNCTION_ADJACENCY_MARGIN = 10
JUNCTION_MIN_SHRINK = 1.0


class CleanParams(C.Structure):
    _fields_ = [
        ("thr_value", C.c_int), ("use_otsu", C.c_int), ("retry_otsu", C.c_int),
        ("kd", C.c_int), ("ke", C.c_int),
        ("sed_hw", C.c_int * MAX_SE), ("see_hw", C.c_int * MAX_SE),
        ("ball_r", C.c_int), ("ball_hw", C.c_int * (2 * MAX_BALL + 1)),
        ("jball_r", C.c_int), ("jball_hw", C.c_int * (2 * MAX_BALL + 1)),
        ("junction_margin", C.c_int),
        ("min_area", C.c_double),
        ("margin", C.c_int),
    ]


class CleanJob(C.Structure):
    _fields_ = [
        ("img", C.c_void_p), ("img_pitch", C.c_longlong),
        ("img_h", C.c_int), ("img_w", C.c_int), ("img_c", C.c_int),
        ("mask", C.c_void_p), ("mask_pitch", C.c_longlong),
        ("mask_x0", C.c_int), ("mask_y0", C.c_int), ("mask_w", C.c_int), ("mask_h", C.c_int),
        ("wx0", C.c_int), ("wy0", C.c_int), ("cw", C.c_int), ("ch", C.c_int),
        ("bbox", C.c_int * 4),
        ("n_neighbors", C.c_int),
        ("neighbors", (C.c_int * 4) * MAX_NEIGHBORS),
        ("work", C.c_void_p),
        ("max_runs", C.c_int), ("page_index", C.c_int),
    ]


class CleanResult(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("used_otsu", C.c_int), ("otsu_thr", C.c_int), ("is_black", C.c_int),
        ("fill_bgr", C.c_int * 3), ("text_bbox", C.c_int * 4),
        ("has_text_color", C.c_int), ("text_color", C.c_int * 4),
        ("n_components", C.c_int), ("n_valid", C.c_int), ("final_start", C.c_int),
        ("final_pixels", C.c_longlong), ("final_area", C.c_double),
        ("gray_sum", C.c_ulonglong), ("gray_cnt", C.c_uint),
    ]


# ---- reference parameter scaling (core/scaling.py:18-96) -----------------------------------------------------
def _normalize_scale(scale) -> float:
    if scale is None or scale <= 0:
        return 1.0
    return float(scale)


def scale_scalar(value, scale, minimum=None, maximum=None) -> float:
    v = value * _normalize_scale(scale)
    if minimum is not None:
        v = max(minimum, v)
    if maximum is not None:
        v = min(maximum, v)
    return v


def scale_area(value, scale, minimum=1.0, maximum=None) -> int:
    s = _normalize_scale(scale)
    v = value * (s * s)
    if minimum is not None:
        v = max(minimum, v)
    if maximum is not None:
        v = min(maximum, v)
    return max(1, int(round(v)))


def scale_kernel_dim(base: int, scale, minimum: int = 1, maximum: int = 63) -> int:
    d = scale_scalar(base, _normalize_scale(scale), minimum=float(minimum), maximum=float(maximum))
    k = max(minimum, int(round(d)))
    k = min(maximum, k)
    if k % 2 == 0:
        k = min(maximum, k + 1)
        if k % 2 == 0:
            k = max(minimum, k - 1)
            if k % 2 == 0:
                k = max(minimum, k + 1)
    return max(minimum, k)


# ---- structuring elements -------------------------------------------------------------------------------
def _cv_round(x: float) -> int:
    """cvRound / saturate_cast<int>(double): round half to even."""
    f = math.floor(x)
    d = x - f
    if d > 0.5 or (d == 0.5 and (int(f) & 1)):
        return int(f) + 1
    return int(f)


def ellipse_rows(k: int) -> List[int]:
    """Half-width of every row of cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)); -1 for an empty row.

    Restates imgproc/src/morph.dispatch.cpp getStructuringElement: r = c = k//2, row i covers
    [c - dx, c + dx] with dx = cvRound(c * sqrt((r*r - dy*dy) / (r*r))).  A 1x1 element is a rectangle.
    """
    if k == 1:
        return [0]
    r = c = k // 2
    inv_r2 = 1.0 / (float(r) * r) if r else 0.0
    rows = []
    for i in range(k):
        dy = i - r
        if abs(dy) <= r:
            dx = _cv_round(c * math.sqrt((r * r - dy * dy) * inv_r2))
            j1, j2 = max(c - dx, 0), min(c + dx + 1, k)
            rows.append((j2 - j1 - 1) // 2 if j2 > j1 else -1)
        else:
            rows.append(-1)
    return rows


# cv2.distanceTransform(DIST_L2, 5): metrics {1, 1.4, 2.1969} as 16.16 fixed point (distransform.cpp)
_CH_A = 1 << 16
_CH_B = int(round(float(np.float32(1.4)) * 65536))
_CH_C = int(round(float(np.float32(2.1969)) * 65536))


def chamfer_fixed(dx: int, dy: int) -> int:
    dx, dy = abs(dx), abs(dy)
    if dx < dy:
        dx, dy = dy, dx
    if 2 * dy <= dx:
        return (dx - 2 * dy) * _CH_A + dy * _CH_C
    return (2 * dy - dx) * _CH_B + (dx - dy) * _CH_C


def chamfer_ball_rows(t: float) -> Tuple[int, List[int]]:
    """Rows of the open chamfer ball {(dx,dy): dist(dx,dy) < t} with dist as cv2 reports it (float32 of the
    16.16 value) and t cast to float32 (NumPy >= 2 compares a float32 array with a Python float in float32,
    core/image/cleaning.py:330).  Returns (radius, half-widths for dy=-radius..radius, -1 = empty)."""
    t32 = np.float32(t)
    scale = np.float32(1.0 / 65536.0)

    def inside(dx, dy) -> bool:
        d = np.float32(np.float32(chamfer_fixed(dx, dy)) * scale)
        return bool(d < t32)

    if not inside(0, 0):
        return 0, [-1]
    r = 0
    while r + 1 <= MAX_BALL and inside(0, r + 1):
        r += 1
    rows = []
    for dy in range(-r, r + 1):
        w = -1
        while w + 1 <= 2 * MAX_BALL and inside(w + 1, dy):
            w += 1
        rows.append(w)
    return r, rows


def build_params(thresholding_value: int = 200, use_otsu_threshold: bool = False, roi_shrink_px: float = 5,
                 processing_scale: float = 1.0, retry_otsu: bool = True) -> CleanParams:
    """Mirror of the parameter block at core/image/cleaning.py:629-648 (+ junction constants :167-168)."""
    p = CleanParams()
    p.thr_value = int(thresholding_value)
    p.use_otsu = int(bool(use_otsu_threshold))
    p.retry_otsu = int(bool(retry_otsu) and not use_otsu_threshold)
    kd = scale_kernel_dim(DILATION_KERNEL_SIZE[0], processing_scale)
    ke = scale_kernel_dim(EROSION_KERNEL_SIZE[0], processing_scale)
    p.kd, p.ke = kd, ke
    for i in range(MAX_SE):
        p.sed_hw[i] = -1
        p.see_hw[i] = -1
    for i, w in enumerate(ellipse_rows(kd)):
        p.sed_hw[i] = w
    for i, w in enumerate(ellipse_rows(ke)):
        p.see_hw[i] = w
    eff_shrink = float(scale_scalar(roi_shrink_px, processing_scale, minimum=0.0, maximum=64.0))
    r, rows = chamfer_ball_rows(eff_shrink)
    p.ball_r = r
    for i in range(2 * MAX_BALL + 1):
        p.ball_hw[i] = -1
        p.jball_hw[i] = -1
    for i, w in enumerate(rows):
        p.ball_hw[i] = w
    jmin = max(1.0, JUNCTION_MIN_SHRINK * _normalize_scale(processing_scale))
    jr, jrows = chamfer_ball_rows(jmin)
    p.jball_r = jr
    for i, w in enumerate(jrows):
        p.jball_hw[i] = w
    p.junction_margin = max(1, int(round(JUNCTION_ADJACENCY_MARGIN * _normalize_scale(processing_scale))))
    p.min_area = float(scale_area(MIN_CONTOUR_AREA, processing_scale, minimum=MIN_CONTOUR_AREA, maximum=5000))
    # window margin: dilation reach + chamfer-ball reach + 2 (knight moves) + 1
    p.margin = kd // 2 + max(r, jr) + 3
    return p


def plan_window(mask_bbox: Sequence[int], img_w: int, img_h: int, params: "CleanParams") -> Tuple[int, int, int, int]:
    """Crop window (x0, y0, w, h) = mask bounding box (x0,y0,x1,y1 exclusive) grown by the margin, clipped to the page.

    With roi_shrink == 0 the reference's shrunk mask is the whole page (``dist >= 0`` everywhere,
    core/image/cleaning.py:330), so such jobs get the whole page as their window."""
    if all(params.ball_hw[i] < 0 for i in range(2 * params.ball_r + 1)):
        return 0, 0, img_w, img_h
    margin = params.margin
    x0, y0, x1, y1 = mask_bbox
    wx0, wy0 = max(0, x0 - margin), max(0, y0 - margin)
    wx1, wy1 = min(img_w, x1 + margin), min(img_h, y1 + margin)
    return wx0, wy0, max(1, wx1 - wx0), max(1, wy1 - wy0)


def default_max_runs(cw: int, ch: int) -> int:
    """Upper bound of the number of horizontal runs in a cw x ch window (never overflows)."""
    return ch * ((cw + 1) // 2) + 1


def workspace_words(cw: int, ch: int, max_runs: int) -> int:
    cwords = (cw + 31) // 32
    w = N_PLANES * cwords * ch + (ch + 1) + 4 * max_runs + 2 * max_runs + 2 * max_runs + max_runs
    return (w + 3) & ~3
