"""Conjoined speech bubbles: grouping geometry on the host, mask splitting on the device.

Reference (core/image/detection.py): `_categorize_detections` :345-405 (secondary RT-DETR boxes contained in a primary
box), `_detect_overlapping_primaries` :408-472 (primary boxes that overlap each other = one bubble the detector cut
into sections; runs on EVERY page with two or more simple boxes, with or without the secondary detector),
`_detect_group_arrangement` :803-839, `_split_overlap_zone_with_box_diagonal` :842-929 / `_split_overlap_zone_with_line`
:675-800 (which line divides an overlap zone, and which side goes to whom) and `_split_conjoined_mask` :971-1035.

The box geometry is a few dozen float64 operations per group and stays in Python, operation for operation like the
reference so that every threshold decision and every line coefficient is the same double.  It produces a *split plan*
(child rectangles, one linear classifier per overlapping pair); the per-pixel work — seeds, zone re-division, nearest
seed for the left-over pixels — is `mtb_split_conjoined` (csrc/conjoined_kernels.cu).

Not restated: the OSB-text-aware variants of the split (text boxes nudge the cut; `require_text_safe_split`), because
OSB text detection is outside this build and the reference passes no text boxes unless
`use_osb_text_verification` is on (default off).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

IOA_THRESHOLD = 0.50                        # detection.py:16
IOA_OVERLAP_THRESHOLD = 0.5                 # :18
SYNTHETIC_CONJOINED_IOA_THRESHOLD = 0.15    # :31-33
AXIS_DOMINANCE_RATIO = 3.0                  # :34-36
MAX_CHILDREN = 15                           # MTB_SPLIT_MAX_CHILDREN


def _as_list(box) -> List[float]:
    return box.tolist() if hasattr(box, "tolist") else list(box)


def _inter_area(a, b) -> float:
    return max(0.0, min(a[2], b[2]) - max(a[0], b[0])) * max(0.0, min(a[3], b[3]) - max(a[1], b[1]))


def _ioa(inner, outer) -> float:
    area = max(0.0, inner[2] - inner[0]) * max(0.0, inner[3] - inner[1])
    return 0.0 if area <= 0 else _inter_area(inner, outer) / area


# ---- grouping --------------------------------------------------------------------------------------------------
def categorize_detections(primary_boxes, secondary_boxes, ioa_threshold: float = IOA_THRESHOLD):
    """A primary box holding two or more (not yet claimed) secondary boxes with IoA > threshold is a conjoined parent;
    a primary that merely duplicates a claimed secondary box disappears; the rest are simple.
    Returns ([(primary index, [secondary indices])], [simple primary indices])."""
    prim = [_as_list(b) for b in (primary_boxes.reshape(-1, 4) if hasattr(primary_boxes, "reshape") else primary_boxes)]
    sec = [_as_list(b) for b in (secondary_boxes.reshape(-1, 4) if hasattr(secondary_boxes, "reshape") else secondary_boxes)]
    conjoined, claimed = [], set()
    for i, pb in enumerate(prim):
        inside = [j for j, sb in enumerate(sec) if j not in claimed and _ioa(sb, pb) > ioa_threshold]
        if len(inside) >= 2:
            conjoined.append((i, inside))
            claimed.update(inside)
    parents = {p for p, _ in conjoined}
    simple = [i for i, pb in enumerate(prim)
              if i not in parents and not any(_ioa(sec[s], pb) > ioa_threshold for s in claimed)]
    return conjoined, simple


def detect_overlapping_primaries(primary_boxes, simple_indices: Sequence[int],
                                 ioa_threshold: float = SYNTHETIC_CONJOINED_IOA_THRESHOLD):
    """Transitive groups (union-find) of simple primary boxes whose mutual IoA exceeds the threshold in either
    direction.  Returns (groups as sorted index lists, in order of first appearance of their root; remaining simple)."""
    simple = list(simple_indices)
    if len(simple) < 2:
        return [], simple
    boxes = {i: _as_list(primary_boxes[i]) for i in simple}
    parent = {}

    def root(x):
        while parent.get(x, x) != x:
            parent[x] = parent.get(parent[x], parent[x])
            x = parent[x]
        return x

    linked = False
    for a_pos, a in enumerate(simple):
        for b in simple[a_pos + 1:]:
            if _ioa(boxes[a], boxes[b]) > ioa_threshold or _ioa(boxes[b], boxes[a]) > ioa_threshold:
                ra, rb = root(a), root(b)
                if ra != rb:
                    parent[rb] = ra
                linked = True
    if not linked:
        return [], simple
    by_root = {}
    for i in simple:
        by_root.setdefault(root(i), []).append(i)
    groups = [sorted(m) for m in by_root.values() if len(m) >= 2]
    grouped = {i for g in groups for i in g}
    return groups, [i for i in simple if i not in grouped]


def group_arrangement(group_boxes) -> Optional[str]:
    """'horizontal' / 'vertical' when every pair of box centres is offset mainly along that axis (ratio 3), else None."""
    if len(group_boxes) < 2:
        return None
    centres = []
    for b in group_boxes:
        b = _as_list(b)
        centres.append(((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0))
    found = None
    for i in range(len(centres)):
        for j in range(i + 1, len(centres)):
            dx, dy = abs(centres[j][0] - centres[i][0]), abs(centres[j][1] - centres[i][1])
            if dx > AXIS_DOMINANCE_RATIO * max(dy, 1e-6):
                kind = "horizontal"
            elif dy > AXIS_DOMINANCE_RATIO * max(dx, 1e-6):
                kind = "vertical"
            else:
                return None
            if found is None:
                found = kind
            elif found != kind:
                return None
    return found


# ---- split plan ------------------------------------------------------------------------------------------------
def box_rect(box, img_h: int, img_w: int) -> Tuple[int, int, int, int]:
    """Pixel rectangle of `_build_rect_mask_from_box` (:568-579): floor/ceil of the clamped box; (0,0,0,0) if empty."""
    x0f, y0f, x1f, y1f = _as_list(box)
    x0, y0 = int(np.floor(max(0, min(x0f, img_w)))), int(np.floor(max(0, min(y0f, img_h))))
    x1, y1 = int(np.ceil(max(0, min(x1f, img_w)))), int(np.ceil(max(0, min(y1f, img_h))))
    return (x0, y0, x1, y1) if (x1 > x0 and y1 > y0) else (0, 0, 0, 0)


class SplitPair(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("mode", C.c_int), ("reserved", C.c_int),
                ("cx", C.c_double), ("cy", C.c_double), ("ax", C.c_double), ("ay", C.c_double)]


@dataclass
class SplitPlan:
    rects: List[Tuple[int, int, int, int]]
    centers: List[Tuple[float, float]]
    pairs: List[Tuple[int, int, int, float, float, float, float]] = field(default_factory=list)  # i, j, mode, cx, cy, ax, ay
    bboxes: List[Tuple[int, int, int, int]] = field(default_factory=list)                        # rounded boxes (dict "bbox")


def _line_classifier(center_a, center_b, line_start, line_end):
    """`_split_overlap_zone_with_line` without text boxes (offset 0): returns (mode, cx, cy, ax, ay) or None when the
    line is degenerate.  mode 1: child a takes v <= 0, b takes v > 0; mode 2: a takes v >= 0, b takes v < 0."""
    lvx = line_end[0] - line_start[0]
    lvy = line_end[1] - line_start[1]
    length = np.hypot(lvx, lvy)
    if length < 1e-6:
        return None
    nx = lvy / length
    ny = -lvx / length

    def signed(px, py):
        return (px - line_start[0]) * nx + (py - line_start[1]) * ny

    side_a = signed(center_a[0], center_a[1]) - 0.0
    side_b = signed(center_b[0], center_b[1]) - 0.0
    if side_a * side_b > 0 or abs(side_a - side_b) < 1e-6:
        # both centres on one side of the line: fall back to the perpendicular bisector of the centres
        return (1, (center_a[0] + center_b[0]) / 2.0, (center_a[1] + center_b[1]) / 2.0,
                center_b[0] - center_a[0], center_b[1] - center_a[1])
    return (1 if side_a < side_b else 2, line_start[0], line_start[1], nx, ny)


def _pair_classifier(box_a, box_b, arrangement: Optional[str]):
    """`_split_overlap_zone_with_box_diagonal` without text boxes: the preferred line by arrangement, then the overlap
    diagonal.  Returns (mode, cx, cy, ax, ay); mode 0 = nobody gets the zone directly."""
    ox0, oy0 = max(box_a[0], box_b[0]), max(box_a[1], box_b[1])
    ox1, oy1 = min(box_a[2], box_b[2]), min(box_a[3], box_b[3])
    if ox1 <= ox0 or oy1 <= oy0:
        return (0, 0.0, 0.0, 0.0, 0.0)
    center_a = ((box_a[0] + box_a[2]) / 2.0, (box_a[1] + box_a[3]) / 2.0)
    center_b = ((box_b[0] + box_b[2]) / 2.0, (box_b[1] + box_b[3]) / 2.0)
    dx, dy = center_b[0] - center_a[0], center_b[1] - center_a[1]
    diag = ((ox1, oy0), (ox0, oy1)) if dx * dy >= 0 else ((ox0, oy0), (ox1, oy1))
    mid_x = float(np.clip((center_a[0] + center_b[0]) / 2.0, ox0, ox1))
    mid_y = float(np.clip((center_a[1] + center_b[1]) / 2.0, oy0, oy1))
    h_line = ((ox0, mid_y), (ox1, mid_y))
    v_line = ((mid_x, oy0), (mid_x, oy1))
    preferred = v_line if arrangement == "horizontal" else h_line if arrangement == "vertical" else diag
    lines = [preferred] if preferred == diag else [preferred, diag]
    for start, end in lines:
        res = _line_classifier(center_a, center_b, start, end)
        if res is not None:
            return res
    return (0, 0.0, 0.0, 0.0, 0.0)


def plan_split(group_boxes, img_h: int, img_w: int) -> SplitPlan:
    """Everything `_split_conjoined_mask` decides from the boxes alone."""
    boxes = [_as_list(b) for b in group_boxes]
    if len(boxes) > MAX_CHILDREN:
        raise ValueError(f"conjoined group of {len(boxes)} children (max {MAX_CHILDREN})")
    plan = SplitPlan(rects=[box_rect(b, img_h, img_w) for b in boxes],
                     centers=[((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0) for b in boxes],
                     bboxes=[(int(round(b[0])), int(round(b[1])), int(round(b[2])), int(round(b[3]))) for b in boxes])
    arrangement = group_arrangement(boxes)
    for i in range(len(boxes)):
        for j in range(i + 1, len(boxes)):
            ri, rj = plan.rects[i], plan.rects[j]
            if min(ri[2], rj[2]) <= max(ri[0], rj[0]) or min(ri[3], rj[3]) <= max(ri[1], rj[1]):
                continue                                    # the pixel rectangles do not meet: no overlap zone
            plan.pairs.append((i, j) + tuple(_pair_classifier(boxes[i], boxes[j], arrangement)))
    return plan


def union_box(boxes: torch.Tensor) -> torch.Tensor:
    """Parent box of a group: min corner / max corner over the member boxes (:1622-1631, :1709-1718)."""
    return torch.cat([boxes[:, :2].min(dim=0).values, boxes[:, 2:].max(dim=0).values])


# ---- device launch ---------------------------------------------------------------------------------------------
def _declare(l) -> None:
    if getattr(l, "_split_declared", False):
        return
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_longlong
    l.mtb_split_conjoined_workspace_bytes.argtypes = [i32, i32, i32]
    l.mtb_split_conjoined_workspace_bytes.restype = i64
    l.mtb_split_conjoined.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(i32), i32,
                                      C.POINTER(SplitPair), vp, vp, i64, vp]
    l.mtb_split_conjoined.restype = i32
    l._split_declared = True


def split_conjoined_device(parent_mask: torch.Tensor, group_boxes, *, include_child_rects: bool = True,
                           window: Optional[Tuple[int, int, int, int]] = None) -> Tuple[torch.Tensor, SplitPlan]:
    """parent_mask: device uint8 HxW (non-zero = bubble).  Returns (uint8 [K][H][W] child masks {0,255}, plan).

    `include_child_rects` ORs every child rectangle into the parent first, as `_build_segmentation_detections` does
    before it splits (:1161-1164, :1218-1221).  `window` (x0,y0,x1,y1) may bound the parent's pixels when the caller
    knows them (a SAM mask is already clipped to its prompt box); default: the whole frame."""
    from ._lib import check, lib, ptr, stream_ptr
    l = lib()
    _declare(l)
    assert parent_mask.is_cuda and parent_mask.dtype == torch.uint8 and parent_mask.dim() == 2
    h, w = int(parent_mask.shape[0]), int(parent_mask.shape[1])
    plan = plan_split(group_boxes, h, w)
    k = len(plan.rects)
    parent = parent_mask.contiguous()
    if include_child_rects:
        parent = parent.clone()
        for (x0, y0, x1, y1) in plan.rects:
            if x1 > x0 and y1 > y0:
                parent[y0:y1, x0:x1] = 255
    if window is None:
        win = (0, 0, w, h)
    else:
        xs = [window[0]] + [r[0] for r in plan.rects if r[2] > r[0]]
        ys = [window[1]] + [r[1] for r in plan.rects if r[2] > r[0]]
        xe = [window[2]] + [r[2] for r in plan.rects if r[2] > r[0]]
        ye = [window[3]] + [r[3] for r in plan.rects if r[2] > r[0]]
        win = (max(0, min(xs)), max(0, min(ys)), min(w, max(xe)), min(h, max(ye)))
    rects = (C.c_int * (4 * k))(*[v for r in plan.rects for v in r])
    centers = (C.c_double * (2 * k))(*[v for c in plan.centers for v in c])
    cwin = (C.c_int * 4)(*win)
    pairs = (SplitPair * max(1, len(plan.pairs)))()
    for q, (i, j, mode, cx, cy, ax, ay) in enumerate(plan.pairs):
        pairs[q].i, pairs[q].j, pairs[q].mode = i, j, mode
        pairs[q].cx, pairs[q].cy, pairs[q].ax, pairs[q].ay = float(cx), float(cy), float(ax), float(ay)
    nbytes = int(l.mtb_split_conjoined_workspace_bytes(win[3] - win[1], win[2] - win[0], k))
    work = torch.empty(nbytes, dtype=torch.uint8, device=parent.device)
    out = torch.empty((k, h, w), dtype=torch.uint8, device=parent.device)
    check(l.mtb_split_conjoined(ptr(parent), h, w, k, rects, centers, cwin, len(plan.pairs), pairs, ptr(out), ptr(work),
                                nbytes, stream_ptr()), "mtb_split_conjoined")
    return out, plan
