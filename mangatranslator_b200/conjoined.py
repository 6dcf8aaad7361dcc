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

OSB text boxes (`use_osb_text_verification`, detection.py:1555-1571): when text boxes belong to BOTH children of an
overlapping pair, the cut is moved along the line's normal into the gap between the two texts (`_split_overlap_zone_with_line`
with `require_text_safe_split`, :700-783; candidates in the order of `_split_overlap_zone_with_box_diagonal` :893-905).  That
decision needs the extent of the overlap zone's PIXELS along the normal and the pixel counts on each side, so for such a
pair (rare) the zone is read back from the device once and the offset is decided here in float64 NumPy, operation for
operation; the kernel then classifies `v - off`.  Pairs without text on both sides never leave the device.
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
OSB_TEXT_MATCH_IOA_THRESHOLD = 0.2          # :20
AMBIGUOUS_TEXT_MATCH_RATIO = 0.85           # :23
OSB_TEXT_CONTAIN_IOA_THRESHOLD = 0.9        # :26
OVERLAP_NUDGE_INSET_RATIO = 0.08            # :29
MIN_OVERLAP_SPLIT_SHARE = 0.08              # :30


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
                ("cx", C.c_double), ("cy", C.c_double), ("ax", C.c_double), ("ay", C.c_double), ("off", C.c_double)]


@dataclass
class SplitPlan:
    rects: List[Tuple[int, int, int, int]]
    centers: List[Tuple[float, float]]
    # i, j, mode, cx, cy, ax, ay, off
    pairs: List[Tuple[int, int, int, float, float, float, float, float]] = field(default_factory=list)
    bboxes: List[Tuple[int, int, int, int]] = field(default_factory=list)                        # rounded boxes (dict "bbox")


# ---- OSB text boxes -> which child they belong to (:91-106, :317-342, :582-638) ------------------------------------------
def _text_matches(t, b) -> bool:
    inter = _inter_area(t, b)
    area = max(0.0, t[2] - t[0]) * max(0.0, t[3] - t[1])
    if inter <= 0.0 or area <= 0.0:
        return False
    mx, my = (t[0] + t[2]) / 2.0, (t[1] + t[3]) / 2.0
    return inter / area >= OSB_TEXT_MATCH_IOA_THRESHOLD or (b[0] <= mx <= b[2] and b[1] <= my <= b[3])


def match_text_boxes_to_bubbles(text_boxes, boxes) -> dict:
    """child index -> the text boxes assigned to it: the child a text box meaningfully overlaps most, unless a second
    child overlaps it nearly as much (ratio >= 0.85: ambiguous, assigned to nobody)."""
    assigned = {i: [] for i in range(len(boxes))}
    for t in text_boxes:
        cands = []
        for i, b in enumerate(boxes):
            bl = _as_list(b)
            a = _inter_area(t[:4], bl)
            if a > 0.0 and _text_matches(t[:4], bl):
                cands.append((i, a))
        cands.sort(key=lambda c: c[1], reverse=True)
        if cands and not (len(cands) > 1 and cands[1][1] / cands[0][1] >= AMBIGUOUS_TEXT_MATCH_RATIO):
            assigned[cands[0][0]].append(t)
    return assigned


def filter_encompassing_text_boxes(text_boxes):
    """Drop a text box that nearly contains (IoA > 0.9) a smaller kept one: a detection spanning both lobes would block
    every text-safe cut."""
    if text_boxes is None or len(text_boxes) <= 1:
        return text_boxes
    bs = [[float(v) for v in np.asarray(t)[:4]] for t in text_boxes]
    n = len(bs)
    alive = [True] * n
    area = [max(0.0, b[2] - b[0]) * max(0.0, b[3] - b[1]) for b in bs]
    for i in range(n):
        if not alive[i]:
            continue
        if area[i] <= 0.0:
            alive[i] = False
            continue
        for j in range(n):
            if i == j or not alive[j] or area[j] <= 0.0 or area[i] <= area[j]:
                continue
            if _ioa(bs[j], bs[i]) > OSB_TEXT_CONTAIN_IOA_THRESHOLD:
                alive[i] = False
                break
    kept = [text_boxes[i] for i in range(n) if alive[i]]
    return np.asarray(kept) if kept else text_boxes


def group_osb_text_boxes(text_boxes, parent_box):
    """The page's OSB text boxes that intersect a group's parent box, without encompassing duplicates; None if none."""
    if text_boxes is None or len(text_boxes) == 0:
        return None
    px0, py0, px1, py1 = _as_list(parent_box)
    hits = [t for t in text_boxes if t[0] < px1 and t[2] > px0 and t[1] < py1 and t[3] > py0]
    return filter_encompassing_text_boxes(np.asarray(hits)) if hits else None


def _text_safe_classifier(zone_xs: np.ndarray, zone_ys: np.ndarray, center_a, center_b, line_start, line_end, tba, tbb):
    """`_split_overlap_zone_with_line(..., require_text_safe_split=True)` for a zone given by its pixel coordinates:
    (mode, cx, cy, ax, ay, off) or None when the line is degenerate, no offset keeps all text corners of a and of b on
    their own sides, or one side would keep less than 8 % of the zone."""
    lvx, lvy = line_end[0] - line_start[0], line_end[1] - line_start[1]
    length = np.hypot(lvx, lvy)
    if length < 1e-6 or len(zone_xs) == 0:
        return None
    nx, ny = lvy / length, -lvx / length

    def signed(px, py):
        return (px - line_start[0]) * nx + (py - line_start[1]) * ny

    dist = signed(zone_xs, zone_ys)
    raw_lo, raw_hi = float(np.min(dist)), float(np.max(dist))
    inset = max(1.0, (raw_hi - raw_lo) * OVERLAP_NUDGE_INSET_RATIO)
    lo, hi = raw_lo + inset, raw_hi - inset
    if lo > hi:
        lo, hi = raw_lo, raw_hi
    for tboxes, centre in ((tba, center_a), (tbb, center_b)):
        cdist = signed(centre[0], centre[1])
        if abs(cdist) < 1e-6:
            continue
        corners = []
        for t in tboxes:
            x0, y0, x1, y1 = [float(v) for v in t[:4]]
            corners += [signed(x0, y0), signed(x1, y0), signed(x0, y1), signed(x1, y1)]
        if not corners:
            continue
        if cdist > 0:
            hi = min(hi, min(corners))
        else:
            lo = max(lo, max(corners))
    if lo > hi:
        return None
    off = (lo + hi) / 2.0
    side_a = signed(center_a[0], center_a[1]) - off
    side_b = signed(center_b[0], center_b[1]) - off
    if side_a * side_b > 0 or abs(side_a - side_b) < 1e-6:
        res = (1, (center_a[0] + center_b[0]) / 2.0, (center_a[1] + center_b[1]) / 2.0,
               center_b[0] - center_a[0], center_b[1] - center_a[1], 0.0)
        v = (zone_xs - res[1]) * res[3] + (zone_ys - res[2]) * res[4]
    else:
        res = (1 if side_a < side_b else 2, line_start[0], line_start[1], nx, ny, off)
        v = dist - off
    if off != 0.0:
        to_a = np.count_nonzero(v <= 0) if res[0] == 1 else np.count_nonzero(v >= 0)
        to_b = np.count_nonzero(v > 0) if res[0] == 1 else np.count_nonzero(v < 0)
        need = max(1, int(np.ceil(len(zone_xs) * MIN_OVERLAP_SPLIT_SHARE)))
        if to_a < need or to_b < need:
            return None
    return res


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
                center_b[0] - center_a[0], center_b[1] - center_a[1], 0.0)
    return (1 if side_a < side_b else 2, line_start[0], line_start[1], nx, ny, 0.0)


def _pair_classifier(box_a, box_b, arrangement: Optional[str], tba=None, tbb=None, zone=None):
    """`_split_overlap_zone_with_box_diagonal`: with text boxes on both sides (`zone` = (xs, ys) of the overlap zone's
    pixels) the first candidate line that admits a text-safe cut; otherwise the preferred line by arrangement, then the
    overlap diagonal.  Returns (mode, cx, cy, ax, ay, off); mode 0 = nobody gets the zone directly."""
    none = (0, 0.0, 0.0, 0.0, 0.0, 0.0)
    ox0, oy0 = max(box_a[0], box_b[0]), max(box_a[1], box_b[1])
    ox1, oy1 = min(box_a[2], box_b[2]), min(box_a[3], box_b[3])
    if ox1 <= ox0 or oy1 <= oy0:
        return none
    center_a = ((box_a[0] + box_a[2]) / 2.0, (box_a[1] + box_a[3]) / 2.0)
    center_b = ((box_b[0] + box_b[2]) / 2.0, (box_b[1] + box_b[3]) / 2.0)
    dx, dy = center_b[0] - center_a[0], center_b[1] - center_a[1]
    diag = ((ox1, oy0), (ox0, oy1)) if dx * dy >= 0 else ((ox0, oy0), (ox1, oy1))
    mid_x = float(np.clip((center_a[0] + center_b[0]) / 2.0, ox0, ox1))
    mid_y = float(np.clip((center_a[1] + center_b[1]) / 2.0, oy0, oy1))
    h_line = ((ox0, mid_y), (ox1, mid_y))
    v_line = ((mid_x, oy0), (mid_x, oy1))
    preferred = v_line if arrangement == "horizontal" else h_line if arrangement == "vertical" else diag
    if tba and tbb and zone is not None and len(zone[0]):
        cands = ([v_line, diag, h_line] if arrangement == "horizontal" else
                 [h_line, diag, v_line] if arrangement == "vertical" else [diag, h_line, v_line])
        for start, end in cands:
            res = _text_safe_classifier(zone[0], zone[1], center_a, center_b, start, end, tba, tbb)
            if res is not None:
                return res
    lines = [preferred] if preferred == diag else [preferred, diag]
    for start, end in lines:
        res = _line_classifier(center_a, center_b, start, end)
        if res is not None:
            return res
    return none


def plan_split(group_boxes, img_h: int, img_w: int, text_boxes=None, zone_pixels=None) -> SplitPlan:
    """Everything `_split_conjoined_mask` decides from the boxes — and, for pairs whose children both own OSB text boxes,
    from the pixels of their overlap zone: `zone_pixels(i, j, rect)` -> (xs, ys) int arrays of the parent's pixels inside
    `rect` = (x0, y0, x1, y1), the intersection of the two child rectangles (called only for such pairs)."""
    boxes = [_as_list(b) for b in group_boxes]
    if len(boxes) > MAX_CHILDREN:
        raise ValueError(f"conjoined group of {len(boxes)} children (max {MAX_CHILDREN})")
    plan = SplitPlan(rects=[box_rect(b, img_h, img_w) for b in boxes],
                     centers=[((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0) for b in boxes],
                     bboxes=[(int(round(b[0])), int(round(b[1])), int(round(b[2])), int(round(b[3]))) for b in boxes])
    arrangement = group_arrangement(boxes)
    text_for = None
    if text_boxes is not None and len(text_boxes) > 0 and len(boxes) > 1:
        text_for = match_text_boxes_to_bubbles(text_boxes, boxes)
    for i in range(len(boxes)):
        for j in range(i + 1, len(boxes)):
            ri, rj = plan.rects[i], plan.rects[j]
            x0, y0, x1, y1 = max(ri[0], rj[0]), max(ri[1], rj[1]), min(ri[2], rj[2]), min(ri[3], rj[3])
            if x1 <= x0 or y1 <= y0:
                continue                                    # the pixel rectangles do not meet: no overlap zone
            tba = text_for.get(i, []) if text_for else []
            tbb = text_for.get(j, []) if text_for else []
            zone = zone_pixels(i, j, (x0, y0, x1, y1)) if (tba and tbb and zone_pixels is not None) else None
            plan.pairs.append((i, j) + tuple(_pair_classifier(boxes[i], boxes[j], arrangement, tba, tbb, zone)))
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
                           window: Optional[Tuple[int, int, int, int]] = None,
                           text_boxes=None) -> Tuple[torch.Tensor, SplitPlan]:
    """parent_mask: device uint8 HxW (non-zero = bubble).  Returns (uint8 [K][H][W] child masks {0,255}, plan).

    `include_child_rects` ORs every child rectangle into the parent first, as `_build_segmentation_detections` does
    before it splits (:1161-1164, :1218-1221).  `window` (x0,y0,x1,y1) may bound the parent's pixels when the caller
    knows them (a SAM mask is already clipped to its prompt box); default: the whole frame."""
    from ._lib import check, lib, ptr, stream_ptr
    l = lib()
    _declare(l)
    assert parent_mask.is_cuda and parent_mask.dtype == torch.uint8 and parent_mask.dim() == 2
    h, w = int(parent_mask.shape[0]), int(parent_mask.shape[1])
    parent = parent_mask.contiguous()
    if include_child_rects:
        parent = parent.clone()
        for b in group_boxes:
            x0, y0, x1, y1 = box_rect(b, h, w)
            if x1 > x0 and y1 > y0:
                parent[y0:y1, x0:x1] = 255

    def zone_pixels(i, j, rect):
        # a pair with OSB text on both sides: its overlap zone comes to the host once (a few thousand pixels)
        x0, y0, x1, y1 = rect
        ys, xs = np.nonzero(parent[y0:y1, x0:x1].cpu().numpy())
        return xs + x0, ys + y0

    plan = plan_split(group_boxes, h, w, text_boxes=text_boxes, zone_pixels=zone_pixels)
    k = len(plan.rects)
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
    for q, (i, j, mode, cx, cy, ax, ay, off) in enumerate(plan.pairs):
        pairs[q].i, pairs[q].j, pairs[q].mode = i, j, mode
        pairs[q].cx, pairs[q].cy, pairs[q].ax, pairs[q].ay = float(cx), float(cy), float(ax), float(ay)
        pairs[q].off = float(off)
    nbytes = int(l.mtb_split_conjoined_workspace_bytes(win[3] - win[1], win[2] - win[0], k))
    work = torch.empty(nbytes, dtype=torch.uint8, device=parent.device)
    out = torch.empty((k, h, w), dtype=torch.uint8, device=parent.device)
    check(l.mtb_split_conjoined(ptr(parent), h, w, k, rects, centers, cwin, len(plan.pairs), pairs, ptr(out), ptr(work),
                                nbytes, stream_ptr()), "mtb_split_conjoined")
    return out, plan
