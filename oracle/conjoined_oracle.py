"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy + OpenCV) of the reference's conjoined-bubble mask splitting,
core/image/detection.py: `_split_conjoined_mask` :971-1035 with `_seed_mask_from_box` :646-672,
`_split_overlap_zone_with_line` :675-800 (incl. the text-safe offset when OSB text boxes belong to both children),
`_match_text_boxes_to_bubbles` :317-342, `_filter_encompassing_osb_text_boxes` :582-619, `_get_group_osb_text_boxes`
:622-638, `_detect_group_arrangement` :803-839,
`_split_overlap_zone_with_box_diagonal` :842-929, `_expand_resolved_masks_within_parent` :932-968, and the grouping
helpers `_categorize_detections` :345-405 / `_detect_overlapping_primaries` :408-472.

Pinned: tests/test_conjoined.py runs it against the UNMODIFIED reference functions on seeded random groups (live, when
/root/reference is present) and against tests/golden/conjoined_golden.json (hashes the reference produced,
oracle/gen_golden_conjoined.py).

`cv2.distanceTransform(DIST_L2, 5)` note.  The wheel in this image routes it through Intel IPP, a closed float32
implementation whose values differ from OpenCV's own C++ code (16.16 fixed point) in the last bits AND depend on the
position of the seed in the image (three different answers for one input: IPP, cv2.ipp.setUseIPP(False), and a plain
float32 two-pass emulation).  Nearest-seed ties are therefore decided by noise under IPP.  The parity target is
OpenCV's own implementation: `ipp=False` (default) switches IPP off around the call; `ipp=True` leaves the wheel's
default so tests can count the pixels IPP's noise flips.
"""
from __future__ import annotations

import contextlib

import cv2
import numpy as np

AXIS_RATIO = 3.0


@contextlib.contextmanager
def _ipp(enabled: bool):
    old = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(bool(enabled))
    try:
        yield
    finally:
        cv2.ipp.setUseIPP(old)


def rect_mask(box, h, w):
    x0f, y0f, x1f, y1f = [float(v) for v in box]
    x0, y0 = int(np.floor(max(0, min(x0f, w)))), int(np.floor(max(0, min(y0f, h))))
    x1, y1 = int(np.ceil(max(0, min(x1f, w)))), int(np.ceil(max(0, min(y1f, h))))
    m = np.zeros((h, w), bool)
    if x1 > x0 and y1 > y0:
        m[y0:y1, x0:x1] = True
    return m


def _arrangement(boxes):
    if len(boxes) < 2:
        return None
    cs = [((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0) for b in boxes]
    seen = None
    for i in range(len(cs)):
        for j in range(i + 1, len(cs)):
            dx, dy = abs(cs[j][0] - cs[i][0]), abs(cs[j][1] - cs[i][1])
            kind = "horizontal" if dx > AXIS_RATIO * max(dy, 1e-6) else "vertical" if dy > AXIS_RATIO * max(dx, 1e-6) else None
            if kind is None or (seen is not None and kind != seen):
                return None
            seen = kind
    return seen


TEXT_MATCH_IOA = 0.2            # OSB_TEXT_MATCH_IOA_THRESHOLD :20
AMBIGUOUS_RATIO = 0.85          # AMBIGUOUS_TEXT_MATCH_RATIO :23
TEXT_CONTAIN_IOA = 0.9          # OSB_TEXT_CONTAIN_IOA_THRESHOLD :26
NUDGE_INSET = 0.08              # OVERLAP_NUDGE_INSET_RATIO :29
MIN_SHARE = 0.08                # MIN_OVERLAP_SPLIT_SHARE :30


def _inter(a, b):
    return max(0.0, min(a[2], b[2]) - max(a[0], b[0])) * max(0.0, min(a[3], b[3]) - max(a[1], b[1]))


def _area(b):
    return max(0.0, b[2] - b[0]) * max(0.0, b[3] - b[1])


def _text_matches(t, b):
    inter, area = _inter(t, b), _area(t)
    if inter <= 0.0 or area <= 0.0:
        return False
    cx, cy = (t[0] + t[2]) / 2.0, (t[1] + t[3]) / 2.0
    return inter / area >= TEXT_MATCH_IOA or (b[0] <= cx <= b[2] and b[1] <= cy <= b[3])


def match_text_boxes(text_boxes, boxes):
    """Each text box goes to the bubble box it meaningfully overlaps most, unless the runner-up is nearly as good."""
    out = {i: [] for i in range(len(boxes))}
    for t in text_boxes:
        hits = []
        for i, b in enumerate(boxes):
            a = _inter(t[:4], b)
            if a > 0.0 and _text_matches(t[:4], b):
                hits.append((i, a))
        hits.sort(key=lambda it: it[1], reverse=True)
        if hits and not (len(hits) > 1 and hits[1][1] / hits[0][1] >= AMBIGUOUS_RATIO):
            out[hits[0][0]].append(t)
    return out


def filter_encompassing(text_boxes):
    """A text box that nearly contains a smaller one is dropped in favour of the smaller one."""
    if text_boxes is None or len(text_boxes) <= 1:
        return text_boxes
    bs = [np.asarray(t)[:4] for t in text_boxes]
    n = len(bs)
    keep = [True] * n
    for i in range(n):
        if not keep[i]:
            continue
        ai = _area(bs[i])
        if ai <= 0.0:
            keep[i] = False
            continue
        for j in range(n):
            if i == j or not keep[j]:
                continue
            aj = _area(bs[j])
            if aj <= 0.0 or ai <= aj:
                continue
            if _ioa(bs[j], bs[i]) > TEXT_CONTAIN_IOA:
                keep[i] = False
                break
    kept = [text_boxes[i] for i in range(n) if keep[i]]
    return np.asarray(kept) if kept else text_boxes


def group_text_boxes(text_boxes, parent_box):
    """The text boxes that intersect a group's parent box (then `filter_encompassing`); None when there are none."""
    if text_boxes is None or len(text_boxes) == 0:
        return None
    px0, py0, px1, py1 = [float(v) for v in (parent_box.tolist() if hasattr(parent_box, "tolist") else parent_box)]
    hits = [t for t in text_boxes if t[0] < px1 and t[2] > px0 and t[1] < py1 and t[3] > py0]
    return filter_encompassing(np.asarray(hits)) if hits else None


def _divide_by_line(zone, ca, cb, p0, p1, tba=None, tbb=None, text_safe=False):
    """Zone pixels to (a, b) by the side of the line p0->p1 they are on; None for a degenerate line — or, in the text-safe
    variant, when no offset keeps every text box of a and of b on its own side, or a side would get < 8 % of the zone."""
    vx, vy = p1[0] - p0[0], p1[1] - p0[1]
    norm = np.hypot(vx, vy)
    if norm < 1e-6:
        return None
    nx, ny = vy / norm, -vx / norm

    def sd(px, py):
        return (px - p0[0]) * nx + (py - p0[1]) * ny

    ys, xs = np.where(zone)
    if len(xs) == 0:
        return None
    dist = sd(xs, ys)
    tba, tbb = tba or [], tbb or []
    text_safe = text_safe and bool(tba) and bool(tbb)
    off = 0.0
    if text_safe:
        raw_lo, raw_hi = float(np.min(dist)), float(np.max(dist))
        inset = max(1.0, (raw_hi - raw_lo) * NUDGE_INSET)
        lo, hi = raw_lo + inset, raw_hi - inset
        if lo > hi:
            lo, hi = raw_lo, raw_hi

        def tighten(tboxes, cdist, lo, hi):
            if abs(cdist) < 1e-6:
                return lo, hi
            cd = []
            for t in tboxes:
                x0, y0, x1, y1 = [float(v) for v in t[:4]]
                for cx, cy in ((x0, y0), (x1, y0), (x0, y1), (x1, y1)):
                    cd.append(sd(cx, cy))
            if not cd:
                return lo, hi
            if cdist > 0:
                hi = min(hi, min(cd))
            else:
                lo = max(lo, max(cd))
            return lo, hi

        lo, hi = tighten(tba, sd(ca[0], ca[1]), lo, hi)
        lo, hi = tighten(tbb, sd(cb[0], cb[1]), lo, hi)
        if lo > hi:
            return None
        off = (lo + hi) / 2.0
    sa = sd(ca[0], ca[1]) - off
    sb = sd(cb[0], cb[1]) - off
    side = dist - off
    if sa * sb > 0 or abs(sa - sb) < 1e-6:
        proj = (xs - (ca[0] + cb[0]) / 2.0) * (cb[0] - ca[0]) + (ys - (ca[1] + cb[1]) / 2.0) * (cb[1] - ca[1])
        to_a, to_b = proj <= 0, proj > 0
    elif sa < sb:
        to_a, to_b = side <= 0, side > 0
    else:
        to_a, to_b = side >= 0, side < 0
    if text_safe and off != 0.0:
        need = max(1, int(np.ceil(len(xs) * MIN_SHARE)))
        if np.count_nonzero(to_a) < need or np.count_nonzero(to_b) < need:
            return None
    ma, mb = np.zeros_like(zone), np.zeros_like(zone)
    ma[ys[to_a], xs[to_a]] = True
    mb[ys[to_b], xs[to_b]] = True
    return ma, mb


def _divide_zone(zone, a, b, arrangement, tba=None, tbb=None):
    none = (np.zeros_like(zone), np.zeros_like(zone))
    ox0, oy0, ox1, oy1 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    if ox1 <= ox0 or oy1 <= oy0 or not zone.any():
        return none
    ca = ((a[0] + a[2]) / 2.0, (a[1] + a[3]) / 2.0)
    cb = ((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0)
    diag = ((ox1, oy0), (ox0, oy1)) if (cb[0] - ca[0]) * (cb[1] - ca[1]) >= 0 else ((ox0, oy0), (ox1, oy1))
    mx = float(np.clip((ca[0] + cb[0]) / 2.0, ox0, ox1))
    my = float(np.clip((ca[1] + cb[1]) / 2.0, oy0, oy1))
    h_line, v_line = ((ox0, my), (ox1, my)), ((mx, oy0), (mx, oy1))
    cands = {"horizontal": [v_line, diag, h_line], "vertical": [h_line, diag, v_line]}.get(arrangement, [diag, h_line, v_line])
    if tba and tbb:                                       # text-safe cut on the first candidate line that admits one
        for p0, p1 in cands:
            got = _divide_by_line(zone, ca, cb, p0, p1, tba, tbb, True)
            if got is not None:
                return got
    first = cands[0]
    for p0, p1 in ([first] if first == diag else [first, diag]):
        got = _divide_by_line(zone, ca, cb, p0, p1)
        if got is not None:
            return got
    return none


def split_conjoined_mask(parent_mask, group_boxes, *, ipp: bool = False, osb_text_boxes=None):
    """uint8 parent mask + K child boxes (+ the group's OSB text boxes) -> K uint8 masks {0,255}."""
    boxes = [[float(v) for v in (b.tolist() if hasattr(b, "tolist") else b)] for b in group_boxes]
    if parent_mask is None or not boxes:
        return []
    base = np.asarray(parent_mask) > 0
    if not base.any():
        return [np.zeros(base.shape, np.uint8) for _ in boxes]
    if len(boxes) == 1:
        return [base.astype(np.uint8) * 255]
    h, w = base.shape
    rects = [rect_mask(b, h, w) for b in boxes]
    owned = [base & r for r in rects]
    for k, m in enumerate(owned):
        if not m.any():                                   # nearest parent pixel to the box centre becomes the seed
            cy, cx = (boxes[k][1] + boxes[k][3]) / 2.0, (boxes[k][0] + boxes[k][2]) / 2.0
            ys, xs = np.where(base)
            n = int(np.argmin((xs - cx) ** 2 + (ys - cy) ** 2))
            m[ys[n], xs[n]] = True
    arrangement = _arrangement(boxes)
    text_for = match_text_boxes(osb_text_boxes, boxes) if osb_text_boxes is not None and len(osb_text_boxes) > 0 else None
    for i in range(len(boxes)):
        for j in range(i + 1, len(boxes)):
            zone = base & rects[i] & rects[j]
            if not zone.any():
                continue
            gi, gj = _divide_zone(zone, boxes[i], boxes[j], arrangement, text_for.get(i, []) if text_for else None,
                                  text_for.get(j, []) if text_for else None)
            owned[i] = (owned[i] & ~zone) | gi
            owned[j] = (owned[j] & ~zone) | gj
    taken = np.zeros_like(base)
    for m in owned:
        taken |= m
    rest = base & ~taken
    if not rest.any():
        return [m.astype(np.uint8) * 255 for m in owned]
    maps = []
    with _ipp(ipp):
        for m in owned:
            if m.any():
                maps.append(cv2.distanceTransform(np.where(m, 0, 1).astype(np.uint8), cv2.DIST_L2, 5).astype(np.float32))
            else:
                maps.append(np.full(base.shape, np.inf, np.float32))
    nearest = np.argmin(np.stack(maps, 0), 0)
    return [((m | (rest & (nearest == k))).astype(np.uint8) * 255) for k, m in enumerate(owned)]


def split_group(parent_mask, group_boxes, *, ipp: bool = False, osb_text_boxes=None):
    """What `_build_segmentation_detections` (:1157-1171, :1213-1228) does around the split: child rectangles are ORed
    into the parent first; returns (masks, rounded child bboxes)."""
    h, w = parent_mask.shape
    parent = np.asarray(parent_mask) > 0
    for b in group_boxes:
        parent = parent | rect_mask(b.tolist() if hasattr(b, "tolist") else b, h, w)
    masks = split_conjoined_mask(parent, group_boxes, ipp=ipp, osb_text_boxes=osb_text_boxes)
    bboxes = []
    for b in group_boxes:
        v = b.tolist() if hasattr(b, "tolist") else b
        bboxes.append((int(round(v[0])), int(round(v[1])), int(round(v[2])), int(round(v[3]))))
    return masks, bboxes


# ---- grouping (detection.py:408-472 `_detect_overlapping_primaries`, :1596-1619) -------------------------------------
def _ioa(inner, outer):
    area = max(0.0, inner[2] - inner[0]) * max(0.0, inner[3] - inner[1])
    if area <= 0:
        return 0.0
    iw = max(0.0, min(inner[2], outer[2]) - max(inner[0], outer[0]))
    ih = max(0.0, min(inner[3], outer[3]) - max(inner[1], outer[1]))
    return iw * ih / area


def overlapping_groups(boxes, threshold: float = 0.15):
    """Connected components (either-direction IoA > threshold) of size >= 2 among the boxes, as sorted index lists in order
    of their first member; and the indices left over as simple bubbles."""
    b = [[float(v) for v in (x.tolist() if hasattr(x, "tolist") else x)] for x in boxes]
    n = len(b)
    comp = list(range(n))

    def find(i):
        while comp[i] != i:
            comp[i] = comp[comp[i]]
            i = comp[i]
        return i

    for i in range(n):
        for j in range(i + 1, n):
            if _ioa(b[i], b[j]) > threshold or _ioa(b[j], b[i]) > threshold:
                ri, rj = find(i), find(j)
                if ri != rj:
                    comp[rj] = ri
    members = {}
    for i in range(n):
        members.setdefault(find(i), []).append(i)
    groups = [sorted(m) for m in members.values() if len(m) >= 2]
    grouped = {i for g in groups for i in g}
    return groups, [i for i in range(n) if i not in grouped]


def union_box(boxes):
    arr = np.asarray([[float(v) for v in (x.tolist() if hasattr(x, "tolist") else x)] for x in boxes], np.float32)
    return np.concatenate([arr[:, :2].min(0), arr[:, 2:].max(0)])
