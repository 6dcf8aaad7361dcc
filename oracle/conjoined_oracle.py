"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy + OpenCV) of the reference's conjoined-bubble mask splitting,
core/image/detection.py: `_split_conjoined_mask` :971-1035 with `_seed_mask_from_box` :646-672,
`_split_overlap_zone_with_line` :675-800 (no OSB text boxes: offset 0), `_detect_group_arrangement` :803-839,
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


def _divide_by_line(zone, ca, cb, p0, p1):
    """Zone pixels to (a, b) by the side of the line p0->p1 they are on; None for a degenerate line."""
    vx, vy = p1[0] - p0[0], p1[1] - p0[1]
    norm = np.hypot(vx, vy)
    if norm < 1e-6:
        return None
    nx, ny = vy / norm, -vx / norm
    ys, xs = np.where(zone)
    dist = (xs - p0[0]) * nx + (ys - p0[1]) * ny
    sa = (ca[0] - p0[0]) * nx + (ca[1] - p0[1]) * ny - 0.0
    sb = (cb[0] - p0[0]) * nx + (cb[1] - p0[1]) * ny - 0.0
    if sa * sb > 0 or abs(sa - sb) < 1e-6:
        proj = (xs - (ca[0] + cb[0]) / 2.0) * (cb[0] - ca[0]) + (ys - (ca[1] + cb[1]) / 2.0) * (cb[1] - ca[1])
        to_a, to_b = proj <= 0, proj > 0
    elif sa < sb:
        to_a, to_b = (dist - 0.0) <= 0, (dist - 0.0) > 0
    else:
        to_a, to_b = (dist - 0.0) >= 0, (dist - 0.0) < 0
    ma, mb = np.zeros_like(zone), np.zeros_like(zone)
    ma[ys[to_a], xs[to_a]] = True
    mb[ys[to_b], xs[to_b]] = True
    return ma, mb


def _divide_zone(zone, a, b, arrangement):
    none = (np.zeros_like(zone), np.zeros_like(zone))
    ox0, oy0, ox1, oy1 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    if ox1 <= ox0 or oy1 <= oy0 or not zone.any():
        return none
    ca = ((a[0] + a[2]) / 2.0, (a[1] + a[3]) / 2.0)
    cb = ((b[0] + b[2]) / 2.0, (b[1] + b[3]) / 2.0)
    diag = ((ox1, oy0), (ox0, oy1)) if (cb[0] - ca[0]) * (cb[1] - ca[1]) >= 0 else ((ox0, oy0), (ox1, oy1))
    mx = float(np.clip((ca[0] + cb[0]) / 2.0, ox0, ox1))
    my = float(np.clip((ca[1] + cb[1]) / 2.0, oy0, oy1))
    first = {"horizontal": ((mx, oy0), (mx, oy1)), "vertical": ((ox0, my), (ox1, my))}.get(arrangement, diag)
    for p0, p1 in ([first] if first == diag else [first, diag]):
        got = _divide_by_line(zone, ca, cb, p0, p1)
        if got is not None:
            return got
    return none


def split_conjoined_mask(parent_mask, group_boxes, *, ipp: bool = False):
    """uint8 parent mask + K child boxes -> K uint8 masks {0,255}."""
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
    for i in range(len(boxes)):
        for j in range(i + 1, len(boxes)):
            zone = base & rects[i] & rects[j]
            if not zone.any():
                continue
            gi, gj = _divide_zone(zone, boxes[i], boxes[j], arrangement)
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


def split_group(parent_mask, group_boxes, *, ipp: bool = False):
    """What `_build_segmentation_detections` (:1157-1171, :1213-1228) does around the split: child rectangles are ORed
    into the parent first; returns (masks, rounded child bboxes)."""
    h, w = parent_mask.shape
    parent = np.asarray(parent_mask) > 0
    for b in group_boxes:
        parent = parent | rect_mask(b.tolist() if hasattr(b, "tolist") else b, h, w)
    masks = split_conjoined_mask(parent, group_boxes, ipp=ipp)
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
