"""TEST INFRASTRUCTURE ONLY — CPU oracle for the bubble-cleaning stage.

A compact restatement, on top of OpenCV (the reference's own dependency, opencv-contrib-python >= 4.8), of what the
reference computes per bubble (core/image/cleaning.py:210-521 `process_single_bubble`) and per page
(core/image/cleaning.py:524-1048 `clean_speech_bubbles`, default path: no Flux, no coloured-bubble classification).
It is pinned two ways (tests/test_clean_oracle.py): against the UNMODIFIED reference imported from /root/reference
when that tree is present (build container), and against the golden vectors in tests/golden/clean_*.json that
oracle/gen_golden.py produced from the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(mangatranslator_b200/) never does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np

# constants of the reference (core/image/cleaning.py:26-39)
_MIDPOINT = 128
_MIN_AREA = 50
_DIL = 7
_ERO = 5
_JUNCTION_MARGIN = 10
_JUNCTION_MIN_SHRINK = 1.0


@dataclass
class BubbleResult:
    ok: bool
    reason: str = ""
    mask: Optional[np.ndarray] = None            # H x W uint8 {0,255}
    fill_bgr: Optional[Tuple[int, int, int]] = None
    text_bbox: Optional[Tuple[int, int, int, int]] = None
    text_color: Optional[Tuple[int, ...]] = None
    used_otsu: bool = False


def _odd_kernel(base: int, scale: float) -> int:
    # core/scaling.py:64-96 (clamp to [1,63], round, bump even sizes to the next odd)
    s = 1.0 if scale is None or scale <= 0 else float(scale)
    k = int(round(min(63.0, max(1.0, base * s))))
    k = min(63, max(1, k))
    if k % 2 == 0:
        k = k + 1 if k + 1 <= 63 else k - 1
    return k


def scaled_params(roi_shrink_px: float, processing_scale: float):
    """(dilate k, erode k, shrink px, min contour area) — core/image/cleaning.py:629-648."""
    s = 1.0 if processing_scale is None or processing_scale <= 0 else float(processing_scale)
    shrink = min(64.0, max(0.0, roi_shrink_px * s))
    area = max(1, int(round(min(5000, max(_MIN_AREA, _MIN_AREA * s * s)))))
    return _odd_kernel(_DIL, s), _odd_kernel(_ERO, s), float(shrink), area


def _shrink(roi: np.ndarray, t: float, bbox, neighbors, scale: float) -> np.ndarray:
    dist = cv2.distanceTransform(roi, cv2.DIST_L2, 5)
    out = np.where(dist >= t, 255, 0).astype(np.uint8)
    if neighbors and bbox is not None:
        # conjoined junction zones keep everything that is >= 1 (scaled) px inside (cleaning.py:155-207)
        am = max(1, int(round(_JUNCTION_MARGIN * scale)))
        jm = max(1.0, _JUNCTION_MIN_SHRINK * scale)
        h, w = roi.shape
        x1, y1, x2, y2 = bbox
        for ox1, oy1, ox2, oy2 in neighbors:
            if x1 - am > ox2 or ox1 - am > x2 or y1 - am > oy2 or oy1 - am > y2:
                continue
            zx1, zy1 = max(0, max(x1, ox1) - am), max(0, max(y1, oy1) - am)
            zx2, zy2 = min(w, min(x2, ox2) + am), min(h, min(y2, oy2) + am)
            if zx2 <= zx1 or zy2 <= zy1:
                continue
            zone = dist[zy1:zy2, zx1:zx2] >= jm
            out[zy1:zy2, zx1:zx2][zone] = 255
    return out


def clean_bubble(mask: np.ndarray, gray: np.ndarray, image: np.ndarray, *, threshold: int, otsu: bool,
                 shrink_px: float, kd: int, ke: int, min_area: float, bbox=None, neighbors=None,
                 scale: float = 1.0) -> BubbleResult:
    """One bubble, one attempt (fixed threshold or Otsu)."""
    h, w = gray.shape
    m = np.where(mask > 0, 255, 0).astype(np.uint8)
    under = gray[m == 255]
    if under.size == 0:
        return BubbleResult(False, "empty mask")
    black = bool(np.mean(under) < _MIDPOINT)
    fill = (0, 0, 0) if black else (255, 255, 255)
    se_d = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (kd, kd))
    se_e = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ke, ke))
    roi = cv2.dilate(m, se_d)
    inside = roi == 255
    g = np.zeros_like(gray)
    g[inside] = gray[inside]
    if black:
        g = cv2.bitwise_not(g)
    if otsu:
        thr, _ = cv2.threshold(g[inside], 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    else:
        thr = threshold
    _, t = cv2.threshold(g, thr, 255, cv2.THRESH_BINARY)
    t = cv2.bitwise_and(t, roi)
    shrunk = _shrink(roi, float(shrink_px), bbox, neighbors, scale)
    t = cv2.bitwise_and(t, shrunk)
    gate = cv2.erode(m, se_e)
    contours, _ = cv2.findContours(t, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    keep = []
    for cnt in contours:
        if cv2.contourArea(cnt) <= min_area:
            continue
        mo = cv2.moments(cnt)
        if mo["m00"] == 0:
            continue
        cx, cy = int(mo["m10"] / mo["m00"]), int(mo["m01"] / mo["m00"])
        if 0 <= cx < w and 0 <= cy < h and gate[cy, cx] == 255:
            keep.append(cnt)
    if not keep:
        return BubbleResult(False, "no valid contour", used_otsu=otsu)
    valid = np.zeros((h, w), np.uint8)
    cv2.drawContours(valid, keep, -1, 255, thickness=cv2.FILLED)
    outer, _ = cv2.findContours(valid, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not outer:
        return BubbleResult(False, "no boundary contour", used_otsu=otsu)
    big = max(outer, key=cv2.contourArea)
    final = np.zeros((h, w), np.uint8)
    cv2.drawContours(final, [big], -1, 255, thickness=cv2.FILLED)
    x, y, bw, bh = cv2.boundingRect(big)
    # text colour: median of the original pixels under the (3x3-eroded) dark part of the shrunk ROI
    txt = cv2.bitwise_and(cv2.bitwise_not(t), shrunk)
    er = cv2.erode(txt, np.ones((3, 3), np.uint8))
    px = image[er == 255]
    if px.size == 0:
        px = image[txt == 255]
    tc = None
    if px.size > 0:
        med = tuple(np.median(px, axis=0).astype(int))
        sat = cv2.cvtColor(np.uint8([[med]]), cv2.COLOR_BGR2HSV)[0][0][1]
        if sat < 25:
            lum = 0.114 * fill[0] + 0.587 * fill[1] + 0.299 * fill[2]
            tc = (0, 0, 0) if lum >= 128 else (255, 255, 255)
        else:
            tc = tuple(int(v) for v in med)
    return BubbleResult(True, "", final, fill, (x, y, x + bw, y + bh), tc, otsu)


def clean_page(image_bgr: np.ndarray, detections: Sequence[dict], *, thresholding_value: int = 200,
               use_otsu_threshold: bool = False, roi_shrink_px: float = 5, processing_scale: float = 1.0):
    """Page-level oracle: (cleaned BGR(A) image, list of per-bubble dicts) like clean_speech_bubbles returns for
    pre-computed SAM-style detections (``sam_mask`` entries)."""
    img = np.ascontiguousarray(image_bgr)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY if img.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
    kd, ke, shrink, min_area = scaled_params(roi_shrink_px, processing_scale)
    s = 1.0 if processing_scale is None or processing_scale <= 0 else float(processing_scale)
    out = img.copy()
    bubbles: List[dict] = []
    for det in detections:
        m = det.get("sam_mask")
        if m is None:
            continue
        kw = dict(threshold=thresholding_value, shrink_px=shrink, kd=kd, ke=ke, min_area=min_area,
                  bbox=det.get("bbox"), neighbors=det.get("conjoined_neighbor_bboxes"), scale=s)
        r = clean_bubble(m, gray, img, otsu=use_otsu_threshold, **kw)
        if not r.ok and not use_otsu_threshold and r.reason != "skip":
            r = clean_bubble(m, gray, img, otsu=True, **kw)   # Otsu retry (cleaning.py:690-734)
        if not r.ok:
            continue
        bubbles.append({"mask": r.mask, "base_mask": np.where(m > 0, 255, 0).astype(np.uint8), "color": r.fill_bgr,
                        "bbox": det.get("bbox"), "is_colored": False, "text_bbox": r.text_bbox,
                        "text_color_bgr": r.text_color, "is_sam": True, "inpainted": False,
                        "used_otsu": r.used_otsu})
    groups = {}
    for b in bubbles:
        groups.setdefault(b["color"], []).append(b["mask"])
    for color, masks in groups.items():
        sel = np.bitwise_or.reduce(masks) == 255
        if out.shape[2] == 4:
            out[sel, :3] = color
        else:
            out[sel] = color
    return out, bubbles
