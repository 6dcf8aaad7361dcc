"""Deterministic synthetic manga pages used by tests, smoke() and bench.py.

There is no network on the build or GPU boxes, so there are no real scans or checkpoints: a page is screentone
noise with B white elliptical speech bubbles (black outline, glyph-like black strokes inside), one black bubble
with white text, one gradient ("coloured") bubble and one overlapping (conjoined) pair.  Everything is a pure
function of ``seed`` (``numpy.random.default_rng(seed)``) so the CPU oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import cv2
import numpy as np


@dataclass
class SynthPage:
    image_rgb: np.ndarray            # H x W x 3 uint8
    boxes_xyxy: np.ndarray           # B x 4 float32 (ground-truth bubble boxes)
    masks: List[np.ndarray] = field(default_factory=list)  # B x (H x W uint8 {0,255}) ideal bubble masks
    kinds: List[str] = field(default_factory=list)         # "white" | "black" | "gradient"


def make_page(seed: int, height: int = 1536, width: int = 1024, n_bubbles: int = 12) -> SynthPage:
    rng = np.random.default_rng(seed)
    img = rng.integers(90, 170, size=(height, width, 3), dtype=np.uint8)
    # a few dark panel borders so the page is not pure noise
    for _ in range(3):
        y = int(rng.integers(height // 8, height - height // 8))
        cv2.line(img, (0, y), (width - 1, y), (20, 20, 20), 5)

    rows, cols = 4, 3
    if n_bubbles > rows * cols:
        rows = int(np.ceil(n_bubbles / cols))
    cell_h, cell_w = height / rows, width / cols
    scale = np.sqrt(height * width / (1536.0 * 1024.0))
    boxes, masks, kinds = [], [], []
    for b in range(n_bubbles):
        r, c = divmod(b, cols)
        ax = int(rng.integers(int(110 * scale), int(160 * scale) + 1))   # semi-axis x
        ay = int(rng.integers(int(110 * scale), int(160 * scale) + 1))   # semi-axis y
        ax = min(ax, int(cell_w * 0.47))
        ay = min(ay, int(cell_h * 0.47))
        cx = int((c + 0.5) * cell_w + rng.integers(-8, 9))
        cy = int((r + 0.5) * cell_h + rng.integers(-8, 9))
        if b == 1:  # conjoined with bubble 0: slide it left until the ellipses overlap
            cx = int(boxes[0][2] + ax * 0.80)
            cy = int((boxes[0][1] + boxes[0][3]) / 2 + 10)
        kind = "white"
        if b == n_bubbles - 1 and n_bubbles >= 3:
            kind = "black"
        elif b == n_bubbles - 2 and n_bubbles >= 4:
            kind = "gradient"
        mask = np.zeros((height, width), np.uint8)
        cv2.ellipse(mask, (cx, cy), (ax, ay), 0, 0, 360, 255, -1)
        fill = (255, 255, 255)
        ink = (0, 0, 0)
        if kind == "black":
            fill, ink = (8, 8, 8), (250, 250, 250)
        if kind == "gradient":
            ramp = np.linspace(215, 255, 2 * ax + 1, dtype=np.float32)
            xs = np.clip(np.arange(width) - (cx - ax), 0, 2 * ax).astype(int)
            g = ramp[xs][None, :].repeat(height, 0).astype(np.uint8)
            sel = mask == 255
            img[sel] = np.stack([g, (g * 0.97).astype(np.uint8), (g * 0.90).astype(np.uint8)], -1)[sel]
        else:
            img[mask == 255] = fill
        cv2.ellipse(img, (cx, cy), (ax, ay), 0, 0, 360, (0, 0, 0) if kind != "black" else (255, 255, 255), 3)
        # glyph-like strokes inside the inner 60 % of the bubble
        n_lines = int(rng.integers(3, 6))
        for li in range(n_lines):
            ty = int(cy - ay * 0.45 + li * (ay * 0.9 / max(n_lines - 1, 1)))
            half = int(ax * 0.55 * np.sqrt(max(0.05, 1 - ((ty - cy) / (ay * 0.9)) ** 2)))
            x = cx - half
            while x < cx + half - 6:
                gw = int(rng.integers(6, 15))
                gh = int(rng.integers(8, 17))
                for _ in range(int(rng.integers(2, 5))):
                    p0 = (int(x + rng.integers(0, gw)), int(ty + rng.integers(-gh // 2, gh // 2 + 1)))
                    p1 = (int(x + rng.integers(0, gw)), int(ty + rng.integers(-gh // 2, gh // 2 + 1)))
                    cv2.line(img, p0, p1, ink, int(rng.integers(1, 4)))
                x += gw + int(rng.integers(3, 8))
        x0, y0, x1, y1 = cx - ax, cy - ay, cx + ax, cy + ay
        jit = rng.uniform(-1.5, 1.5, size=4)
        boxes.append([max(0.0, x0 + jit[0]), max(0.0, y0 + jit[1]), min(width - 1.0, x1 + jit[2]),
                      min(height - 1.0, y1 + jit[3])])
        masks.append(mask)
        kinds.append(kind)
    return SynthPage(image_rgb=img, boxes_xyxy=np.asarray(boxes, np.float32), masks=masks, kinds=kinds)


def detections_from_page(page: SynthPage) -> list:
    """Detection dicts in the reference's format (core/image/detection.py:1119-1131): bbox ints + sam_mask u8."""
    dets = []
    for box, m in zip(page.boxes_xyxy, page.masks):
        x0, y0, x1, y1 = [int(round(float(v))) for v in box]
        dets.append({"bbox": (x0, y0, x1, y1), "confidence": 0.9, "class": "speech_bubble", "sam_mask": m.copy()})
    return dets
