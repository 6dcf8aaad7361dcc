"""TEST INFRASTRUCTURE ONLY — golden vectors for conjoined-bubble splitting, produced by the UNMODIFIED reference
(`core.image.detection._split_conjoined_mask`, with the child rectangles ORed into the parent like
`_build_segmentation_detections` does) on seeded cases; OpenCV's IPP dispatch is switched off while the reference runs
(see oracle/conjoined_oracle.py).  Writes tests/golden/conjoined_golden.json (case parameters + sha256 per child mask).

    python oracle/gen_golden_conjoined.py          # build container only: needs /root/reference
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _refimport  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_case(seed: int, h: int, w: int, k: int, layout: str):
    """Seeded group: k overlapping child boxes laid out side by side / stacked / diagonally, and a blobby parent mask
    (union of ellipses around the boxes plus stray blobs) clipped to the union box like a SAM mask would be."""
    rng = np.random.default_rng(seed)
    bw, bh = int(w * rng.uniform(0.18, 0.3)), int(h * rng.uniform(0.16, 0.28))
    x, y = w * rng.uniform(0.05, 0.2), h * rng.uniform(0.05, 0.25)
    boxes = []
    for i in range(k):
        jx, jy = rng.uniform(-0.06, 0.06) * bw, rng.uniform(-0.06, 0.06) * bh
        boxes.append([x + jx, y + jy, x + jx + bw * rng.uniform(0.9, 1.15), y + jy + bh * rng.uniform(0.9, 1.15)])
        step = rng.uniform(0.55, 0.85)
        if layout == "h":
            x += bw * step
        elif layout == "v":
            y += bh * step
        else:
            x += bw * step * 0.8
            y += bh * step * (0.7 if layout == "d" else -0.0) + (bh * 0.5 if layout == "z" and i % 2 == 0 else 0)
    boxes = np.asarray(boxes, np.float32)
    boxes[:, [0, 2]] = np.clip(boxes[:, [0, 2]], 0, w)
    boxes[:, [1, 3]] = np.clip(boxes[:, [1, 3]], 0, h)
    ux0, uy0, ux1, uy1 = boxes[:, 0].min(), boxes[:, 1].min(), boxes[:, 2].max(), boxes[:, 3].max()
    mask = np.zeros((h, w), np.uint8)
    for b in boxes:                                      # bubble bodies bulge a little beyond their boxes
        c = (int((b[0] + b[2]) / 2), int((b[1] + b[3]) / 2))
        ax = (int((b[2] - b[0]) * rng.uniform(0.5, 0.62)), int((b[3] - b[1]) * rng.uniform(0.5, 0.62)))
        cv2.ellipse(mask, c, ax, float(rng.uniform(-20, 20)), 0, 360, 255, -1)
    for _ in range(int(rng.integers(2, 7))):             # stray blobs inside the union box, outside the children
        c = (int(rng.uniform(ux0, ux1)), int(rng.uniform(uy0, uy1)))
        cv2.circle(mask, c, int(rng.integers(2, 14)), 255, -1)
    clip = np.zeros_like(mask)
    clip[int(np.floor(uy0)):int(np.ceil(uy1)), int(np.floor(ux0)):int(np.ceil(ux1))] = 255
    return boxes, mask & clip


CASES = [  # name, seed, H, W, K, layout
    ("pair_h", 1, 600, 800, 2, "h"), ("pair_v", 2, 800, 600, 2, "v"), ("pair_d", 3, 700, 700, 2, "d"),
    ("triple_h", 4, 500, 1000, 3, "h"), ("triple_z", 5, 900, 900, 3, "z"), ("quad_d", 6, 1536, 1024, 4, "d"),
    ("pair_h_page", 7, 1536, 1024, 2, "h"), ("triple_v_page", 8, 1536, 1024, 3, "v"),
]


def main():
    _refimport.import_reference()
    import core.image.detection as ref
    cv2.ipp.setUseIPP(False)
    out = {}
    for name, seed, h, w, k, layout in CASES:
        boxes, mask = make_case(seed, h, w, k, layout)
        parent = mask > 0
        tb = torch.from_numpy(boxes)
        for b in tb:
            parent = np.logical_or(parent, ref._build_rect_mask_from_box(b, h, w) > 0)
        covered = np.zeros_like(parent)
        for b in tb:
            covered |= ref._build_rect_mask_from_box(b, h, w) > 0
        masks = ref._split_conjoined_mask(parent, [b for b in tb])
        out[name] = dict(seed=seed, H=h, W=w, K=k, layout=layout, boxes=boxes.tolist(), parent_sha256=sha(mask),
                         arrangement=ref._detect_group_arrangement([b for b in tb]),
                         masks_sha256=[sha(m) for m in masks], mask_pixels=[int((m > 0).sum()) for m in masks],
                         leftover_pixels=int((parent & ~covered).sum()))      # pixels decided by the nearest-seed rule
        print(name, out[name]["arrangement"], out[name]["mask_pixels"], "leftover", out[name]["leftover_pixels"])
    with open(os.path.join(ROOT, "tests", "golden", "conjoined_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/conjoined_golden.json")


if __name__ == "__main__":
    main()
