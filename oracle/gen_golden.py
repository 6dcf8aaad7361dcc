"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.json by running the UNMODIFIED reference
(/root/reference, imported through oracle/_refimport.py) on seeded synthetic inputs.

Run in the build container (the reference tree does not exist on the GPU box):
    python oracle/gen_golden.py
The fixtures are small (hashes, boxes, colours) and committed; tests compare both the oracle restatement and the
CUDA path against them.
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _refimport  # noqa: E402
from mangatranslator_b200 import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CLEAN_CASES = [
    # name, seed, H, W, rgba, thresholding_value, use_otsu, roi_shrink_px
    ("p1024x768_s0", 0, 768, 1024, False, 200, False, 5),
    ("p1536x1024_s1", 1, 1536, 1024, False, 200, False, 5),
    ("p1536x1024_s2_rgba", 2, 1536, 1024, True, 200, False, 5),
    ("p1536x1024_s3_otsu", 3, 1536, 1024, False, 200, True, 5),
    ("p800x1150_s4_thr180", 4, 1150, 800, False, 180, False, 3),
    ("p640x480_s5_failing", 5, 480, 640, False, 254, False, 5),   # threshold so high that bubbles fail -> Otsu retry
]


def gen_clean():
    core = _refimport.import_reference()
    import core.image.cleaning as ref_clean
    out = {}
    for name, seed, h, w, rgba, thr, otsu, shrink in CLEAN_CASES:
        n_b = 12 if h * w > 600_000 else 6
        page = synth.make_page(seed, h, w, n_bubbles=n_b)
        dets = synth.detections_from_page(page)
        if name.endswith("s1"):  # give the conjoined pair its neighbour boxes like detection.py:1197-1205 would
            dets[0]["conjoined_neighbor_bboxes"] = [dets[1]["bbox"]]
            dets[1]["conjoined_neighbor_bboxes"] = [dets[0]["bbox"]]
        pil = Image.fromarray(page.image_rgb)
        if rgba:
            pil = pil.convert("RGBA")
        scale = (h * w / 1e6) ** 0.5
        img, info = ref_clean.clean_speech_bubbles(pil, "x.pt", pre_computed_detections=dets,
                                                   thresholding_value=thr, use_otsu_threshold=otsu,
                                                   roi_shrink_px=shrink, processing_scale=scale)
        out[name] = {
            "seed": seed, "H": h, "W": w, "rgba": rgba, "thresholding_value": thr, "use_otsu": otsu,
            "roi_shrink_px": shrink, "n_bubbles": n_b, "processing_scale": scale,
            "conjoined": name.endswith("s1"),
            "cleaned_sha256": sha(img), "cleaned_shape": list(img.shape),
            "bubbles": [{
                "bbox": [int(v) for v in b["bbox"]],
                "color": [int(v) for v in b["color"]],
                "text_bbox": [int(v) for v in b["text_bbox"]] if b["text_bbox"] is not None else None,
                "text_color_bgr": [int(v) for v in b["text_color_bgr"]] if b["text_color_bgr"] is not None else None,
                "mask_sha256": sha(b["mask"]), "mask_pixels": int((b["mask"] > 0).sum()),
            } for b in info],
        }
        print(name, "bubbles", len(info))
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    g = gen_clean()
    with open(os.path.join(ROOT, "tests", "golden", "clean_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote tests/golden/clean_golden.json")
