"""TEST INFRASTRUCTURE ONLY — golden vectors for the safe text box, produced by the UNMODIFIED reference function
`calculate_centroid_expansion_box` (core/image/image_utils.py:173-348) on the seeded masks of tests/helpers.py
(`safebox_mask`: ten families x 12 seeds, plus twelve full-frame bubbles of a 1536x1024 page).  Each case stores the
mask's SHA-256 (so a drifting generator is noticed), the padding and the reference's return value — centroid as
float.hex() so no digit is lost — or its error message.  The reference is run twice, with OpenCV's IPP dispatch on (the
wheel's default) and off (OpenCV's own distance transform); the script refuses to write if the two disagree.

    python oracle/gen_golden_safebox.py          # build container only: needs /root/reference
"""
import json
import os
import sys

import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _refimport  # noqa: E402
from helpers import SAFEBOX_KINDS, safebox_mask, safebox_page_masks, sha  # noqa: E402

N_SEEDS = 12 * len(SAFEBOX_KINDS)


def run_reference(fn, err, mask, padding):
    try:
        box, c = fn(mask, padding)
        return {"box": [int(v) for v in box], "centroid": [float(c[0]).hex(), float(c[1]).hex()]}
    except err as e:
        return {"error": str(e)}


def main():
    _refimport.import_reference()
    import core.image.image_utils as RU
    from utils.exceptions import ImageProcessingError
    import utils.logging as RL
    RL.log_message = lambda *a, **k: None
    RU.log_message = lambda *a, **k: None
    cases = []
    inputs = [(f"seed{seed}", *safebox_mask(seed)) for seed in range(N_SEEDS)]
    inputs += [(f"page0_bubble{i}", m, 6.0) for i, m in enumerate(safebox_page_masks(0))]
    for name, mask, pad in inputs:
        outs = []
        for ipp in (True, False):
            cv2.ipp.setUseIPP(ipp)
            outs.append(run_reference(RU.calculate_centroid_expansion_box, ImageProcessingError, mask, pad))
        cv2.ipp.setUseIPP(True)
        if outs[0] != outs[1]:
            raise SystemExit(f"{name}: IPP on/off disagree: {outs}")
        cases.append({"name": name, "shape": list(mask.shape), "padding": pad, "mask_sha256": sha(mask), **outs[0]})
    out = {"reference": "core/image/image_utils.py:173-348 calculate_centroid_expansion_box (unmodified, live)",
           "cv2": cv2.__version__, "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "safebox_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    n_err = sum("error" in c for c in cases)
    print(f"wrote {path}: {len(cases)} cases, {n_err} reference failures")


if __name__ == "__main__":
    main()
