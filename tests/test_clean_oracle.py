"""CPU: the cleaning oracle (oracle/clean_oracle.py) is pinned against (a) the golden vectors generated from the
unmodified reference (tests/golden/clean_golden.json, oracle/gen_golden.py) and (b) the reference itself when
/root/reference is present (build container only)."""
import numpy as np
import pytest

import clean_oracle
from helpers import build_clean_case, check_bubbles_against_golden, load_clean_golden, sha

GOLD = load_clean_golden()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_reference_golden(name):
    g = GOLD[name]
    bgr, dets = build_clean_case(g)
    out, bubbles = clean_oracle.clean_page(bgr, dets, thresholding_value=g["thresholding_value"],
                                           use_otsu_threshold=g["use_otsu"], roi_shrink_px=g["roi_shrink_px"],
                                           processing_scale=g["processing_scale"])
    assert list(out.shape) == g["cleaned_shape"]
    check_bubbles_against_golden(bubbles, g)
    assert sha(out) == g["cleaned_sha256"]


def test_oracle_matches_live_reference_random_bubbles():
    import _refimport
    if not _refimport.available():
        pytest.skip("reference tree not present (GPU box)")
    import cv2
    core = _refimport.import_reference()
    import core.image.cleaning as ref
    rng = np.random.default_rng(7)
    for it in range(12):
        h, w = int(rng.integers(150, 400)), int(rng.integers(150, 400))
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        mask = np.zeros((h, w), np.uint8)
        cv2.ellipse(mask, (int(rng.integers(0, w)), int(rng.integers(0, h))), (int(rng.integers(30, w // 2)),
                    int(rng.integers(30, h // 2))), 0, 0, 360, 255, -1)
        img[mask > 0] = 250
        for _ in range(20):
            p0 = (int(rng.integers(0, w)), int(rng.integers(0, h)))
            cv2.line(img, p0, (p0[0] + int(rng.integers(-25, 26)), p0[1] + int(rng.integers(-25, 26))), (10, 10, 10), 2)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        scale = float(rng.choice([0.8868, 1.0, 1.2541]))
        kd, ke, shrink, min_area = clean_oracle.scaled_params(5, scale)
        se_d = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (kd, kd))
        se_e = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ke, ke))
        try:
            r = ref.process_single_bubble(mask, gray, h, w, 200, False, shrink, False, None, True, se_d, se_e,
                                          min_area, False, None, scale, img)
        except Exception:
            r = None
        o = clean_oracle.clean_bubble(mask, gray, img, threshold=200, otsu=False, shrink_px=shrink, kd=kd, ke=ke,
                                      min_area=min_area, scale=scale)
        assert (r is not None) == o.ok
        if r is not None:
            assert np.array_equal(r[0], o.mask) and tuple(r[1]) == tuple(o.fill_bgr)
            assert tuple(r[4]) == tuple(o.text_bbox)
            assert (None if r[5] is None else tuple(int(v) for v in r[5])) == o.text_color
