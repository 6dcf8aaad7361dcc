"""GPU: the CUDA cleaning path (through the reference-shaped stage function and the C ABI) is bit-exact against the
cv2 oracle and the golden vectors generated from the unmodified reference."""
import numpy as np
import pytest
import torch
from PIL import Image

import clean_oracle
from helpers import build_clean_case, check_bubbles_against_golden, load_clean_golden, sha

pytestmark = pytest.mark.gpu
GOLD = load_clean_golden()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_clean_speech_bubbles_matches_reference_golden(name):
    from mangatranslator_b200.core.image.cleaning import clean_speech_bubbles
    g = GOLD[name]
    bgr, dets = build_clean_case(g)
    pil = Image.fromarray(np.ascontiguousarray(bgr[:, :, [2, 1, 0, 3]] if g["rgba"] else bgr[:, :, ::-1]),
                          "RGBA" if g["rgba"] else "RGB")
    out, bubbles = clean_speech_bubbles(pil, "x.pt", pre_computed_detections=dets,
                                        thresholding_value=g["thresholding_value"],
                                        use_otsu_threshold=g["use_otsu"], roi_shrink_px=g["roi_shrink_px"],
                                        processing_scale=g["processing_scale"])
    assert list(out.shape) == g["cleaned_shape"]
    check_bubbles_against_golden(bubbles, g)
    assert sha(out) == g["cleaned_sha256"]


def test_clean_batch_matches_oracle_random_pages():
    """Batched launch (several pages, ragged bubble counts, an empty mask, a page with no detections)."""
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.image.cleaning import clean_pages_device
    dev = torch.device("cuda:0")
    pages, dets_all, bgrs = [], [], []
    for seed, (h, w, nb) in enumerate([(768, 1024, 12), (600, 400, 5), (1536, 1024, 12), (300, 300, 0)]):
        pg = synth.make_page(100 + seed, h, w, n_bubbles=max(nb, 1))
        dets = synth.detections_from_page(pg)[:nb]
        if seed == 1:
            dets.append({"bbox": (1, 1, 5, 5), "sam_mask": np.zeros((h, w), np.uint8)})  # empty mask -> skipped
        bgr = np.ascontiguousarray(pg.image_rgb[:, :, ::-1])
        bgrs.append(bgr)
        pages.append(torch.from_numpy(bgr).to(dev))
        dets_all.append(dets)
    scale = 1.1
    batch = clean_pages_device(pages, dets_all, processing_scale=scale)
    for pi, (bgr, dets) in enumerate(zip(bgrs, dets_all)):
        exp_img, exp_b = clean_oracle.clean_page(bgr, dets, processing_scale=scale)
        assert np.array_equal(batch.pages_out[pi].cpu().numpy(), exp_img), pi
        ok = [(k, r) for k, r in enumerate(batch.results[pi]) if r is not None and r.status == 0]
        assert len(ok) == len(exp_b)
        for (k, r), e in zip(ok, exp_b):
            assert np.array_equal(batch.export_mask(pi, k).cpu().numpy(), e["mask"])
            assert tuple(r.text_bbox) == tuple(e["text_bbox"]) and tuple(r.fill_bgr) == tuple(e["color"])


def test_clean_full_size_properties():
    """BASELINE size (1536x1024, 12 bubbles): idempotence — cleaning an already cleaned page with the same masks
    leaves it unchanged, and every final mask lies inside its dilated detection mask."""
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.image.cleaning import clean_pages_device
    dev = torch.device("cuda:0")
    pg = synth.make_page(42)
    dets = synth.detections_from_page(pg)
    page = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).to(dev)
    scale = (1536 * 1024 / 1e6) ** 0.5
    b1 = clean_pages_device([page], [dets], processing_scale=scale)
    b2 = clean_pages_device([b1.pages_out[0]], [dets], processing_scale=scale)
    assert torch.equal(b1.pages_out[0], b2.pages_out[0])
    import cv2
    for k, (d, r) in enumerate(zip(dets, b1.results[0])):
        assert r.status == 0
        m = b1.export_mask(0, k).cpu().numpy()
        roi = cv2.dilate(d["sam_mask"], cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (9, 9)))
        assert not np.any((m > 0) & (roi == 0))
