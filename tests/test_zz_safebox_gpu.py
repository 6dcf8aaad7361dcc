"""GPU: the safe-text-box kernels (mtb_safe_boxes, through the C ABI) against the reference's golden vectors and the CPU
oracle — bit-exact boxes, centroid doubles, anchor rule, anchor pixel, maximal squared distance, mask bounds, errors.
Reference: core/image/image_utils.py:173-348 `calculate_centroid_expansion_box`.
(The first four tests ran green on a B200 in round 1, profiles/r01_safebox_timing.json; the clean -> text-box chain test
was written after the round's GPU budget was spent, which is why the file sorts last.)"""
import json
import os

import numpy as np
import pytest
import torch

import safebox_oracle as O
from helpers import ROOT, safebox_mask, safebox_page_masks, sha
from mangatranslator_b200 import safebox_host as S

pytestmark = pytest.mark.gpu

with open(os.path.join(ROOT, "tests", "golden", "safebox_golden.json")) as f:
    GOLD = json.load(f)["cases"]


def _expected(case):
    if "error" in case:
        return case["error"]
    return tuple(case["box"]), tuple(float.fromhex(v) for v in case["centroid"])


def _decode(rec):
    try:
        return S.decode(rec)
    except ValueError as e:
        return e.args[0]


def _oracle(mask, pad, trace=None):
    try:
        return O.safe_box(mask, pad, trace)
    except O.SafeBoxError as e:
        return e.args[0]


def _check_against_oracle(mask, pad, rec):
    t = {}
    exp = _oracle(mask, pad, t)
    assert _decode(rec) == exp
    if not isinstance(exp, str):
        assert (int(rec["moved"]), tuple(rec["anchor"]), int(rec["max_d2"])) == (t["moved"], t["anchor"], t["max_d2"])


def test_golden_vectors_one_launch_per_padding():
    """All golden masks (ragged sizes, rows not multiples of 16 bytes) grouped by padding, one launch per group."""
    from helpers import safebox_mask as gen
    dev = torch.device("cuda")
    by_pad = {}
    for c in GOLD:
        if c["name"].startswith("seed"):
            m, pad = gen(int(c["name"][4:]))
        else:
            m, pad = safebox_page_masks(0)[int(c["name"].split("bubble")[1])], c["padding"]
        assert sha(m) == c["mask_sha256"]
        by_pad.setdefault(pad, []).append((c, m))
    for pad, items in by_pad.items():
        recs = S.safe_boxes_device([torch.from_numpy(m).to(dev) for _, m in items], pad)
        for (c, m), rec in zip(items, recs):
            assert _decode(rec) == _expected(c), c["name"]
            if "error" not in c:
                ys, xs = np.nonzero(m)
                assert list(rec["mask_bbox"]) == [xs.min(), ys.min(), xs.max(), ys.max()], c["name"]


def test_random_masks_match_oracle_including_anchor_rule():
    dev = torch.device("cuda")
    masks, pads = zip(*[safebox_mask(s) for s in range(4000, 4120)])
    seen = set()
    for pad in sorted(set(pads)):
        idx = [i for i, p in enumerate(pads) if p == pad]
        recs = S.safe_boxes_device([torch.from_numpy(masks[i]).to(dev) for i in idx], pad)
        for i, rec in zip(idx, recs):
            _check_against_oracle(masks[i], pad, rec)
            seen.add(int(rec["moved"]) if int(rec["status"]) == 0 else -1)
    assert seen >= {-1, 0, 1, 2}


def test_full_size_page_and_bbox_bounded_workspace():
    """Twelve full-frame 1536x1024 masks in one launch; the same with the workspace sized from the bubbles' boxes;
    a pitched view (rows of a wider buffer) and an empty mask."""
    dev = torch.device("cuda")
    masks = safebox_page_masks(1)
    dm = [torch.from_numpy(m).to(dev) for m in masks]
    recs = S.safe_boxes_device(dm, 6.0)
    for m, rec in zip(masks, recs):
        _check_against_oracle(m, 6.0, rec)
    boxes = []
    for m in masks:
        ys, xs = np.nonzero(m)
        boxes.append((int(xs.min()) - 3, int(ys.min()) - 3, int(xs.max()) + 4, int(ys.max()) + 4))
    recs2 = S.safe_boxes_device(dm, 6.0, boxes)
    assert recs2.tobytes() == recs.tobytes()
    # determinism + idempotence on the same inputs
    assert S.safe_boxes_device(dm, 6.0).tobytes() == recs.tobytes()
    # rows of a wider buffer (pitch != width, base not 16-byte aligned)
    wide = torch.zeros((masks[0].shape[0], masks[0].shape[1] + 37), dtype=torch.uint8, device=dev)
    wide[:, 5:5 + masks[0].shape[1]] = dm[3]
    view = wide[:, 5:5 + masks[0].shape[1]]
    rec = S.safe_boxes_device([view, torch.zeros_like(dm[0])], 6.0)
    assert rec[0].tobytes() == recs[3].tobytes()
    assert int(rec[1]["status"]) == S.ST_EMPTY_MASK
    # a too-small planned box is reported, not overrun
    small = S.safe_boxes_device([dm[0]], 6.0, [(0, 0, 8, 8)])
    assert int(small[0]["status"]) == S.ST_WORKSPACE


def test_reference_shaped_function():
    from mangatranslator_b200.core.image import calculate_centroid_expansion_box
    from mangatranslator_b200.utils.exceptions import ImageProcessingError
    m, _ = safebox_mask(0)
    assert calculate_centroid_expansion_box(m, 4.0) == O.safe_box(m, 4.0)
    assert calculate_centroid_expansion_box(torch.from_numpy(m).cuda(), 2.5) == O.safe_box(m, 2.5)
    with pytest.raises(ImageProcessingError, match="Invalid or empty mask provided"):
        calculate_centroid_expansion_box(np.zeros((20, 20), np.uint8))
    with pytest.raises(ImageProcessingError, match="Safe area calculation failed"):
        calculate_centroid_expansion_box(m, 500.0)


def test_text_boxes_of_a_cleaned_page_follow_the_reference_chain():
    """clean -> final masks -> safe boxes, all on the device (CleanBatchResult.text_boxes), against the CPU chain the
    reference runs: clean_speech_bubbles' `mask` records -> calculate_centroid_expansion_box per bubble."""
    import clean_oracle
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.image.cleaning import clean_pages_device
    dev = torch.device("cuda")
    pg = synth.make_page(0, 768, 1024, n_bubbles=6)
    dets = synth.detections_from_page(pg)
    bgr = np.ascontiguousarray(pg.image_rgb[:, :, ::-1])
    scale = (768 * 1024 / 1e6) ** 0.5
    batch = clean_pages_device([torch.from_numpy(bgr).to(dev)], [dets], processing_scale=scale)
    _, bubbles = clean_oracle.clean_page(bgr, dets, processing_scale=scale)
    got = [b for b in batch.text_boxes(4.0)[0] if b is not None]
    assert len(got) == len(bubbles) > 0
    n_boxes = 0
    for g, b in zip(got, bubbles):
        assert g == _oracle(b["mask"], 4.0)
        n_boxes += not isinstance(g, str)
    assert n_boxes > 0
