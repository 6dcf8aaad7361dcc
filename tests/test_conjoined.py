"""Conjoined-bubble grouping and mask splitting (reference core/image/detection.py:345-472, 646-1035).

CPU:  * the grouping / arrangement geometry equals the UNMODIFIED reference functions on seeded random boxes (live, when
        /root/reference is present);
      * the oracle (oracle/conjoined_oracle.py) equals the reference's `_split_conjoined_mask` live and through
        tests/golden/conjoined_golden.json (hashes the reference produced, OpenCV's IPP dispatch off);
      * the split PLAN the host code hands to the kernel, walked in NumPy exactly like the kernel walks it
        (tests/split_emul.py), gives the oracle's masks bit for bit — including the closed-form chamfer norm that
        replaces cv2.distanceTransform.
GPU:  * `mtb_split_conjoined` equals the golden hashes and the oracle bit for bit;
      * detect_speech_bubbles / detect_pages_device route overlapping primaries through it like the reference.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import _refimport
import conjoined_oracle as O
from gen_golden_conjoined import CASES, make_case
from helpers import ROOT
from split_emul import apply_plan

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "conjoined_golden.json")))
needs_ref = pytest.mark.skipif(not _refimport.available(), reason="reference tree not present (GPU box)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _random_boxes(rng, n, w=1000, h=1400, overlap=True):
    out = []
    for _ in range(n):
        if overlap and out and rng.random() < 0.6:
            b = out[int(rng.integers(0, len(out)))]
            x0 = b[0] + rng.uniform(-0.8, 0.8) * (b[2] - b[0])
            y0 = b[1] + rng.uniform(-0.8, 0.8) * (b[3] - b[1])
        else:
            x0, y0 = rng.uniform(0, w * 0.8), rng.uniform(0, h * 0.8)
        out.append([x0, y0, x0 + rng.uniform(20, 300), y0 + rng.uniform(20, 300)])
    return torch.tensor(out, dtype=torch.float32)


# ---- grouping geometry vs the live reference --------------------------------------------------------------------
@needs_ref
def test_grouping_geometry_matches_live_reference():
    from mangatranslator_b200 import conjoined as Cj
    _refimport.import_reference()
    import core.image.detection as ref
    rng = np.random.default_rng(0)
    n_groups = n_conj = 0
    for _ in range(300):
        prim = _random_boxes(rng, int(rng.integers(1, 9)))
        sec = _random_boxes(rng, int(rng.integers(0, 9)))
        if len(sec):
            # secondary boxes mostly inside some primary, like RT-DETR children of a conjoined bubble
            for s in range(len(sec)):
                if rng.random() < 0.7:
                    p = prim[int(rng.integers(0, len(prim)))]
                    fx, fy = rng.uniform(0, 0.6), rng.uniform(0, 0.6)
                    sec[s] = torch.tensor([p[0] + fx * (p[2] - p[0]), p[1] + fy * (p[3] - p[1]),
                                           p[0] + (fx + 0.45) * (p[2] - p[0]), p[1] + (fy + 0.45) * (p[3] - p[1])])
            got = Cj.categorize_detections(prim, sec)
            exp = ref._categorize_detections(prim, sec)
            assert got == exp
            n_conj += len(exp[0])
        simple = sorted(rng.choice(len(prim), size=int(rng.integers(0, len(prim) + 1)), replace=False).tolist())
        got = Cj.detect_overlapping_primaries(prim, simple)
        exp = ref._detect_overlapping_primaries(prim, list(simple))
        assert got[0] == exp[0] and list(got[1]) == list(exp[1])
        n_groups += len(exp[0])
        grp = [b for b in prim[: int(rng.integers(1, len(prim) + 1))]]
        assert Cj.group_arrangement(grp) == ref._detect_group_arrangement(grp)
        # the oracle's own grouping (used by oracle/pipeline_oracle.py) is pinned to the same reference function
        allp = list(range(len(prim)))
        og, osimple = O.overlapping_groups(prim)
        eg, esimple = ref._detect_overlapping_primaries(prim, allp)
        assert og == eg and osimple == list(esimple)
    assert n_groups > 50 and n_conj > 20            # the random cases do exercise both kinds of group


# ---- the oracle vs the reference ----------------------------------------------------------------------------------
def _parent_with_rects(mask, boxes):
    parent = mask > 0
    for b in boxes:
        parent = parent | O.rect_mask(b, *mask.shape)
    return parent


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_golden(case):
    name, seed, h, w, k, layout = case
    g = GOLD[name]
    boxes, mask = make_case(seed, h, w, k, layout)
    assert sha(mask) == g["parent_sha256"] and boxes.tolist() == g["boxes"]
    masks, _ = O.split_group(mask, [b for b in torch.from_numpy(boxes)])
    assert [sha(m) for m in masks] == g["masks_sha256"]
    assert [int((m > 0).sum()) for m in masks] == g["mask_pixels"]


@needs_ref
def test_oracle_matches_live_reference_on_random_groups():
    import cv2
    _refimport.import_reference()
    import core.image.detection as ref
    rng = np.random.default_rng(5)
    flipped = total_rest = 0
    for t in range(40):
        h, w = int(rng.integers(60, 400)), int(rng.integers(60, 400))
        k = int(rng.integers(1, 5))
        boxes = _random_boxes(rng, k, w=w * 0.8, h=h * 0.8)
        boxes = boxes * torch.tensor([1, 1, 0.5, 0.5]) + torch.tensor([0, 0, 0.5, 0.5]) * boxes[:, [0, 1, 0, 1]]
        if t % 7 == 3:
            boxes[0] = torch.tensor([5.0, 5.0, 5.0, 9.0])             # empty child rectangle -> nearest-parent-pixel seed
        mask = (rng.random((h, w)) < 0.02).astype(np.uint8) * 255
        import cv2 as _cv
        mask = _cv.dilate(mask, np.ones((5, 5), np.uint8))
        parent = mask > 0 if t % 3 == 0 else _parent_with_rects(mask, boxes.tolist())
        blist = [b for b in boxes]
        cv2.ipp.setUseIPP(False)
        exp = ref._split_conjoined_mask(parent, blist)
        cv2.ipp.setUseIPP(True)
        exp_ipp = ref._split_conjoined_mask(parent, blist)
        got = O.split_conjoined_mask(parent, blist)
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b)
        flipped += sum(int((a != b).sum()) for a, b in zip(exp, exp_ipp)) // 2
        total_rest += int(parent.sum())
    # IPP's float chamfer only flips near-ties of the nearest-seed rule: a vanishing share of the parent pixels
    assert flipped <= 0.002 * total_rest, (flipped, total_rest)


def _text_case(rng, t):
    """Two to four overlapping child boxes of a bubble-like parent and OSB text boxes: usually one inside each child away
    from the overlap (a text-safe cut exists), sometimes spanning it, sometimes ambiguous or nested — every branch of
    the text-aware split."""
    h, w = int(rng.integers(120, 360)), int(rng.integers(160, 420))
    k = int(rng.integers(2, 5))
    horizontal = rng.random() < 0.5
    boxes = []
    for i in range(k):
        if horizontal:
            x0 = 10 + i * (w - 40) / k * rng.uniform(0.8, 1.0)
            boxes.append([x0, rng.uniform(5, 30), x0 + (w - 40) / k * rng.uniform(1.15, 1.6), h - rng.uniform(5, 30)])
        else:
            y0 = 10 + i * (h - 40) / k * rng.uniform(0.8, 1.0)
            boxes.append([rng.uniform(5, 30), y0, w - rng.uniform(5, 30), y0 + (h - 40) / k * rng.uniform(1.15, 1.6)])
    if t % 5 == 4:                                                  # a diagonal pair
        boxes = [[10.0, 10.0, w * 0.6, h * 0.6], [w * 0.4, h * 0.4, w - 10.0, h - 10.0]]
    texts = []
    for b in boxes:
        bw, bh = b[2] - b[0], b[3] - b[1]
        r = rng.random()
        if r < 0.7:                                                 # inside, small, near the middle of the child
            cx, cy = b[0] + bw * rng.uniform(0.35, 0.65), b[1] + bh * rng.uniform(0.35, 0.65)
            tw, th = bw * rng.uniform(0.1, 0.35), bh * rng.uniform(0.1, 0.35)
            texts.append([cx - tw / 2, cy - th / 2, cx + tw / 2, cy + th / 2])
        elif r < 0.85:                                              # wide: reaches into the neighbour
            texts.append([b[0] - bw * 0.2, b[1] + bh * 0.3, b[2] + bw * 0.2, b[1] + bh * 0.6])
    if texts and t % 4 == 1:                                        # a big box that nearly contains a small one
        s0 = texts[0]
        texts.append([s0[0] - 3, s0[1] - 3, s0[2] + 30, s0[3] + 20])
    if t % 6 == 2:
        texts.append([float(w + 50), float(h + 50), float(w + 90), float(h + 80)])      # outside everything
    yy, xx = np.mgrid[0:h, 0:w]
    parent = np.zeros((h, w), bool)
    for b in boxes:                                                 # an ellipse per child: a conjoined bubble outline
        parent |= ((xx - (b[0] + b[2]) / 2) / ((b[2] - b[0]) / 2 + 1e-6)) ** 2 + ((yy - (b[1] + b[3]) / 2) / ((b[3] - b[1]) / 2 + 1e-6)) ** 2 <= 1.0
    return h, w, torch.tensor(boxes, dtype=torch.float32), np.asarray(texts, np.float32).reshape(-1, 4), parent


@needs_ref
def test_text_aware_split_oracle_matches_live_reference():
    """OSB text boxes that belong to both children move the cut so that no text box is cut (:700-783, :893-905): the
    oracle's restatement equals the unmodified `_split_conjoined_mask`, `_match_text_boxes_to_bubbles` and
    `_get_group_osb_text_boxes` on seeded groups; the text boxes change the result in a good share of them."""
    import cv2
    _refimport.import_reference()
    import core.image.detection as ref
    rng = np.random.default_rng(17)
    moved = 0
    for t in range(60):
        h, w, boxes, texts, parent = _text_case(rng, t)
        blist = [b for b in boxes]
        union = torch.cat([boxes[:, :2].min(0).values, boxes[:, 2:].max(0).values])
        want_group = ref._get_group_osb_text_boxes(texts if len(texts) else None, union)
        got_group = O.group_text_boxes(texts if len(texts) else None, union)
        assert (want_group is None) == (got_group is None)
        if want_group is not None:
            assert np.array_equal(np.asarray(want_group), np.asarray(got_group))
            wm = ref._match_text_boxes_to_bubbles(want_group, blist)
            gm = O.match_text_boxes(got_group, [b.tolist() for b in blist])
            assert {i: [tuple(x) for x in v] for i, v in wm.items()} == {i: [tuple(x) for x in v] for i, v in gm.items()}
        full = parent.copy()
        for b in boxes.tolist():
            full |= O.rect_mask(b, h, w)
        cv2.ipp.setUseIPP(False)
        try:
            exp = ref._split_conjoined_mask(full, blist, osb_text_boxes=want_group)
            plain = ref._split_conjoined_mask(full, blist)
        finally:
            cv2.ipp.setUseIPP(True)
        got = O.split_conjoined_mask(full, blist, osb_text_boxes=got_group)
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b), t
        moved += int(any(not np.array_equal(a, b) for a, b in zip(exp, plain)))
    assert moved >= 15, moved


def _plan_with_text(parent, boxes, texts, h, w):
    from mangatranslator_b200 import conjoined as Cj
    full = parent.copy()
    for b in boxes.tolist():
        full |= O.rect_mask(b, h, w)

    def zone_pixels(i, j, rect):
        x0, y0, x1, y1 = rect
        ys, xs = np.nonzero(full[y0:y1, x0:x1])
        return xs + x0, ys + y0

    return Cj.plan_split(boxes, h, w, text_boxes=texts, zone_pixels=zone_pixels)


def test_text_aware_split_plan_walked_in_numpy_matches_oracle():
    """The product's host decisions for the text-aware split (which candidate line, which offset: conjoined.plan_split)
    walked through the NumPy statement of the kernel give the oracle's masks bit for bit; the offsets are non-zero in a
    good share of the cases; and the text-box helpers equal the oracle's."""
    from mangatranslator_b200 import conjoined as Cj
    rng = np.random.default_rng(17)
    with_offset = 0
    for t in range(60):
        h, w, boxes, texts, parent = _text_case(rng, t)
        union = torch.cat([boxes[:, :2].min(0).values, boxes[:, 2:].max(0).values])
        group = Cj.group_osb_text_boxes(texts if len(texts) else None, union)
        want_group = O.group_text_boxes(texts if len(texts) else None, union)
        assert (group is None) == (want_group is None)
        if group is not None:
            assert np.array_equal(np.asarray(group), np.asarray(want_group))
        plan = _plan_with_text(parent, boxes, group, h, w)
        with_offset += int(any(p[7] != 0.0 for p in plan.pairs))
        got = apply_plan(parent.astype(np.uint8) * 255, plan, include_child_rects=True)
        exp, _ = O.split_group(parent.astype(np.uint8) * 255, [b for b in boxes], osb_text_boxes=want_group)
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b), t
    assert with_offset >= 15, with_offset


# ---- the kernel's algorithm (plan + closed-form chamfer) walked in NumPy ----------------------------------------
@pytest.mark.parametrize("case", CASES[:6], ids=[c[0] for c in CASES[:6]])
def test_split_plan_walked_in_numpy_matches_golden(case):
    from mangatranslator_b200 import conjoined as Cj
    name, seed, h, w, k, layout = case
    boxes, mask = make_case(seed, h, w, k, layout)
    plan = Cj.plan_split(torch.from_numpy(boxes), h, w)
    assert Cj.group_arrangement(boxes.tolist()) == GOLD[name]["arrangement"]
    masks = apply_plan(mask, plan)
    assert [sha(m) for m in masks] == GOLD[name]["masks_sha256"]


def test_split_plan_handles_degenerate_groups():
    from mangatranslator_b200 import conjoined as Cj
    rng = np.random.default_rng(9)
    h, w = 120, 160
    mask = np.zeros((h, w), np.uint8)
    mask[20:90, 30:140] = 255
    # touching pixel rectangles whose float boxes do not overlap (mode 0), an empty child box, identical centres
    groups = [
        [[30.0, 20.0, 80.4, 90.0], [80.6, 20.0, 140.0, 90.0]],
        [[30.0, 20.0, 90.0, 90.0], [200.0, 200.0, 200.0, 260.0]],
        [[30.0, 20.0, 100.0, 90.0], [30.0, 20.0, 100.0, 90.0]],
        [[30.0, 20.0, 100.0, 60.0], [60.0, 40.0, 140.0, 90.0], [50.0, 30.0, 120.0, 80.0]],
    ]
    for g in groups:
        for with_rects in (True, False):
            parent = _parent_with_rects(mask, g) if with_rects else mask > 0
            exp = O.split_conjoined_mask(parent, g)
            got = apply_plan(parent.astype(np.uint8) * 255, Cj.plan_split(g, h, w), include_child_rects=False)
            for a, b in zip(got, exp):
                assert np.array_equal(a, b)
    del rng


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_split_kernel_matches_reference_golden(case):
    from mangatranslator_b200 import conjoined as Cj
    name, seed, h, w, k, layout = case
    boxes, mask = make_case(seed, h, w, k, layout)
    dev = torch.device("cuda:0")
    ys, xs = np.nonzero(mask)
    for window in (None, (int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1)):
        out, plan = Cj.split_conjoined_device(torch.from_numpy(mask).to(dev), torch.from_numpy(boxes), window=window)
        got = out.cpu().numpy()
        assert [sha(m) for m in got] == GOLD[name]["masks_sha256"]
        assert plan.bboxes == [tuple(int(round(float(v))) for v in b) for b in boxes]


@pytest.mark.gpu
def test_split_kernel_degenerate_groups_match_oracle():
    from mangatranslator_b200 import conjoined as Cj
    dev = torch.device("cuda:0")
    h, w = 120, 160
    mask = np.zeros((h, w), np.uint8)
    mask[20:90, 30:140] = 255
    groups = [
        [[30.0, 20.0, 80.4, 90.0], [80.6, 20.0, 140.0, 90.0]],
        [[30.0, 20.0, 90.0, 90.0], [200.0, 200.0, 200.0, 260.0]],
        [[30.0, 20.0, 100.0, 90.0], [30.0, 20.0, 100.0, 90.0]],
        [[30.0, 20.0, 100.0, 60.0], [60.0, 40.0, 140.0, 90.0], [50.0, 30.0, 120.0, 80.0]],
        [[30.0, 20.0, 100.0, 60.0]],
    ]
    for g in groups:
        for with_rects in (True, False):
            exp = O.split_group(mask, g)[0] if with_rects else O.split_conjoined_mask(mask, g)
            out, _ = Cj.split_conjoined_device(torch.from_numpy(mask).to(dev), g, include_child_rects=with_rects)
            for a, b in zip(out.cpu().numpy(), exp):
                assert np.array_equal(a, b)
    # an all-zero parent gives all-zero children
    out, _ = Cj.split_conjoined_device(torch.zeros((h, w), dtype=torch.uint8, device=dev), groups[0], include_child_rects=False)
    assert int(out.sum()) == 0
    # the largest group the kernel takes: a chain of MAX_CHILDREN overlapping boxes (105 candidate pairs)
    big = np.zeros((200, 900), np.uint8)
    big[30:170, 10:890] = 255
    chain = [[20.0 + 55 * i, 40.0 + 3 * (i % 3), 95.0 + 55 * i, 150.0 - 2 * (i % 2)] for i in range(Cj.MAX_CHILDREN)]
    exp = O.split_group(big, chain)[0]
    out, plan = Cj.split_conjoined_device(torch.from_numpy(big).to(dev), chain)
    assert len(plan.pairs) >= Cj.MAX_CHILDREN - 1
    for a, b in zip(out.cpu().numpy(), exp):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        Cj.plan_split(chain + [[5.0, 5.0, 50.0, 50.0]], 200, 900)


# ---- the whole detection flow with duck-typed models, against the UNMODIFIED reference -----------------------------
class _Boxes:
    def __init__(self, xyxy, conf, cls):
        self.xyxy, self.conf, self.cls = xyxy, conf, cls

    def __len__(self):
        return len(self.xyxy)


class _FakeDetector:
    """What the stage code touches of an ultralytics model / the RT-DETR adapter."""

    def __init__(self, boxes, confs, classes, names, want_imgsz):
        self.names = names
        self._r = (torch.tensor(boxes, dtype=torch.float32), torch.tensor(confs, dtype=torch.float32),
                   torch.tensor(classes, dtype=torch.float32))
        self.want_imgsz = want_imgsz

    def __call__(self, im, conf, device, verbose, imgsz, retina_masks=None):
        from types import SimpleNamespace
        assert imgsz == self.want_imgsz
        return [SimpleNamespace(boxes=_Boxes(*self._r), masks=None, orig_shape=im.shape[:2], names=self.names)]


class _FakeSamInputs(dict):
    def to(self, device):
        return self


class _FakeSamProcessor:
    """Full-resolution 'logits': an ellipse that bulges 8 % beyond each prompt box, plus two stray blobs."""

    def __call__(self, image, input_boxes=None, return_tensors="pt"):
        w, h = image.size
        return _FakeSamInputs(pixel_values=torch.zeros(1, 3, 8, 8), original_sizes=torch.tensor([[h, w]]),
                              input_boxes=input_boxes)

    def post_process_masks(self, pred_masks, original_sizes):
        return [pred_masks]


class _FakeSamModel:
    dtype = torch.float32

    def __call__(self, multimask_output=False, pixel_values=None, original_sizes=None, input_boxes=None):
        from types import SimpleNamespace
        h, w = [int(v) for v in original_sizes[0]]
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        out = []
        for b in input_boxes[0]:
            cx, cy, rx, ry = (b[0] + b[2]) / 2, (b[1] + b[3]) / 2, (b[2] - b[0]) * 0.54, (b[3] - b[1]) * 0.54
            m = (((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0).float()
            m[int(b[1]) + 2:int(b[1]) + 9, int(b[0]) + 3:int(b[0]) + 11] = 1.0          # stray blob in the box corner
            out.append(m)
        return SimpleNamespace(pred_masks=torch.stack(out).unsqueeze(1))


FLOW_CASES = {
    # primaries: two overlapping pairs (synthetic groups), a chain of three, and two lone bubbles
    "synthetic_groups": dict(
        primary=[[60, 80, 300, 330], [250, 120, 520, 360], [620, 90, 820, 300], [640, 260, 850, 470], [100, 600, 330, 820],
                 [290, 640, 520, 850], [480, 610, 700, 840], [760, 700, 900, 860], [50, 950, 250, 1150]],
        secondary=None),
    # a secondary detector that sees two children inside primary 0, one missed bubble and one text_free region
    "with_secondary": dict(
        primary=[[60, 80, 520, 360], [620, 90, 820, 300], [640, 260, 850, 470], [100, 600, 330, 820], [700, 900, 860, 1100]],
        secondary=dict(boxes=[[70, 90, 300, 350], [270, 100, 515, 355], [300, 1200, 520, 1400], [690, 890, 870, 1110],
                              [625, 95, 815, 295]],
                       confs=[0.8, 0.7, 0.6, 0.9, 0.5], classes=[0, 0, 0, 2, 0],
                       names={0: "bubble", 1: "text_bubble", 2: "text_free"})),
    # OSB text verification on (the reference's default): text boxes inside both lobes of the overlapping pairs move the
    # cut into the gap between the texts; one text box sticks out of a lone bubble (its box grows); one is ambiguous
    # between two lobes; one nearly contains another
    "synthetic_groups_with_osb_text": dict(
        primary=[[60, 80, 300, 330], [250, 120, 520, 360], [620, 90, 820, 300], [640, 260, 850, 470], [100, 600, 330, 820],
                 [290, 640, 520, 850], [480, 610, 700, 840], [760, 700, 900, 860], [50, 950, 250, 1150]],
        secondary=None,
        osb=[[90, 150, 262, 260], [310, 180, 480, 300], [640, 120, 800, 235], [660, 330, 830, 440], [120, 650, 250, 780],
             [340, 690, 470, 800], [545, 650, 680, 800], [850, 740, 930, 800], [255, 200, 300, 240], [86, 146, 290, 290]]),
}


def _run_reference_flow(case, seg_model):
    import cv2
    _refimport.import_reference()
    import core.image.detection as ref
    from core.caching import get_cache
    from core.ml.model_manager import ModelType, get_model_manager
    mm = get_model_manager()
    n = len(case["primary"])
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = _FakeDetector(case["primary"], [0.95 - 0.03 * i for i in range(n)], [0] * n,
                                                            {0: "speech_bubble"}, 1600)
    mm.models[ModelType.SAM2] = (_FakeSamProcessor(), _FakeSamModel())
    sec = case["secondary"]
    if sec is not None:
        mm.models[ModelType.RTDETR_CONJOINED_BUBBLE] = _FakeDetector(sec["boxes"], sec["confs"], sec["classes"], sec["names"], 640)
    else:
        mm.models.pop(ModelType.RTDETR_CONJOINED_BUBBLE, None)
    osb = case.get("osb")
    if osb is not None:
        mm.models[ModelType.YOLO_OSBTEXT] = _FakeDetector(osb, [0.9] * len(osb), [0] * len(osb), {0: "text"}, 640)
    get_cache().clear_all()
    from PIL import Image
    pil = Image.fromarray(np.full((1536, 1024, 3), 200, np.uint8))
    cv2.ipp.setUseIPP(False)
    try:
        return ref.detect_speech_bubbles(__import__("pathlib").Path("x.png"), "x.pt", 0.6, seg_model=seg_model, conjoined_detection=sec is not None,
                                         image_override=pil, device=torch.device("cpu"), osb_text_verification=osb is not None)
    finally:
        cv2.ipp.setUseIPP(True)
        for k in (ModelType.YOLO_SPEECH_BUBBLE, ModelType.SAM2, ModelType.RTDETR_CONJOINED_BUBBLE, ModelType.YOLO_OSBTEXT):
            mm.models.pop(k, None)


def _run_our_flow(case, seg_model, monkeypatch=None):
    from PIL import Image
    from mangatranslator_b200.core.caching import get_cache
    from mangatranslator_b200.core.image import detection as D
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    mm = get_model_manager()
    saved = dict(mm.models)
    n = len(case["primary"])
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = _FakeDetector(case["primary"], [0.95 - 0.03 * i for i in range(n)], [0] * n,
                                                            {0: "speech_bubble"}, 1600)
    mm.models[ModelType.SAM2] = (_FakeSamProcessor(), _FakeSamModel())
    sec = case["secondary"]
    if sec is not None:
        mm.models[ModelType.RTDETR_CONJOINED_BUBBLE] = _FakeDetector(sec["boxes"], sec["confs"], sec["classes"], sec["names"], 640)
    else:
        mm.models.pop(ModelType.RTDETR_CONJOINED_BUBBLE, None)
    osb = case.get("osb")
    if osb is not None:
        mm.models[ModelType.YOLO_OSBTEXT] = _FakeDetector(osb, [0.9] * len(osb), [0] * len(osb), {0: "text"}, 640)
    get_cache().clear()
    if monkeypatch is not None:          # CPU run: the kernel's algorithm walked in NumPy stands in for the kernel
        from mangatranslator_b200 import conjoined as Cj

        def emul(parent_mask, group_boxes, device, text_boxes=None):
            h, w = parent_mask.shape
            full = np.asarray(parent_mask) > 0
            for b in group_boxes:
                x0, y0, x1, y1 = Cj.box_rect(b, h, w)
                full[y0:y1, x0:x1] = True

            def zone_pixels(i, j, rect):
                ys, xs = np.nonzero(full[rect[1]:rect[3], rect[0]:rect[2]])
                return xs + rect[0], ys + rect[1]
            plan = Cj.plan_split(group_boxes, h, w, text_boxes=text_boxes, zone_pixels=zone_pixels)
            return apply_plan(parent_mask, plan, include_child_rects=True)
        monkeypatch.setattr(D, "_split_group_on_device", emul)
    try:
        pil = Image.fromarray(np.full((1536, 1024, 3), 200, np.uint8))
        return D.detect_speech_bubbles(__import__("pathlib").Path("x.png"), "x.pt", 0.6, seg_model=seg_model, conjoined_detection=sec is not None,
                                       image_override=pil, device=torch.device("cpu"), osb_text_verification=osb is not None)
    finally:
        mm.models.clear()
        mm.models.update(saved)


def _flow_digest(result):
    dets, free = result
    return dict(free=[[float(v) for v in b] for b in free],
                dets=[dict(bbox=[int(v) for v in d["bbox"]], confidence=round(float(d["confidence"]), 6), cls=d["class"],
                           neighbors=[[int(v) for v in b] for b in d.get("conjoined_neighbor_bboxes", [])] if "conjoined_neighbor_bboxes" in d else None,
                           mask_sha256=sha(np.asarray(d["sam_mask"])), mask_pixels=int((np.asarray(d["sam_mask"]) > 0).sum()))
                      for d in dets])


FLOW_GOLD_PATH = os.path.join(ROOT, "tests", "golden", "conjoined_flow_golden.json")


@needs_ref
@pytest.mark.parametrize("name", list(FLOW_CASES))
@pytest.mark.parametrize("seg_model", ["sam2", "yolo"])
def test_detection_flow_matches_live_reference(name, seg_model, monkeypatch):
    exp = _flow_digest(_run_reference_flow(FLOW_CASES[name], seg_model))
    got = _flow_digest(_run_our_flow(FLOW_CASES[name], seg_model, monkeypatch))
    assert got == exp
    assert any(d["neighbors"] for d in exp["dets"])                      # the case does contain conjoined children
    gold = json.load(open(FLOW_GOLD_PATH))
    assert gold[f"{name}/{seg_model}"] == exp                             # the committed fixture is what the reference says


@pytest.mark.parametrize("name", list(FLOW_CASES))
@pytest.mark.parametrize("seg_model", ["sam2", "yolo"])
def test_detection_flow_matches_reference_golden_cpu(name, seg_model, monkeypatch):
    gold = json.load(open(FLOW_GOLD_PATH))
    assert _flow_digest(_run_our_flow(FLOW_CASES[name], seg_model, monkeypatch)) == gold[f"{name}/{seg_model}"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FLOW_CASES))
@pytest.mark.parametrize("seg_model", ["sam2", "yolo"])
def test_detection_flow_on_device_matches_reference_golden(name, seg_model):
    """Same flow, the split running in mtb_split_conjoined."""
    gold = json.load(open(FLOW_GOLD_PATH))
    assert _flow_digest(_run_our_flow(FLOW_CASES[name], seg_model)) == gold[f"{name}/{seg_model}"]


@pytest.mark.gpu
def test_device_page_path_groups_overlapping_boxes_and_splits_on_device():
    """detect_pages_device with overlapping boxes: simple bubbles first, then the children of every synthetic group with
    neighbour lists; each child mask equals the oracle's split of the SAME parent mask the device SAM produced for the
    group's union box (so SAM's float noise is out of the comparison and the rest is bit-exact)."""
    from mangatranslator_b200 import conjoined as Cj
    from mangatranslator_b200 import synth
    from mangatranslator_b200 import weights as W
    from mangatranslator_b200.core.image.detection import detect_pages_device
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    from mangatranslator_b200.sam2 import Sam2B200
    from mangatranslator_b200.sam2_api import Sam2ModelB200, Sam2ProcessorB200
    from mangatranslator_b200.yolo import YoloB200
    dev = torch.device("cuda:0")
    mm = get_model_manager()
    mm.unload_all()
    ycfg = W.yolo_cfg("n")
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = YoloB200(W.yolo_state_dict(0, ycfg), ycfg, dev)
    cfg, sd = W.sam2_model_and_state(0)
    net = Sam2B200(sd, cfg, dev)
    mm.models[ModelType.SAM2] = (Sam2ProcessorB200(net), Sam2ModelB200(net))
    try:
        h, w = 640, 512
        pg = synth.make_page(3, h, w, n_bubbles=4)
        page = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).to(dev)
        boxes = np.array([[40, 50, 220, 240], [180, 90, 400, 300], [60, 380, 250, 600], [300, 400, 480, 590],
                          [330, 330, 500, 480]], np.float32)
        tb = torch.from_numpy(boxes)
        groups, simple = Cj.detect_overlapping_primaries(tb, list(range(len(boxes))))
        assert groups == [[0, 1], [3, 4]] and simple == [2]
        dets = detect_pages_device([page], injected_boxes=[boxes], imgsz=640, own_masks=True)[0]
        assert [d["bbox"] for d in dets] == [tuple(int(round(float(v))) for v in boxes[k]) for k in (2, 0, 1, 3, 4)]
        assert "conjoined_neighbor_bboxes" not in dets[0]
        assert dets[1]["conjoined_neighbor_bboxes"] == [dets[2]["bbox"]] and dets[4]["conjoined_neighbor_bboxes"] == [dets[3]["bbox"]]
        rgb = page[:, :, [2, 1, 0]].contiguous()
        pos = 1
        for g in groups:
            parent = net.decode(net.encode(rgb), Cj.union_box(tb[g]).unsqueeze(0), (h, w))[0].cpu().numpy()
            exp, _ = O.split_group(parent, [b for b in tb[g]])
            for k in range(len(g)):
                got = dets[pos + k]["sam_mask"].cpu().numpy()
                assert np.array_equal(got, exp[k])
                x0, y0, x1, y1 = dets[pos + k]["mask_bbox"]
                assert got[y0:y1, x0:x1].sum() == got.sum()                  # mask_bbox bounds the child mask
            pos += len(g)
        # the cleaning stage takes the children (with their neighbour boxes) like any other detection
        from mangatranslator_b200.core.image.cleaning import clean_pages_device
        import clean_oracle
        scale = (h * w / 1e6) ** 0.5
        batch = clean_pages_device([page], [dets], processing_scale=scale)
        host_dets = [dict(d, sam_mask=d["sam_mask"].cpu().numpy()) for d in dets]
        exp_page, _ = clean_oracle.clean_page(page.cpu().numpy(), host_dets, processing_scale=scale)
        assert np.array_equal(batch.pages_out[0].cpu().numpy(), exp_page)
    finally:
        mm.unload_all()
