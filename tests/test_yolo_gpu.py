"""GPU: B200 YOLOv8-seg (tcgen05 conv plans with channel-slice concats, SPPF/upsample kernels, DFL decode, NMS,
scale_boxes, reference dedup/containment) against the CPU fp32 oracle (oracle/yolo_oracle.py) and the reference's own
post-NMS functions."""
import numpy as np
import pytest
import torch

import yolo_oracle as Y

pytestmark = pytest.mark.gpu
NANO = dict(nc=1, depth=0.33, width=0.25, max_ch=1024)
MEDIUM = dict(nc=1, depth=0.67, width=0.75, max_ch=768)


def _image(seed, h, w):
    """Low-frequency colour field + noise: no constant regions, so no two anchors see the same receptive field."""
    import cv2
    rng = np.random.default_rng(seed)
    low = rng.uniform(0, 255, size=(h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    low = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(low + rng.normal(0, 12, size=(h, w, 3)), 0, 255).astype(np.uint8)


def _setup(cfg, h, w, seed, imgsz):
    from mangatranslator_b200.yolo import YoloB200
    from mangatranslator_b200.preproc import letterbox_device
    m = Y.make_model(seed, bias_objects=-7.0, **cfg)
    img = _image(seed, h, w)
    dev = torch.device("cuda:0")
    net = YoloB200(m.state_dict(), m.cfg, dev)
    lb = letterbox_device(torch.from_numpy(img).to(dev), imgsz, swap_rb=True)
    g = net.forward_letterboxed(lb)
    torch.cuda.synchronize()
    x = Y.preprocess(img, imgsz)
    assert tuple(x.shape[2:]) == tuple(lb.shape[:2])
    return m, net, g, x, img, lb


def _well_posed_conf(m, x, lo=20, hi=60):
    """A confidence threshold that ~lo..hi anchors clear, placed in the widest score gap, and the decision margins of
    the oracle at that threshold (score gap at the cut, min gap between kept scores, min |IoU - 0.7|)."""
    with torch.no_grad():
        pred, _ = m(x)
    sc = torch.sort(pred[0, 4], descending=True).values
    gaps = sc[lo - 1:hi - 1] - sc[lo:hi]
    k = lo + int(torch.argmax(gaps))
    conf = float((sc[k - 1] + sc[k]) / 2)
    top = sc[:k]
    p = pred[0].t()
    b = p[p[:, 4] > conf]
    xyxy = torch.cat((b[:, :2] - b[:, 2:4] / 2, b[:, :2] + b[:, 2:4] / 2), 1)
    iou = Y.box_iou_matrix(xyxy)
    d = (iou - 0.7).abs()
    d.fill_diagonal_(1.0)
    if d.numel() == 0 or b.shape[0] < 2:
        return conf, 0.0, 0.0, 0.0
    return conf, float(gaps.max()), float((top[:-1] - top[1:]).min()), float(d.min())


@pytest.mark.parametrize("cfg,hw,imgsz", [(NANO, (200, 320), 320), (MEDIUM, (192, 160), 192)], ids=["nano", "medium"])
def test_head_outputs_match_oracle(cfg, hw, imgsz):
    m, net, g, x, img, lb = _setup(cfg, hw[0], hw[1], 3, imgsz)
    with torch.no_grad():
        raw, proto = m.heads_raw(x)
    # tolerance: 1e-3 relative to the tensor's magnitude (the synthetic weights drive intermediate activations to
    # |x| ~ 50-100, so a 1e-5 relative conv error is ~1e-3 absolute on O(1..10) head outputs)
    def close(got, ref):
        return (got - ref).abs().max().item() < 1e-3 * max(1.0, float(ref.abs().max()))
    for (box, cls, mc), (gb, gc, gm, fh, fw, st) in zip(raw, g["levels"]):
        assert close(gb.cpu()[0].permute(2, 0, 1), box[0])
        assert close(gc.cpu()[0, :, :, :1].permute(2, 0, 1), cls[0])
        assert close(gm.cpu()[0].permute(2, 0, 1), mc[0])
    assert close(g["proto"].cpu()[0].permute(2, 0, 1), proto[0])


@pytest.mark.parametrize("cfg,hw,imgsz", [(NANO, (200, 320), 320), (NANO, (333, 250), 320)], ids=["wide", "tall"])
def test_detections_match_oracle_and_reference_dedup(cfg, hw, imgsz):
    # index parity is only well-posed when the oracle's own decisions are not knife-edge: walk seeds until the
    # oracle's margins (score gap at the cut, gaps between kept scores, |IoU - 0.7|) are far above numerical noise
    for seed in range(5, 40):
        mm = Y.make_model(seed, bias_objects=-7.0, **cfg)
        xx = Y.preprocess(_image(seed, hw[0], hw[1]), imgsz)
        conf, cut_gap, min_gap, iou_margin = _well_posed_conf(mm, xx)
        if cut_gap > 1e-4 and min_gap > 5e-5 and iou_margin > 1e-3:
            break
    else:
        pytest.fail("no well-posed synthetic case found")
    m, net, g, x, img, lb = _setup(cfg, hw[0], hw[1], seed, imgsz)
    ref = Y.predict(m, img, conf, imgsz)
    det, cnt, final_idx = net.detect(g, conf, hw, tuple(lb.shape[:2]), apply_reference_dedup=True)
    torch.cuda.synchronize()
    n_nms, n_final = int(cnt[0]), int(cnt[1])
    assert n_nms == ref["xyxy"].shape[0] and n_nms > 3, (n_nms, ref["xyxy"].shape[0])
    d = det[:n_nms].cpu()
    assert torch.equal(d[:, 6].long(), ref["anchors"])          # bit-exact NMS indices (anchor ids, in score order)
    assert (d[:, :4] - ref["xyxy"]).abs().max().item() < 1e-2   # pixels
    assert (d[:, 4] - ref["conf"]).abs().max().item() < 1e-4
    # the reference's own post-NMS logic on the same boxes (restated below from detection.py:204-295)
    boxes = d[:, :4].tolist()
    confs = d[:, 4].tolist()
    keep = _ref_dedup(boxes, confs, 0.7)
    keep2 = _ref_remove_contained([boxes[i] for i in keep], 0.9)
    exp = [keep[i] for i in keep2]
    assert final_idx[:n_final].cpu().tolist() == exp


def _iou(a, b):
    ix = max(0.0, min(a[2], b[2]) - max(a[0], b[0])) * max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    ua = max(0.0, a[2] - a[0]) * max(0.0, a[3] - a[1]) + max(0.0, b[2] - b[0]) * max(0.0, b[3] - b[1]) - ix
    return ix / ua if ua > 0 else 0.0


def _ioa(inner, outer):
    ai = max(0.0, inner[2] - inner[0]) * max(0.0, inner[3] - inner[1])
    if ai <= 0:
        return 0.0
    ix = max(0.0, min(inner[2], outer[2]) - max(inner[0], outer[0])) * max(0.0, min(inner[3], outer[3]) - max(inner[1], outer[1]))
    return ix / ai


def _ref_dedup(boxes, confs, thr):
    order = sorted(range(len(boxes)), key=lambda i: confs[i], reverse=True)
    keep = []
    for i in order:
        if not any(_iou(boxes[i], boxes[k]) > thr for k in keep):
            keep.append(i)
    return keep


def _ref_remove_contained(boxes, thr):
    n = len(boxes)
    alive = [True] * n
    for i in range(n):
        if not alive[i]:
            continue
        for j in range(n):
            if i == j or not alive[j]:
                continue
            if _ioa(boxes[i], boxes[j]) > thr:
                alive[i] = False
                break
    return [i for i in range(n) if alive[i]]


def test_post_nms_logic_matches_live_reference():
    """When the reference tree is present, its _deduplicate_primary_boxes/_remove_contained_boxes agree with the
    restatement used above (so the GPU indices are pinned to the reference itself)."""
    import _refimport
    if not _refimport.available():
        pytest.skip("reference tree not present")
    core = _refimport.import_reference()
    import core.image.detection as ref
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(2, 25))
        xy = rng.uniform(0, 500, size=(n, 2))
        wh = rng.uniform(20, 200, size=(n, 2))
        b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        b[n // 2] = b[0] + rng.uniform(-3, 3, size=4).astype(np.float32)       # near duplicate
        c = rng.uniform(0.3, 1.0, size=n).astype(np.float32)
        tb, tc = torch.from_numpy(b), torch.from_numpy(c)
        kb, keep = ref._deduplicate_primary_boxes(tb, tc, 0.7)
        assert keep == _ref_dedup(tb.tolist(), tc.tolist(), 0.7)
        kb2, idx = ref._remove_contained_boxes(kb, [("primary", i) for i in keep], 0.9)
        assert [i for (_, i) in idx] == [keep[i] for i in _ref_remove_contained(kb.tolist(), 0.9)]


def test_call_shape_and_retina_masks_match_oracle():
    """`model(image_bgr, conf=..., imgsz=..., retina_masks=True)[0]` like the reference calls it (detection.py:1338-1345):
    boxes, scores and the native-resolution masks (process_mask_native) against the oracle."""
    from mangatranslator_b200.yolo import YoloB200
    cfg, hw, imgsz = NANO, (333, 250), 320
    for seed in range(5, 40):
        mm = Y.make_model(seed, bias_objects=-7.0, **cfg)
        img = _image(seed, hw[0], hw[1])
        conf, cut_gap, min_gap, iou_margin = _well_posed_conf(mm, Y.preprocess(img, imgsz))
        if cut_gap > 1e-4 and min_gap > 5e-5 and iou_margin > 1e-3:
            break
    ref = Y.predict(mm, img, conf, imgsz)
    net = YoloB200(mm.state_dict(), mm.cfg, torch.device("cuda:0"))
    res = net(np.ascontiguousarray(img), conf=conf, device=None, verbose=False, imgsz=imgsz, retina_masks=True)[0]
    assert res.orig_shape == hw and len(res.boxes) == ref["xyxy"].shape[0] == len(res.masks)
    assert (res.boxes.xyxy.cpu() - ref["xyxy"]).abs().max().item() < 1e-2
    got = res.masks.data.cpu().numpy() > 0.5
    exp = ref["masks"].numpy()
    assert got.shape == exp.shape
    assert (got != exp).mean() < 1e-3          # knife-edge pixels of the `> 0` decision only
    assert exp.any()


def test_batched_plan_equals_page_by_page():
    """Pages of one size through ONE plan with a batch dimension (`forward_letterboxed_batch` / `detect_batch`, what
    `detect_pages_device` does for a group): head tensors, detection tables, counts and the reference-dedup index lists
    are bit-identical to running the pages one by one."""
    from mangatranslator_b200.preproc import letterbox_device
    from mangatranslator_b200.yolo import YoloB200
    cfg, hw, imgsz = NANO, (200, 320), 320
    m = Y.make_model(4, bias_objects=-3.0, **cfg)
    dev = torch.device("cuda:0")
    net = YoloB200(m.state_dict(), m.cfg, dev)
    imgs = [_image(20 + i, *hw) for i in range(3)]
    lbs = [letterbox_device(torch.from_numpy(im).to(dev), imgsz, swap_rb=True) for im in imgs]
    single = []
    for lb in lbs:
        g = net.forward_letterboxed(lb)
        det, cnt, fin = net.detect(g, 0.3, hw, tuple(lb.shape[:2]), apply_reference_dedup=True)
        single.append((det.clone(), cnt.clone(), fin.clone(), [lv[0].clone() for lv in g["levels"]]))
    gb = net.forward_letterboxed_batch(lbs)
    det, cnt, fin = net.detect_batch(gb, 0.3, hw, tuple(lbs[0].shape[:2]), apply_reference_dedup=True)
    torch.cuda.synchronize()
    assert int(cnt[:, 0].min()) > 0
    for i, (d1, c1, f1, heads) in enumerate(single):
        for lv, h1 in zip(gb["levels"], heads):
            assert torch.equal(lv[0][i], h1[0])
        n, nf = int(c1[0]), int(c1[1])
        assert torch.equal(cnt[i], c1)
        assert torch.equal(det[i, :n], d1[:n]) and torch.equal(fin[i, :nf], f1[:nf])
