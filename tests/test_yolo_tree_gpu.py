"""GPU: the module-tree detectors (panel YOLO11 / OSB-text YOLO12, mangatranslator_b200/yolo_tree.py) against the CPU
oracle of the same tree (oracle/yolo_tree_oracle.py): head tensors, NMS indices bit-exact, boxes and scores; the depthwise
convolution kernel against torch; `detect_panels` / OSB-text expansion through the model manager."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import yolo_oracle as Y
import yolo_tree_oracle as O

pytestmark = pytest.mark.gpu

FAMILIES = [("11", dict()), ("12", dict(a2_residual=True, mlp_ratio=1.2))]


def _image(seed, h, w):
    import cv2
    rng = np.random.default_rng(seed)
    low = rng.uniform(0, 255, size=(h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    low = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(low + rng.normal(0, 12, size=(h, w, 3)), 0, 255).astype(np.uint8)


def _tree(family, kw, seed, img, imgsz, nc=3):
    from mangatranslator_b200 import yolo_tree as T
    tree = T.synthetic_tree(family, "s", nc=nc, seed=seed, names={0: "body", 1: "frame", 2: "text"}, **kw)
    O.calibrate(tree, img, imgsz, cls_mean=-6.0, cls_std=1.2)
    return tree


def _logit(p):
    return torch.log(p) - torch.log1p(-p)


def _well_posed_conf(pred, lo=6, hi=16):
    """A confidence threshold that lo..hi anchors clear, in the widest gap of the class LOGITS, and the oracle's decision
    margins there: logit gap at the cut, min logit gap between kept anchors (their order is the NMS order), min
    |IoU - 0.7| between kept boxes, min probability margin of the winning class."""
    sc_all = pred[0, 4:].max(0).values
    sc = torch.sort(sc_all, descending=True).values
    lg = _logit(sc.double().clamp(1e-12, 1 - 1e-12))
    gaps = lg[lo - 1:hi - 1] - lg[lo:hi]
    k = lo + int(torch.argmax(gaps))
    conf = float((sc[k - 1] + sc[k]) / 2)
    top = lg[:k]
    p = pred[0].t()
    keep = p[:, 4:].max(1).values > conf
    b, cls = p[keep], p[keep][:, 4:].argmax(1)
    xyxy = torch.cat((b[:, :2] - b[:, 2:4] / 2, b[:, :2] + b[:, 2:4] / 2), 1) + cls[:, None].float() * 7680
    d = (Y.box_iou_matrix(xyxy) - 0.7).abs()
    d.fill_diagonal_(1.0)
    two = torch.sort(b[:, 4:], 1, descending=True).values
    cls_margin = float((two[:, 0] - two[:, 1]).min()) if two.shape[1] > 1 else 1.0
    return conf, float(gaps.max()), float((top[:-1] - top[1:]).min()), float(d.min()), cls_margin


@pytest.mark.parametrize("family,kw", FAMILIES, ids=["yolo11", "yolo12"])
def test_heads_and_detections_match_oracle(family, kw):
    from mangatranslator_b200.preproc import letterbox_device
    from mangatranslator_b200.yolo_tree import YoloTreeB200
    hw, imgsz = (300, 420), 448
    dev = torch.device("cuda:0")
    for seed in range(3, 30):
        img = _image(seed, *hw)
        tree = _tree(family, kw, seed, img, imgsz)
        x = Y.preprocess(img, imgsz)
        pred, heads = O.forward(tree, x)
        conf, cut_gap, min_gap, iou_margin, cls_margin = _well_posed_conf(pred)
        # margins far above the head error of these seeded networks (a few 1e-2 in logits, a few 0.1 px in boxes: see
        # profiles/r02_yolo_tree_layer_errors.txt — a random deep network amplifies the 1e-5 operand rounding of its first
        # layer by 1.1-2x per block; a trained one does not)
        if cut_gap > 0.2 and min_gap > 0.05 and iou_margin > 0.03 and cls_margin > 0.02:
            break
    else:
        pytest.fail("no well-posed synthetic case found")
    net = YoloTreeB200(tree, dev)
    lb = letterbox_device(torch.from_numpy(img).to(dev), imgsz, swap_rb=True)
    assert tuple(x.shape[2:]) == tuple(lb.shape[:2])
    g = net.forward_letterboxed(lb)
    torch.cuda.synchronize()
    worst = 0.0
    for (box, cls), (gb, gc, _, fh, fw, st) in zip(heads, g["levels"]):
        for got, ref in ((gb.cpu()[0].permute(2, 0, 1), box[0]), (gc.cpu()[0, :, :, :net.nc].permute(2, 0, 1), cls[0])):
            err = (got - ref).abs().max().item()
            worst = max(worst, err)
            assert err < max(1e-3, 4e-3 * float(ref.abs().max())), (family, st, err)      # measured: 1.0e-3 / 3e-4 relative
    print(f"yolo{family}: head tensors max abs err {worst:.2e}")
    ref = O.predict(tree, img, conf, imgsz)
    det, cnt, _ = net.detect(g, conf, hw, tuple(lb.shape[:2]), apply_reference_dedup=False)
    torch.cuda.synchronize()
    n = int(cnt[0])
    assert n == ref["xyxy"].shape[0] and n > 3, (n, ref["xyxy"].shape[0])
    d = det[:n].cpu()
    assert torch.equal(d[:, 6].long(), ref["anchors"])            # bit-exact NMS indices (anchor ids, in score order)
    assert torch.equal(d[:, 5].long(), ref["cls"].long())
    box_err, conf_err = (d[:, :4] - ref["xyxy"]).abs().max().item(), (d[:, 4] - ref["conf"]).abs().max().item()
    print(f"yolo{family}: {n} detections, boxes max err {box_err:.3f} px, scores max err {conf_err:.2e}")
    assert box_err < 0.5 and conf_err < 5e-3
    # the reference call shape gives the same rows
    out = net(img, conf=conf, imgsz=imgsz)[0]
    assert len(out.boxes) == n and torch.equal(out.boxes.xyxy.cpu(), d[:, :4]) and out.masks is None
    assert out.orig_shape == hw and net.names[1] == "frame"
    # a second, differently shaped page through the same object (its own plan), then the first again: bit-identical
    other = _image(99, 260, 200)
    net(other, conf=conf, imgsz=imgsz)
    again = net(img, conf=conf, imgsz=imgsz)[0]
    assert torch.equal(again.boxes.xyxy, out.boxes.xyxy) and torch.equal(again.boxes.conf, out.boxes.conf)


def test_segment_head_serves_as_the_speech_bubble_detector(tmp_path):
    """A YOLO11-seg checkpoint (what the default `yolo_2` file may be) loads through `load_yolo_speech_bubble`: refused by
    the hard-wired YOLOv8-seg reader, executed from its module tree, with mask coefficients, prototypes and retina masks
    (`process_mask_native`) against the oracle, and accepted by the device-resident page path."""
    import sys
    from mangatranslator_b200 import yolo_tree as T
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    from mangatranslator_b200.preproc import letterbox_device
    from mangatranslator_b200.yolo import YoloB200
    from test_yolo_tree_host import _save_with_fake_package
    hw, imgsz = (300, 420), 448
    for seed in range(3, 40):
        img = _image(seed, *hw)
        tree = T.synthetic_tree("11", "s", nc=1, seed=seed, names={0: "bubble"}, segment=True)
        O.calibrate(tree, img, imgsz, cls_mean=-6.0, cls_std=1.2)
        pred, heads, seg = O.forward(tree, Y.preprocess(img, imgsz))
        conf, cut_gap, min_gap, iou_margin, _ = _well_posed_conf(pred[:, :5])
        if cut_gap > 0.2 and min_gap > 0.05 and iou_margin > 0.03:
            break
    else:
        pytest.fail("no well-posed synthetic case found")
    path, stash = _save_with_fake_package(tmp_path, tree)
    sys.modules.update(stash)
    mm = get_model_manager()
    mm.unload_model(ModelType.YOLO_SPEECH_BUBBLE)
    try:
        net = mm.load_yolo_speech_bubble(str(path))
        assert isinstance(net, T.YoloTreeB200) and isinstance(net, YoloB200) and net.has_masks and net.names == {0: "bubble"}
        dev = net.device
        lb = letterbox_device(torch.from_numpy(img).to(dev), imgsz, swap_rb=True)
        g = net.forward_letterboxed(lb)
        torch.cuda.synchronize()
        for (gb, gc, gm, fh, fw, st), mc in zip(g["levels"], seg[0]):
            err = (gm.cpu()[0].permute(2, 0, 1) - mc[0]).abs().max().item()
            assert err < max(1e-3, 4e-3 * float(mc.abs().max())), ("mask coefficients", st, err)
        perr = (g["proto"].cpu()[0].permute(2, 0, 1) - seg[1][0]).abs().max().item()
        assert perr < max(1e-3, 4e-3 * float(seg[1].abs().max())), ("prototypes", perr)
        ref = O.predict(tree, img, conf, imgsz)
        out = net(img, conf=conf, imgsz=imgsz, retina_masks=True)[0]
        n = len(out.boxes)
        assert n == len(ref["conf"]) and n > 2
        assert (out.boxes.xyxy.cpu() - ref["xyxy"]).abs().max().item() < 0.5
        got, want = out.masks.data.cpu() > 0, ref["masks"]
        assert tuple(got.shape) == tuple(want.shape) == (n, hw[0], hw[1])
        flips = (got != want).flatten(1).sum(1)
        area = want.flatten(1).sum(1).clamp_min(1)
        print(f"yolo11-seg: {n} detections, prototypes max err {perr:.2e}, flipped mask pixels per detection "
              f"{flips.tolist()} of {area.tolist()}")
        assert int(want.sum()) > 0 and torch.all(flips <= torch.clamp(0.01 * area, min=30)), (flips, area)
        # the device-resident page path takes the tree model like the hard-wired one
        det, cnt, final_idx = net.detect(g, conf, hw, tuple(lb.shape[:2]), apply_reference_dedup=True)
        assert int(cnt[0]) == n and 0 < int(cnt[1]) <= n
    finally:
        mm.unload_model(ModelType.YOLO_SPEECH_BUBBLE)


@pytest.mark.parametrize("k,act,planes", [(3, 1, 2), (7, 0, 2), (3, 0, 1)])
def test_depthwise_conv_kernel_matches_torch(k, act, planes):
    from mangatranslator_b200 import planes as P
    from mangatranslator_b200._lib import check, lib, ptr, stream_ptr
    from mangatranslator_b200.yolo_tree import _declare
    l = lib()
    _declare(l)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(k * 10 + act)
    n, h, w, ct_in, ci, c, ct_out, co = 2, 19, 23, 48, 16, 32, 64, 8
    x = torch.randn((n, h, w, ct_in), generator=g)
    wt = torch.randn((c, 1, k, k), generator=g) * 0.3
    b = torch.randn((c,), generator=g) * 0.2
    xp = P.split_planes(x.to(dev), planes)                                      # [planes][n][h][w][ct_in]
    xq = xp.float().sum(0)                                                     # what the planes actually hold
    ref = F.conv2d(xq[..., ci:ci + c].permute(0, 3, 1, 2).cpu(), wt, b, padding=k // 2, groups=c)
    ref = F.silu(ref) if act else ref
    y = torch.full((planes, n, h, w, ct_out), 7.0, dtype=torch.bfloat16, device=dev)
    wd = wt.reshape(c, k * k).t().contiguous().to(dev)
    check(l.mtb_dwconv(ptr(xp), ptr(y), n, h, w, ct_in, ci, ct_out, co, c, k, ptr(wd), ptr(b.to(dev)), act, planes,
                       stream_ptr()), "mtb_dwconv")
    torch.cuda.synchronize()
    got = y.float().sum(0)[..., co:co + c].permute(0, 3, 1, 2).cpu()
    tol = 2e-5 if planes == 2 else 1.2e-2 * float(ref.abs().max())            # one bf16 plane: 8 significant bits out
    assert (got - ref).abs().max().item() < tol * max(1.0, float(ref.abs().max()) if planes == 2 else 1.0)
    untouched = torch.ones(ct_out, dtype=torch.bool)
    untouched[co:co + c] = False
    assert torch.all(y[..., untouched].float() == 7.0)                         # neighbours of the slice are not written
    # aliasing and odd channel counts are refused, not mangled
    assert l.mtb_dwconv(ptr(xp), ptr(xp), n, h, w, ct_in, ci, ct_in, ci, c, k, ptr(wd), ptr(b.to(dev)), act, planes,
                        stream_ptr()) != 0
    assert l.mtb_dwconv(ptr(xp), ptr(y), n, h, w, ct_in, ci, ct_out, co, 12, k, ptr(wd), ptr(b.to(dev)), act, planes,
                        stream_ptr()) != 0


def test_panels_and_osb_text_through_the_model_manager(monkeypatch):
    """`detect_panels` (:1817-1921) and the OSB-text expansion inside `detect_speech_bubbles` (:1555-1567) with the
    loaders' seeded synthetic trees: frames only, int tuples, cached second call; without the explicit opt-in the
    loaders refuse to invent weights."""
    from PIL import Image
    from mangatranslator_b200.core.image.detection import _expand_boxes_with_osb_text, detect_panels
    from mangatranslator_b200.core.caching import get_cache
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    from mangatranslator_b200.utils.exceptions import ModelError
    mm = get_model_manager()
    for t in (ModelType.YOLO_PANEL, ModelType.YOLO_OSBTEXT):
        mm.unload_model(t)
    monkeypatch.delenv("MTB200_SYNTHETIC_PANEL", raising=False)
    monkeypatch.delenv("MTB200_SYNTHETIC_OSBTEXT", raising=False)
    img = Image.fromarray(_image(4, 420, 300)[:, :, ::-1].copy())
    with pytest.raises(ModelError, match="no checkpoint"):
        detect_panels(None, 0.25, image_override=img)
    with pytest.raises(ModelError, match="no checkpoint"):
        mm.load_yolo_osbtext()
    monkeypatch.setenv("MTB200_SYNTHETIC_PANEL", "1")
    monkeypatch.setenv("MTB200_SYNTHETIC_OSBTEXT", "1")
    try:
        panels = detect_panels(None, 0.25, image_override=img)
        model = mm.load_yolo_panel()
        assert model.names[2] == "frame" and len(model.tree["layers"]) == 24          # the YOLO11-L layout
        res = model(np.ascontiguousarray(np.asarray(img)[:, :, ::-1]), conf=0.25, imgsz=640)[0]
        want = [] if res.boxes is None else [tuple(int(round(v)) for v in b) for b, c in
                                             zip(res.boxes.xyxy.cpu().tolist(), res.boxes.cls.cpu().tolist()) if int(c) == 2]
        assert panels == want and all(isinstance(v, int) for p in panels for v in p)
        osb = mm.load_yolo_osbtext()
        assert len(osb.tree["layers"]) == 22 and osb.nc == 1                            # the YOLO12x layout
        cv = np.ascontiguousarray(np.asarray(img)[:, :, ::-1])
        boxes = torch.tensor([[40.0, 60.0, 200.0, 260.0], [150.0, 300.0, 290.0, 400.0]])
        get_cache().clear_all()
        out = _expand_boxes_with_osb_text(cv, img, boxes.clone(), get_cache(), mm, mm.device, 0.25, "", True)
        assert out.shape == boxes.shape and torch.all(out[:, :2] <= boxes[:, :2]) and torch.all(out[:, 2:] >= boxes[:, 2:])
    finally:
        for t in (ModelType.YOLO_PANEL, ModelType.YOLO_OSBTEXT):
            mm.unload_model(t)
