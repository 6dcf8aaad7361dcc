"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle of the YOLOv8-seg speech-bubble detector call.

The reference delegates to a third-party package that is absent here: `ultralytics>=8.3.94` (requirements.txt:21);
call sites core/ml/model_manager.py:711-743 (`YOLO(path)`) and core/image/detection.py:1338-1345
(`model(image_bgr, conf=..., imgsz=1600|640, retina_masks=True)[0]`).  This file restates, from the published
ultralytics sources (nn/modules/{conv,block,head}.py, data/augment.py LetterBox, utils/ops.py non_max_suppression /
scale_boxes / process_mask_native), the computation of that call for a YOLOv8-seg model (`yolo_1` =
kitsumed/yolov8m_seg-speech-bubble is YOLOv8m-seg; the default `yolo_2` checkpoint's family is unknown offline).
BatchNorm is assumed already folded into the convolutions (ultralytics fuses at first predict), so a Conv block is
conv(+bias)+SiLU.  PARITY UNPINNED: ultralytics and the checkpoints are not installable here; the restatement is
checked for internal consistency only and is the yard-stick for the CUDA path on seeded random weights.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _c(x: float, width: float, max_ch: int) -> int:
    return int(math.ceil(min(x, max_ch) * width / 8) * 8)


class Conv(nn.Module):
    """Conv2d(bias, BN folded) + SiLU, pad = k // 2."""

    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=True)
        self.act = act

    def forward(self, x):
        y = self.conv(x)
        return F.silu(y) if self.act else y


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut):
        super().__init__()
        self.cv1, self.cv2, self.add = Conv(c, c, 3), Conv(c, c, 3), shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n, shortcut):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        self.cv1, self.cv2, self.k = Conv(c1, c1 // 2, 1), Conv(c1 // 2 * 4, c2, 1), k

    def forward(self, x):
        x = self.cv1(x)
        y1 = F.max_pool2d(x, self.k, 1, self.k // 2)
        y2 = F.max_pool2d(y1, self.k, 1, self.k // 2)
        y3 = F.max_pool2d(y2, self.k, 1, self.k // 2)
        return self.cv2(torch.cat((x, y1, y2, y3), 1))


class Proto(nn.Module):
    def __init__(self, c1, c_, c2):
        super().__init__()
        self.cv1 = Conv(c1, c_, 3)
        self.upsample = nn.ConvTranspose2d(c_, c_, 2, 2, 0, bias=True)
        self.cv2 = Conv(c_, c_, 3)
        self.cv3 = Conv(c_, c2, 1)

    def forward(self, x):
        return self.cv3(self.cv2(self.upsample(self.cv1(x))))


class Segment(nn.Module):
    """Detect + mask-coefficient branches + Proto (head.py Segment)."""

    def __init__(self, nc, nm, npr, ch):
        super().__init__()
        self.nc, self.nm, self.reg_max = nc, nm, 16
        c2, c3 = max(16, ch[0] // 4, self.reg_max * 4), max(ch[0], min(nc, 100))
        c4 = max(ch[0] // 4, nm)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, nm, 1)) for x in ch)
        self.proto = Proto(ch[0], npr, nm)


class YoloV8Seg(nn.Module):
    """YOLOv8-seg graph (ultralytics cfg/models/v8/yolov8-seg.yaml).  depth/width/max_ch: m = (0.67, 0.75, 768)."""

    def __init__(self, nc=1, depth=0.67, width=0.75, max_ch=768, nm=32, npr=256):
        super().__init__()
        d = lambda n: max(round(n * depth), 1)
        c = lambda x: _c(x, width, max_ch)
        self.cfg = dict(nc=nc, depth=depth, width=width, max_ch=max_ch, nm=nm, npr=npr)
        self.l0 = Conv(3, c(64), 3, 2)
        self.l1 = Conv(c(64), c(128), 3, 2)
        self.l2 = C2f(c(128), c(128), d(3), True)
        self.l3 = Conv(c(128), c(256), 3, 2)
        self.l4 = C2f(c(256), c(256), d(6), True)
        self.l5 = Conv(c(256), c(512), 3, 2)
        self.l6 = C2f(c(512), c(512), d(6), True)
        self.l7 = Conv(c(512), c(1024), 3, 2)
        self.l8 = C2f(c(1024), c(1024), d(3), True)
        self.l9 = SPPF(c(1024), c(1024), 5)
        self.l12 = C2f(c(1024) + c(512), c(512), d(3), False)
        self.l15 = C2f(c(512) + c(256), c(256), d(3), False)
        self.l16 = Conv(c(256), c(256), 3, 2)
        self.l18 = C2f(c(256) + c(512), c(512), d(3), False)
        self.l19 = Conv(c(512), c(512), 3, 2)
        self.l21 = C2f(c(512) + c(1024), c(1024), d(3), False)
        self.head = Segment(nc, nm, c(npr), (c(256), c(512), c(1024)))
        self.strides = (8, 16, 32)

    def features(self, x):
        x = self.l1(self.l0(x))
        x = self.l2(x)
        p3 = self.l4(self.l3(x))
        p4 = self.l6(self.l5(p3))
        p5 = self.l9(self.l8(self.l7(p4)))
        n4 = self.l12(torch.cat([F.interpolate(p5, scale_factor=2, mode="nearest"), p4], 1))
        n3 = self.l15(torch.cat([F.interpolate(n4, scale_factor=2, mode="nearest"), p3], 1))
        m4 = self.l18(torch.cat([self.l16(n3), n4], 1))
        m5 = self.l21(torch.cat([self.l19(m4), p5], 1))
        return [n3, m4, m5]

    def heads_raw(self, x):
        """Per level (box logits [B,64,h,w], class logits [B,nc,h,w], mask coeffs [B,nm,h,w]) and proto."""
        feats = self.features(x)
        h = self.head
        return [(h.cv2[i](f), h.cv3[i](f), h.cv4[i](f)) for i, f in enumerate(feats)], h.proto(feats[0])

    def forward(self, x):
        """-> (pred [B, 4+nc+nm, A] with xywh boxes in input pixels and sigmoid scores, proto [B, nm, h/4, w/4])."""
        feats = self.features(x)
        h = self.head
        proto = h.proto(feats[0])
        box, cls, mc, anchors, strides = [], [], [], [], []
        for i, f in enumerate(feats):
            b, _, fh, fw = f.shape
            box.append(h.cv2[i](f).view(b, 64, -1))
            cls.append(h.cv3[i](f).view(b, h.nc, -1))
            mc.append(h.cv4[i](f).view(b, h.nm, -1))
            sx = torch.arange(fw, dtype=torch.float32, device=f.device) + 0.5
            sy = torch.arange(fh, dtype=torch.float32, device=f.device) + 0.5
            yy, xx = torch.meshgrid(sy, sx, indexing="ij")
            anchors.append(torch.stack((xx, yy), -1).view(-1, 2))
            strides.append(torch.full((fh * fw, 1), float(self.strides[i]), device=f.device))
        box, cls, mc = torch.cat(box, 2), torch.cat(cls, 2), torch.cat(mc, 2)
        anchors, strides = torch.cat(anchors).t().unsqueeze(0), torch.cat(strides).t()
        b, _, a = box.shape
        dist = box.view(b, 4, 16, a).transpose(2, 1).softmax(1)            # DFL: softmax over 16 bins
        dist = (dist * torch.arange(16, dtype=torch.float32, device=box.device).view(1, 16, 1, 1)).sum(1)
        lt, rb = dist.chunk(2, 1)
        x1y1, x2y2 = anchors - lt, anchors + rb
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides      # xywh
        return torch.cat((dbox, cls.sigmoid(), mc), 1), proto


# ---- predictor pre/post-processing ------------------------------------------------------------------------
def letterbox_params(h0: int, w0: int, imgsz: int, stride: int = 32):
    """LetterBox(auto=True, scaleup=True): returns (new_unpad (w,h), (top,bottom,left,right), out (h,w), gain)."""
    r = min(imgsz / h0, imgsz / w0)
    nw, nh = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = (imgsz - nw) % stride, (imgsz - nh) % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return (nw, nh), (top, bottom, left, right), (nh + top + bottom, nw + left + right), r


def letterbox(img_bgr: np.ndarray, imgsz: int):
    import cv2
    h0, w0 = img_bgr.shape[:2]
    (nw, nh), (t, b, l, r), _, _ = letterbox_params(h0, w0, imgsz)
    im = img_bgr
    if (w0, h0) != (nw, nh):
        im = cv2.resize(im, (nw, nh), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(im, t, b, l, r, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def preprocess(img_bgr: np.ndarray, imgsz: int) -> torch.Tensor:
    lb = letterbox(img_bgr, imgsz)
    x = np.ascontiguousarray(lb[..., ::-1].transpose(2, 0, 1))          # BGR->RGB, HWC->CHW
    return torch.from_numpy(x).float().div(255.0).unsqueeze(0)


def box_iou_matrix(b: torch.Tensor) -> torch.Tensor:
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(b[:, None, :2], b[None, :, :2])
    rb = torch.min(b[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area[:, None] + area[None, :] - inter)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thr: float) -> torch.Tensor:
    """torchvision.ops.nms semantics: greedy in descending score order, suppress IoU > thr."""
    order = torch.argsort(scores, descending=True, stable=True)
    iou = box_iou_matrix(boxes[order])
    keep, dead = [], torch.zeros(len(order), dtype=torch.bool)
    for i in range(len(order)):
        if dead[i]:
            continue
        keep.append(int(order[i]))
        dead |= iou[i] > iou_thr
    return torch.tensor(keep, dtype=torch.long)


def non_max_suppression(pred: torch.Tensor, nc: int, conf_thres: float, iou_thres: float = 0.7, max_det: int = 300,
                        max_nms: int = 30000, max_wh: int = 7680):
    """utils/ops.py non_max_suppression for one image, single-label.  pred [4+nc+nm, A] -> [n, 6+nm]."""
    p = pred.t()
    scores, cls = p[:, 4:4 + nc].max(1)
    sel = scores > conf_thres
    p, scores, cls = p[sel], scores[sel], cls[sel]
    if p.shape[0] == 0:
        return torch.zeros((0, 6 + p.shape[1] - 4 - nc)), torch.zeros((0,), dtype=torch.long)
    anchors = torch.nonzero(sel)[:, 0]
    xy, wh = p[:, :2], p[:, 2:4]
    xyxy = torch.cat((xy - wh / 2, xy + wh / 2), 1)
    x = torch.cat((xyxy, scores[:, None], cls[:, None].float(), p[:, 4 + nc:]), 1)
    order = x[:, 4].argsort(descending=True, stable=True)[:max_nms]
    x, anchors = x[order], anchors[order]
    off = x[:, 5:6] * max_wh
    keep = nms(x[:, :4] + off, x[:, 4], iou_thres)[:max_det]
    return x[keep], anchors[keep]


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape) -> torch.Tensor:
    """utils/ops.py scale_boxes: letterboxed -> original pixels, clipped."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    b = boxes.clone()
    b[:, [0, 2]] -= pad_x
    b[:, [1, 3]] -= pad_y
    b /= gain
    b[:, [0, 2]] = b[:, [0, 2]].clamp(0, img0_shape[1])
    b[:, [1, 3]] = b[:, [1, 3]].clamp(0, img0_shape[0])
    return b


def process_mask_native(proto: torch.Tensor, coeffs: torch.Tensor, boxes: torch.Tensor, img1_shape, img0_shape):
    """utils/ops.py process_mask_native: coeff @ proto -> strip letterbox pad -> bilinear to the original size ->
    zero outside the box -> > 0."""
    c, mh, mw = proto.shape
    masks = (coeffs @ proto.view(c, -1)).view(-1, mh, mw)
    gain = min(mh / img0_shape[0], mw / img0_shape[1])
    pad_w, pad_h = (mw - img0_shape[1] * gain) / 2, (mh - img0_shape[0] * gain) / 2
    top, left = int(round(pad_h - 0.1)), int(round(pad_w - 0.1))
    bottom, right = mh - int(round(pad_h + 0.1)), mw - int(round(pad_w + 0.1))
    masks = masks[:, top:bottom, left:right]
    masks = F.interpolate(masks[None], tuple(img0_shape), mode="bilinear", align_corners=False)[0]
    h, w = img0_shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, dtype=boxes.dtype)[None, None, :]
    cc = torch.arange(h, dtype=boxes.dtype)[None, :, None]
    masks = masks * ((r >= x1) * (r < x2) * (cc >= y1) * (cc < y2))
    return masks.gt(0.0)


@torch.no_grad()
def predict(model: YoloV8Seg, img_bgr: np.ndarray, conf: float, imgsz: int, retina_masks: bool = True):
    """-> dict(xyxy [n,4] original px, conf [n], cls [n], masks [n,H,W] bool, anchors [n])."""
    h0, w0 = img_bgr.shape[:2]
    x = preprocess(img_bgr, imgsz)
    pred, proto = model(x)
    det, anchors = non_max_suppression(pred[0], model.cfg["nc"], conf)
    if det.shape[0] == 0:
        return dict(xyxy=torch.zeros((0, 4)), conf=torch.zeros(0), cls=torch.zeros(0),
                    masks=torch.zeros((0, h0, w0), dtype=torch.bool), anchors=anchors, raw=pred, proto=proto)
    boxes = scale_boxes(x.shape[2:], det[:, :4], (h0, w0))
    masks = process_mask_native(proto[0], det[:, 6:], boxes, x.shape[2:], (h0, w0))
    return dict(xyxy=boxes, conf=det[:, 4], cls=det[:, 5], masks=masks, anchors=anchors, raw=pred, proto=proto)


def make_model(seed: int = 0, bias_objects: float = 0.0, cls_gain: float = 1.0, signal_init: bool = True,
               **cfg) -> YoloV8Seg:
    """Seeded random weights (no checkpoint offline).

    PyTorch's default conv init shrinks the signal at every layer, so a 60-layer random network's outputs are almost
    input-independent (every anchor gets the same score and NMS degenerates into tie-breaking).  `signal_init` draws
    variance-preserving weights instead (N(0, 2.6/fan_in), small random biases) so activations stay O(1) and scores
    differ from anchor to anchor.  `bias_objects` / `cls_gain` shift and scale the class logits so a controllable
    fraction of anchors clears the confidence threshold with margins far above the numerical noise."""
    torch.manual_seed(seed)
    m = YoloV8Seg(**cfg).eval()
    with torch.no_grad():
        if signal_init:
            for mod in m.modules():
                if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
                    fan_in = mod.weight[0].numel() if isinstance(mod, nn.Conv2d) else mod.weight.shape[0]
                    mod.weight.normal_(0.0, (2.6 / fan_in) ** 0.5)
                    if mod.bias is not None:
                        mod.bias.normal_(0.0, 0.1)
        if signal_init:
            # calibrate the last 1x1 conv of every head branch on a probe image so the head outputs are O(1)
            g = torch.Generator().manual_seed(seed + 1)
            probe = torch.rand((1, 3, 160, 160), generator=g)
            raw, _ = m.heads_raw(probe)
            for i, (box, cls, mc) in enumerate(raw):
                for seq, out, target in ((m.head.cv2[i], box, 2.0), (m.head.cv3[i], cls, 1.5), (m.head.cv4[i], mc, 1.0)):
                    sd = float(out.std())
                    if sd > 0:
                        seq[-1].weight.mul_(target / sd)
                        seq[-1].bias.mul_(target / sd)
        for seq in m.head.cv3:
            seq[-1].weight.mul_(cls_gain)
            seq[-1].bias.fill_(bias_objects)
    return m
