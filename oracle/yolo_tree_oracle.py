"""TEST INFRASTRUCTURE ONLY — CPU (torch fp32) statement of the ultralytics blocks that the panel detector (YOLO11,
/root/reference/core/image/detection.py:1817-1921) and the OSB-text detector (YOLO12, detection.py:120-201) are made of,
evaluated on the node tree of mangatranslator_b200/yolo_tree.py.  PARITY UNPINNED: `ultralytics` (requirements.txt:21,
>= 8.3.94) is not installed here and there are no checkpoints offline, so each block's forward rule below is restated
from memory of ultralytics nn/modules/{conv,block,head}.py:

  Conv        act(bn(conv(x))), autopad k // 2, SiLU                       (BatchNorm already folded in the tree)
  Bottleneck  x + cv2(cv1(x)) if add else cv2(cv1(x))
  C2f / C3k2  y = cv1(x).chunk(2); y += [m(y[-1]) for m in self.m]; cv2(cat(y))       (m: Bottleneck or C3k)
  C3 / C3k    cv3(cat(m(cv1(x)), cv2(x)))
  SPPF        y = [cv1(x)]; y += [maxpool5(y[-1])] * 3; cv2(cat(y))
  C2PSA       a, b = cv1(x).split(c); b = m(b); cv2(cat(a, b));   PSABlock: x = x + attn(x); x = x + ffn(x)
  Attention   qkv(x).view(B, heads, 2 kd + hd, N).split(kd, kd, hd); softmax(q^T k * kd^-0.5); v attn^T + pe(v); proj
  A2C2f       y = [cv1(x)]; y += [m(y[-1]) ...]; y = cv2(cat(y)); x + gamma * y if gamma is not None else y
  ABlock      x = x + attn(x); x = x + mlp(x);   AAttn: per-head [q k v] of hd each, tokens split into `area` contiguous
              row bands, softmax(q^T k * hd^-0.5), + pe(v) (7x7 depthwise), proj
  Detect      per level: cat(cv2(x), cv3(x)); DFL (softmax over 16 bins, expectation) -> ltrb -> xywh * stride; sigmoid cls
  Segment     Detect + cv4(x) mask coefficients per anchor + Proto(x[0]): cv3(cv2(ConvTranspose2d(2,2)(cv1(x)))), SiLU 1x1

Pre / post-processing (LetterBox, non_max_suppression, scale_boxes) are the functions of oracle/yolo_oracle.py."""
from __future__ import annotations

from typing import List

import numpy as np
import torch
import torch.nn.functional as F

import yolo_oracle


_CALIBRATING = False


def conv(node, x):
    y = F.conv2d(x, node["w"], node["b"], stride=node["s"], padding=node["p"], groups=node["g"])
    if _CALIBRATING:
        # what training's BatchNorm does for a real checkpoint: per-channel zero mean / unit variance pre-activations
        mean, std = y.mean((0, 2, 3)), y.std((0, 2, 3)).clamp_min(1e-6)
        node["w"] = (node["w"] / std.view(-1, 1, 1, 1)).contiguous()
        node["b"] = ((node["b"] - mean) / std).contiguous()
        y = (y - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1)
    return F.silu(y) if node["act"] else y


@torch.no_grad()
def calibrate(tree: dict, img_bgr: np.ndarray, imgsz: int = 640, cls_mean: float = -5.0, cls_std: float = 1.5,
              box_std: float = 2.0) -> dict:
    """Condition a seeded synthetic tree IN PLACE like a trained network: one forward over a probe image during which
    every convolution is rescaled to zero-mean / unit-variance outputs (the job BatchNorm does in a real checkpoint), then
    the class logits are set to N(cls_mean, cls_std^2) so that a few dozen of the anchors clear the callers' thresholds
    by margins far above numerical noise.  Test infrastructure: both the CUDA path and this oracle then run the SAME
    conditioned tree."""
    global _CALIBRATING
    _CALIBRATING = True
    try:
        forward(tree, yolo_oracle.preprocess(img_bgr, imgsz))
    finally:
        _CALIBRATING = False
    det = tree["layers"][-1]
    if det["t"] == "Segment":       # prototypes / coefficients of O(1): mask logits of a few units, like a trained model
        pr = det["proto"]
        pr["upsample"]["w"], pr["upsample"]["b"] = pr["upsample"]["w"] * 0.5, pr["upsample"]["b"] * 0.5
    for br in det["cv2"]:
        br[-1]["w"], br[-1]["b"] = br[-1]["w"] * box_std, br[-1]["b"] * box_std
    for br in det["cv3"]:
        br[-1]["w"], br[-1]["b"] = br[-1]["w"] * cls_std, br[-1]["b"] * cls_std + cls_mean
    return tree


def seq(nodes, x):
    for n in nodes:
        x = conv(n, x)
    return x


def bottleneck(node, x):
    y = conv(node["cv2"], conv(node["cv1"], x))
    return x + y if node["add"] else y


def c3(node, x):
    a = conv(node["cv1"], x)
    for b in node["m"]:
        a = block(b, a)
    return conv(node["cv3"], torch.cat((a, conv(node["cv2"], x)), 1))


def c2f(node, x):
    y = list(conv(node["cv1"], x).chunk(2, 1))
    for b in node["m"]:
        y.append(block(b, y[-1]))
    return conv(node["cv2"], torch.cat(y, 1))


def sppf(node, x):
    y = [conv(node["cv1"], x)]
    k = int(node["k"])
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], k, 1, k // 2))
    return conv(node["cv2"], torch.cat(y, 1))


def attention(a, x):
    B, C, H, W = x.shape
    N = H * W
    nh, kd, hd = a["num_heads"], a["key_dim"], a["head_dim"]
    qkv = conv(a["qkv"], x)
    q, k, v = qkv.view(B, nh, kd * 2 + hd, N).split([kd, kd, hd], dim=2)
    attn = (q.transpose(-2, -1) @ k) * a["scale"]
    attn = attn.softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + conv(a["pe"], v.reshape(B, C, H, W))
    return conv(a["proj"], y)


def psablock(node, x):
    y = attention(node["attn"], x)
    x = x + y if node["add"] else y
    y = seq(node["ffn"], x)
    return x + y if node["add"] else y


def c2psa(node, x):
    c = int(node["c"])
    a, b = conv(node["cv1"], x).split((c, c), dim=1)
    for m in node["m"]:
        b = psablock(m, b)
    return conv(node["cv2"], torch.cat((a, b), 1))


def aattn(a, x):
    B, C, H, W = x.shape
    N = H * W
    nh, hd, area = a["num_heads"], a["head_dim"], int(a["area"])
    qkv = conv(a["qkv"], x).flatten(2).transpose(1, 2)
    if area > 1:
        qkv = qkv.reshape(B * area, N // area, C * 3)
        B, N, _ = qkv.shape
    q, k, v = qkv.view(B, N, nh, hd * 3).permute(0, 2, 3, 1).split([hd, hd, hd], dim=2)
    attn = (q.transpose(-2, -1) @ k) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).permute(0, 3, 1, 2)
    v = v.permute(0, 3, 1, 2)
    if area > 1:
        y = y.reshape(B // area, N * area, C)
        v = v.reshape(B // area, N * area, C)
        B, N, _ = y.shape
    y = y.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    v = v.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    return conv(a["proj"], y + conv(a["pe"], v))


def ablock(node, x):
    x = x + aattn(node["attn"], x)
    return x + seq(node["mlp"], x)


def a2c2f(node, x):
    y = [conv(node["cv1"], x)]
    for m in node["m"]:
        t = y[-1]
        if isinstance(m, list):
            for ab in m:
                t = ablock(ab, t)
        else:
            t = block(m, t)
        y.append(t)
    out = conv(node["cv2"], torch.cat(y, 1))
    if node.get("gamma") is not None:
        return x + node["gamma"].view(1, -1, 1, 1) * out
    return out


def block(node, x):
    return {"Conv": conv, "Bottleneck": bottleneck, "C3": c3, "C2f": c2f, "SPPF": sppf, "C2PSA": c2psa, "A2C2f": a2c2f,
            "PSABlock": psablock, "ABlock": ablock}[node["t"]](node, x)


@torch.no_grad()
def forward(tree: dict, x: torch.Tensor):
    """x [1][3][H][W] in [0,1] -> (pred [1][4+nc][A] (xywh letterbox px, class probabilities), heads [(box, cls)] raw)"""
    outs: List[torch.Tensor] = []
    cur = x
    seg = None
    for node in tree["layers"]:
        f = node.get("f", -1)
        srcs = [cur if j == -1 else outs[j] for j in (f if isinstance(f, (list, tuple)) else [f])]
        t = node["t"]
        if t == "Concat":
            cur = torch.cat(srcs, 1)
        elif t == "Upsample":
            cur = F.interpolate(srcs[0], scale_factor=2, mode="nearest")
        elif t in ("Detect", "Segment"):
            heads = [(seq(node["cv2"][i], s), seq(node["cv3"][i], s)) for i, s in enumerate(srcs)]
            if t == "Segment":
                mcs = [seq(node["cv4"][i], s) for i, s in enumerate(srcs)]
                pr = node["proto"]
                p = F.conv_transpose2d(conv(pr["cv1"], srcs[0]), pr["upsample"]["w"], pr["upsample"]["b"], stride=2)
                seg = (mcs, conv(pr["cv3"], conv(pr["cv2"], p)))
            cur = heads
        else:
            cur = block(node, srcs[0])
        outs.append(cur)
    heads = outs[-1]
    det = tree["layers"][-1]
    nc = int(det["nc"])
    box, cls, anchors, strides = [], [], [], []
    for i, (b, c) in enumerate(heads):
        bs, _, fh, fw = b.shape
        box.append(b.view(bs, 64, -1))
        cls.append(c.view(bs, nc, -1))
        sx = torch.arange(fw, dtype=torch.float32) + 0.5
        sy = torch.arange(fh, dtype=torch.float32) + 0.5
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        anchors.append(torch.stack((xx, yy), -1).view(-1, 2))
        strides.append(torch.full((fh * fw, 1), float(det["stride"][i])))
    box, cls = torch.cat(box, 2), torch.cat(cls, 2)
    anchors, strides = torch.cat(anchors).t().unsqueeze(0), torch.cat(strides).t()
    bs, _, a = box.shape
    dist = box.view(bs, 4, 16, a).transpose(2, 1).softmax(1)
    dist = (dist * torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)).sum(1)
    lt, rb = dist.chunk(2, 1)
    x1y1, x2y2 = anchors - lt, anchors + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
    pred = torch.cat((dbox, cls.sigmoid()), 1)
    if seg is not None:          # Segment: mask coefficients ride behind the class scores, the prototypes come along
        pred = torch.cat((pred, torch.cat([m.view(bs, m.shape[1], -1) for m in seg[0]], 2)), 1)
        return pred, heads, seg
    return pred, heads


@torch.no_grad()
def predict(tree: dict, img_bgr: np.ndarray, conf: float, imgsz: int = 640):
    """-> dict(xyxy [n,4] original px, conf [n], cls [n], anchors [n], heads)"""
    h0, w0 = img_bgr.shape[:2]
    x = yolo_oracle.preprocess(img_bgr, imgsz)
    out = forward(tree, x)
    pred, heads, seg = out if len(out) == 3 else (out[0], out[1], None)
    nc = int(tree["layers"][-1]["nc"])
    det, anchors = yolo_oracle.non_max_suppression(pred[0], nc, conf)
    if det.shape[0] == 0:
        return dict(xyxy=torch.zeros((0, 4)), conf=torch.zeros(0), cls=torch.zeros(0), anchors=anchors, heads=heads, raw=pred,
                    masks=torch.zeros((0, h0, w0), dtype=torch.bool), seg=seg)
    boxes = yolo_oracle.scale_boxes(x.shape[2:], det[:, :4], (h0, w0))
    masks = None
    if seg is not None:
        masks = yolo_oracle.process_mask_native(seg[1][0], det[:, 6:], boxes, x.shape[2:], (h0, w0))
    return dict(xyxy=boxes, conf=det[:, 4], cls=det[:, 5], anchors=anchors, heads=heads, raw=pred, masks=masks, seg=seg)
