"""B200 YOLOv8-seg speech-bubble detector — the object `ModelManager.load_yolo_speech_bubble()` returns.

Duck-types what the reference touches on an ultralytics model (core/image/detection.py:1338-1350, 525-556):
`m(image_bgr, conf=..., device=..., verbose=False, imgsz=..., retina_masks=True) -> [Results]` with
`Results.boxes.{xyxy,conf,cls}`, `Results.masks`, `Results.orig_shape`, and `m.names`.

The whole graph runs on tcgen05 conv plans (bf16x3 = fp32-grade by default).  Concats never copy: every producer
writes straight into its channel slice of the consumer's concat buffer (the conv ABI takes channel offsets), C2f splits
are channel-offset reads, SPPF pools and FPN upsamples write slices too.  Head decode (DFL/sigmoid/confidence), NMS,
scale_boxes and the reference's own IoU-dedup + containment filter run in two small kernels (detect_kernels.cu).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import planes as P
from ._lib import check, lib, ptr, stream_ptr
from .ops import ConvPlan


class YoloLevel(C.Structure):
    _fields_ = [("box", C.c_void_p), ("cls", C.c_void_p), ("H", C.c_int), ("W", C.c_int), ("stride", C.c_int),
                ("reserved", C.c_int)]


class NmsParams(C.Structure):
    _fields_ = [("N", C.c_int), ("max_cand", C.c_int), ("max_det", C.c_int), ("iou_thr", C.c_float),
                ("max_wh", C.c_float), ("gain", C.c_float), ("pad_x", C.c_int), ("pad_y", C.c_int),
                ("img_w", C.c_int), ("img_h", C.c_int), ("dedup_iou", C.c_double), ("contain_ioa", C.c_double),
                ("apply_dedup", C.c_int)]


def _declare(l) -> None:
    if getattr(l, "_det_declared", False):
        return
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    l.mtb_maxpool.argtypes = [vp, vp] + [i32] * 10 + [vp]
    l.mtb_upsample2x.argtypes = [vp, vp] + [i32] * 9 + [vp]
    l.mtb_yolo_decode.argtypes = [C.POINTER(YoloLevel), i32, i32, i32, i32, f32, i32, vp, vp, vp, vp]
    l.mtb_nms.argtypes = [C.POINTER(NmsParams), vp, vp, vp, vp, vp, vp, vp, vp, vp]
    l.mtb_image_to_planes.argtypes = [vp, i32, i32, i32, i32, f32, C.POINTER(f32), vp, i32, i32, vp]
    l.mtb_yolo_masks.argtypes = [vp, i32, i32, i32, C.POINTER(vp), C.POINTER(i32), vp, vp] + [i32] * 7 + [vp, vp]
    l.mtb_yolo_masks.restype = i32
    for n in ("mtb_maxpool", "mtb_upsample2x", "mtb_yolo_decode", "mtb_nms", "mtb_image_to_planes"):
        getattr(l, n).restype = i32
    l._det_declared = True


def letterbox_params(h0: int, w0: int, imgsz: int, stride: int = 32):
    """ultralytics LetterBox(auto=True, scaleup=True) geometry: (new w,h), (top,bottom,left,right), out (h,w), gain."""
    r = min(imgsz / h0, imgsz / w0)
    nw, nh = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = ((imgsz - nw) % stride) / 2, ((imgsz - nh) % stride) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return (nw, nh), (top, bottom, left, right), (nh + top + bottom, nw + left + right), r


class _Slice:
    """A channel slice [off, off+c) of an NHWC plane buffer."""
    __slots__ = ("buf", "off", "c")

    def __init__(self, buf, off, c):
        self.buf, self.off, self.c = buf, off, c


class Boxes:
    def __init__(self, xyxy, conf, cls):
        self.xyxy, self.conf, self.cls = xyxy, conf, cls

    def __len__(self):
        return int(self.xyxy.shape[0])


class Masks:
    """`Results.masks` as far as the reference touches it: `.data` (n,H,W float {0,1}) and `len()`."""

    def __init__(self, data):
        self.data = data

    def __len__(self):
        return int(self.data.shape[0])


class Results:
    def __init__(self, boxes: Optional[Boxes], masks, orig_shape, names):
        self.boxes, self.masks, self.orig_shape, self.names = boxes, masks, orig_shape, names


class YoloB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: dict, device: torch.device, *,
                 precision: str = "bf16x3", names: Optional[dict] = None):
        self.l = lib()
        _declare(self.l)
        self.device = device
        self.planes = 2 if precision == "bf16x3" else 1
        self.cfg = dict(cfg)
        self.nc = cfg.get("nc", 1)
        self.names = names or {i: ("speech_bubble" if i == 0 else f"class{i}") for i in range(self.nc)}
        self.sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self._w: Dict[str, tuple] = {}
        self._plans: Dict[tuple, dict] = {}
        depth, width, max_ch = cfg.get("depth", 0.67), cfg.get("width", 0.75), cfg.get("max_ch", 768)
        self._d = lambda n: max(round(n * depth), 1)
        self._c = lambda x: int(math.ceil(min(x, max_ch) * width / 8) * 8)

    # ---- weights ---------------------------------------------------------------------------------------------
    def _conv_w(self, name: str):
        if name not in self._w:
            w = self.sd[name + ".weight"]
            self._w[name] = (P.conv_weight_to_planes(w, self.planes), P.pad_bias(self.sd.get(name + ".bias"), w.shape[0]))
        return self._w[name]

    def _deconv_w(self, name: str):
        """ConvTranspose2d(k=2,s=2) as a 1x1 conv to 4*Cout channels + pixel-shuffle store."""
        if name not in self._w:
            w = self.sd[name + ".weight"]          # [Cin][Cout][2][2]
            cin, cout = w.shape[0], w.shape[1]
            w1 = w.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1).contiguous()   # [(dy*2+dx)*Cout + co][ci]
            b = self.sd[name + ".bias"].repeat(4)
            self._w[name] = (P.conv_weight_to_planes(w1, self.planes), P.pad_bias(b, 4 * cout))
        return self._w[name]

    # ---- graph -----------------------------------------------------------------------------------------------
    def _build(self, n: int, h: int, w: int) -> dict:
        dev, pl = self.device, self.planes
        c, d = self._c, self._d
        steps: List[tuple] = []
        keep: List[torch.Tensor] = []

        def buf(hh, ww, ch):
            t = torch.zeros((pl, n, hh, ww, ch), dtype=torch.bfloat16, device=dev)
            keep.append(t)
            return t

        def conv(src: _Slice, name: str, dst: _Slice, k=1, s=1, act="silu", residual: Optional[_Slice] = None):
            wgt = self._conv_w(name)
            steps.append(("conv", ConvPlan(src.buf, wgt[0], wgt[1], dst.buf, k=k, stride=s, pad=k // 2, act=act,
                                           x_coff=src.off, out_coff=dst.off,
                                           residual=None if residual is None else residual.buf,
                                           res_coff=0 if residual is None else residual.off)))

        def c2f(src: _Slice, name: str, dst: _Slice, nb: int, shortcut: bool, hh: int, ww: int):
            c2 = dst.c
            hc = c2 // 2
            cat = buf(hh, ww, (2 + nb) * hc)
            tmp = buf(hh, ww, hc)
            conv(src, f"{name}.cv1.conv", _Slice(cat, 0, 2 * hc), 1)
            for i in range(nb):
                sin = _Slice(cat, (1 + i) * hc, hc)
                conv(sin, f"{name}.m.{i}.cv1.conv", _Slice(tmp, 0, hc), 3)
                conv(_Slice(tmp, 0, hc), f"{name}.m.{i}.cv2.conv", _Slice(cat, (2 + i) * hc, hc), 3,
                     residual=sin if shortcut else None)
            conv(_Slice(cat, 0, (2 + nb) * hc), f"{name}.cv2.conv", dst, 1)

        c64, c128, c256, c512, c1024 = c(64), c(128), c(256), c(512), c(1024)
        h2, w2, h4, w4, h8, w8, h16, w16, h32, w32 = h // 2, w // 2, h // 4, w // 4, h // 8, w // 8, h // 16, w // 16, h // 32, w // 32
        x_in = buf(h, w, 8)
        t0, t1, t2 = buf(h2, w2, c64), buf(h4, w4, c128), buf(h4, w4, c128)
        cat14 = buf(h8, w8, c512 + c256)        # [up(n4) | p3]
        cat11 = buf(h16, w16, c1024 + c512)     # [up(p5) | p4]
        cat17 = buf(h16, w16, c256 + c512)      # [l16 | n4]
        cat20 = buf(h32, w32, c512 + c1024)     # [l19 | p5]
        p3 = _Slice(cat14, c512, c256)
        p4 = _Slice(cat11, c1024, c512)
        p5 = _Slice(cat20, c512, c1024)
        n4 = _Slice(cat17, c256, c512)
        t3, t5, t7, t8 = buf(h8, w8, c256), buf(h16, w16, c512), buf(h32, w32, c1024), buf(h32, w32, c1024)
        n3, m4, m5 = buf(h8, w8, c256), buf(h16, w16, c512), buf(h32, w32, c1024)

        conv(_Slice(x_in, 0, 8), "l0.conv", _Slice(t0, 0, c64), 3, 2)
        conv(_Slice(t0, 0, c64), "l1.conv", _Slice(t1, 0, c128), 3, 2)
        c2f(_Slice(t1, 0, c128), "l2", _Slice(t2, 0, c128), d(3), True, h4, w4)
        conv(_Slice(t2, 0, c128), "l3.conv", _Slice(t3, 0, c256), 3, 2)
        c2f(_Slice(t3, 0, c256), "l4", p3, d(6), True, h8, w8)
        conv(p3, "l5.conv", _Slice(t5, 0, c512), 3, 2)
        c2f(_Slice(t5, 0, c512), "l6", p4, d(6), True, h16, w16)
        conv(p4, "l7.conv", _Slice(t7, 0, c1024), 3, 2)
        c2f(_Slice(t7, 0, c1024), "l8", _Slice(t8, 0, c1024), d(3), True, h32, w32)
        # SPPF
        hc = c1024 // 2
        sp = buf(h32, w32, 4 * hc)
        conv(_Slice(t8, 0, c1024), "l9.cv1.conv", _Slice(sp, 0, hc), 1)
        for i in range(3):
            steps.append(("maxpool", (sp, n, h32, w32, 4 * hc, i * hc, 4 * hc, (i + 1) * hc, hc, 5)))
        conv(_Slice(sp, 0, 4 * hc), "l9.cv2.conv", p5, 1)
        # FPN / PAN
        steps.append(("up", (p5, _Slice(cat11, 0, c1024), n, h32, w32)))
        c2f(_Slice(cat11, 0, c1024 + c512), "l12", n4, d(3), False, h16, w16)
        steps.append(("up", (n4, _Slice(cat14, 0, c512), n, h16, w16)))
        c2f(_Slice(cat14, 0, c512 + c256), "l15", _Slice(n3, 0, c256), d(3), False, h8, w8)
        conv(_Slice(n3, 0, c256), "l16.conv", _Slice(cat17, 0, c256), 3, 2)
        c2f(_Slice(cat17, 0, c256 + c512), "l18", _Slice(m4, 0, c512), d(3), False, h16, w16)
        conv(_Slice(m4, 0, c512), "l19.conv", _Slice(cat20, 0, c512), 3, 2)
        c2f(_Slice(cat20, 0, c512 + c1024), "l21", _Slice(m5, 0, c1024), d(3), False, h32, w32)
        # Segment head
        feats = [(n3, c256, h8, w8, 8), (m4, c512, h16, w16, 16), (m5, c1024, h32, w32, 32)]
        nc, nm = self.nc, self.cfg.get("nm", 32)
        ncp = P.pad_to(nc, 16)
        c2h = max(16, c256 // 4, 64)
        c3h = max(c256, min(nc, 100))
        c4h = max(c256 // 4, nm)
        levels = []
        for i, (fb, fc, fh, fw, st) in enumerate(feats):
            f = _Slice(fb, 0, fc)
            outs = {}
            for br, hcx, oc, ocp in (("cv2", c2h, 64, 64), ("cv3", c3h, nc, ncp), ("cv4", c4h, nm, nm)):
                a, b2 = buf(fh, fw, hcx), buf(fh, fw, hcx)
                conv(f, f"head.{br}.{i}.0.conv", _Slice(a, 0, hcx), 3)
                conv(_Slice(a, 0, hcx), f"head.{br}.{i}.1.conv", _Slice(b2, 0, hcx), 3)
                o = torch.zeros((n, fh, fw, ocp), dtype=torch.float32, device=dev)
                keep.append(o)
                wgt = self._conv_w(f"head.{br}.{i}.2")
                steps.append(("conv", ConvPlan(b2, wgt[0], wgt[1], o, k=1, act=None)))
                outs[br] = o
            levels.append((outs["cv2"], outs["cv3"], outs["cv4"], fh, fw, st))
        # Proto
        npr = c(self.cfg.get("npr", 256))
        pa, pb = buf(h8, w8, npr), buf(h4, w4, npr)
        pc = buf(h4, w4, npr)
        conv(_Slice(n3, 0, c256), "head.proto.cv1.conv", _Slice(pa, 0, npr), 3)
        wgt = self._deconv_w("head.proto.upsample")
        steps.append(("conv", ConvPlan(pa, wgt[0], wgt[1], pb, k=1, act=None, pixel_shuffle=True)))
        conv(_Slice(pb, 0, npr), "head.proto.cv2.conv", _Slice(pc, 0, npr), 3)
        proto = torch.zeros((n, h4, w4, nm), dtype=torch.float32, device=dev)
        wgt = self._conv_w("head.proto.cv3.conv")
        steps.append(("conv", ConvPlan(pc, wgt[0], wgt[1], proto, k=1, act="silu")))
        total_anchors = sum(fh * fw for (_, _, _, fh, fw, _) in levels)
        max_cand = min(total_anchors, 30000)
        return dict(steps=steps, keep=keep, x_in=x_in, levels=levels, proto=proto, ncp=ncp, max_cand=max_cand,
                    total_anchors=total_anchors,
                    cand=torch.zeros((n, max_cand, 6), dtype=torch.float32, device=dev),
                    cand_anchor=torch.zeros((n, max_cand), dtype=torch.int32, device=dev),
                    count=torch.zeros((n,), dtype=torch.int32, device=dev),
                    order=torch.zeros((n, max_cand), dtype=torch.int32, device=dev),
                    dead=torch.zeros((n, max_cand), dtype=torch.uint8, device=dev),
                    det=torch.zeros((n, 300, 8), dtype=torch.float32, device=dev),
                    det_count=torch.zeros((n, 2), dtype=torch.int32, device=dev),
                    final_idx=torch.zeros((n, 300), dtype=torch.int32, device=dev))

    def _get(self, n, h, w):
        key = (n, h, w)
        if key not in self._plans:
            self._plans[key] = self._build(n, h, w)
        return self._plans[key]

    def _run_graph(self, g: dict) -> None:
        l, st, pl = self.l, stream_ptr(), self.planes
        for kind, arg in g["steps"]:
            if kind == "conv":
                arg.run()
            elif kind == "maxpool":
                t, n, hh, ww, ct, ci, ct2, co, c, k = arg
                check(l.mtb_maxpool(ptr(t), ptr(t), n, hh, ww, ct, ci, ct2, co, c, k, pl, st), "mtb_maxpool")
            else:
                src, dst, n, hh, ww = arg
                check(l.mtb_upsample2x(ptr(src.buf), ptr(dst.buf), n, hh, ww, src.buf.shape[-1], src.off,
                                       dst.buf.shape[-1], dst.off, src.c, pl, st), "mtb_upsample2x")

    # ---- public ----------------------------------------------------------------------------------------------
    def forward_letterboxed(self, lb_rgb_u8: torch.Tensor):
        """lb_rgb_u8: device uint8 [H][W][3] RGB letterboxed input.  Runs the graph; returns the graph dict."""
        from . import graphs
        h, w, c = lb_rgb_u8.shape
        g = self._get(1, h, w)
        if "lb_in" not in g:
            g["lb_in"] = torch.empty((h, w, 3), dtype=torch.uint8, device=self.device)

        def body():
            zero = (C.c_float * 3)(0.0, 0.0, 0.0)
            check(self.l.mtb_image_to_planes(ptr(g["lb_in"]), h, w, 3, 0, 1.0 / 255.0, zero, ptr(g["x_in"]), 8, self.planes,
                                             stream_ptr()), "mtb_image_to_planes")
            self._run_graph(g)

        g["lb_in"].copy_(lb_rgb_u8[:, :, :3])
        if graphs.ENABLED:
            if "cuda_graph" not in g:
                g["cuda_graph"] = graphs.CapturedGraph(body)
            g["cuda_graph"].replay()
        else:
            body()
        return g

    def forward_letterboxed_batch(self, lbs: List[torch.Tensor]):
        """Several letterboxed inputs of ONE size through one plan with a batch dimension: the deep layers of a single
        1600-pixel page fill only a fraction of the 148 SMs (stride 32: 50 x 34 pixels = 14 tiles of 128), a group of 8
        pages fills them.  Same kernels, same per-output arithmetic: results are bit-identical to page-by-page runs."""
        from . import graphs
        n = len(lbs)
        h, w, _ = lbs[0].shape
        assert all(tuple(lb.shape) == (h, w, 3) for lb in lbs)
        g = self._get(n, h, w)
        if "lb_in" not in g:
            g["lb_in"] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)

        def body():
            zero = (C.c_float * 3)(0.0, 0.0, 0.0)
            # [n][h][w][3] is one image of n*h rows: the planes come out as [planes][n][h][w][8]
            check(self.l.mtb_image_to_planes(ptr(g["lb_in"]), n * h, w, 3, 0, 1.0 / 255.0, zero, ptr(g["x_in"]), 8, self.planes,
                                             stream_ptr()), "mtb_image_to_planes")
            self._run_graph(g)

        for i, lb in enumerate(lbs):
            g["lb_in"][i].copy_(lb[:, :, :3])
        if graphs.ENABLED:
            if "cuda_graph" not in g:
                g["cuda_graph"] = graphs.CapturedGraph(body)
            g["cuda_graph"].replay()
        else:
            body()
        return g

    def detect_batch(self, g: dict, conf: float, orig_hw: Tuple[int, int], lb_hw: Tuple[int, int], *, iou: float = 0.7,
                     apply_reference_dedup: bool = True):
        """`detect` for a batched plan (all pages of one original size).  Returns (det [n][300][8], counts [n][2],
        final_idx [n][300])."""
        l, st = self.l, stream_ptr()
        n = int(g["det"].shape[0])
        lv = (YoloLevel * 3)()
        for i, (box, cls, _, fh, fw, s) in enumerate(g["levels"]):
            lv[i].box, lv[i].cls, lv[i].H, lv[i].W, lv[i].stride = box.data_ptr(), cls.data_ptr(), fh, fw, s
        check(l.mtb_yolo_decode(lv, 3, n, self.nc, g["ncp"], float(conf), g["max_cand"], ptr(g["cand"]),
                                ptr(g["cand_anchor"]), ptr(g["count"]), st), "mtb_yolo_decode")
        h0, w0 = orig_hw
        gain = min(lb_hw[0] / h0, lb_hw[1] / w0)
        p = NmsParams()
        p.N, p.max_cand, p.max_det = n, g["max_cand"], 300
        p.iou_thr, p.max_wh, p.gain = float(iou), 7680.0, float(gain)
        p.pad_x = int(round((lb_hw[1] - w0 * gain) / 2 - 0.1))
        p.pad_y = int(round((lb_hw[0] - h0 * gain) / 2 - 0.1))
        p.img_w, p.img_h = w0, h0
        p.dedup_iou, p.contain_ioa, p.apply_dedup = 0.7, 0.9, int(apply_reference_dedup)
        check(l.mtb_nms(C.byref(p), ptr(g["cand"]), ptr(g["cand_anchor"]), ptr(g["count"]), ptr(g["order"]),
                        ptr(g["dead"]), ptr(g["det"]), ptr(g["det_count"]), ptr(g["final_idx"]), st), "mtb_nms")
        return g["det"], g["det_count"], g["final_idx"]

    def detect(self, g: dict, conf: float, orig_hw: Tuple[int, int], lb_hw: Tuple[int, int], *, iou: float = 0.7,
               apply_reference_dedup: bool = True):
        """Decode + NMS + scale_boxes (+ the reference's dedup/containment).  Returns device tensors
        (det [300][8], counts [2], final_idx [300])."""
        l, st = self.l, stream_ptr()
        lv = (YoloLevel * 3)()
        for i, (box, cls, _, fh, fw, s) in enumerate(g["levels"]):
            lv[i].box, lv[i].cls, lv[i].H, lv[i].W, lv[i].stride = box.data_ptr(), cls.data_ptr(), fh, fw, s
        check(l.mtb_yolo_decode(lv, 3, 1, self.nc, g["ncp"], float(conf), g["max_cand"], ptr(g["cand"]),
                                ptr(g["cand_anchor"]), ptr(g["count"]), st), "mtb_yolo_decode")
        h0, w0 = orig_hw
        gain = min(lb_hw[0] / h0, lb_hw[1] / w0)
        p = NmsParams()
        p.N, p.max_cand, p.max_det = 1, g["max_cand"], 300
        p.iou_thr, p.max_wh, p.gain = float(iou), 7680.0, float(gain)
        p.pad_x = int(round((lb_hw[1] - w0 * gain) / 2 - 0.1))
        p.pad_y = int(round((lb_hw[0] - h0 * gain) / 2 - 0.1))
        p.img_w, p.img_h = w0, h0
        p.dedup_iou, p.contain_ioa, p.apply_dedup = 0.7, 0.9, int(apply_reference_dedup)
        check(l.mtb_nms(C.byref(p), ptr(g["cand"]), ptr(g["cand_anchor"]), ptr(g["count"]), ptr(g["order"]),
                        ptr(g["dead"]), ptr(g["det"]), ptr(g["det_count"]), ptr(g["final_idx"]), st), "mtb_nms")
        return g["det"][0], g["det_count"][0], g["final_idx"][0]

    def __call__(self, image_bgr: np.ndarray, conf: float = 0.25, device=None, verbose: bool = False, imgsz: int = 640,
                 retina_masks: bool = True, **_):
        """Reference call shape (core/image/detection.py:1338-1345).  The letterbox resize itself is done by
        `letterbox_device` (bilinear, cv2.INTER_LINEAR semantics) — see mangatranslator_b200/preproc.py."""
        from .preproc import letterbox_device
        h0, w0 = image_bgr.shape[:2]
        img = torch.from_numpy(np.ascontiguousarray(image_bgr)).to(self.device)
        lb = letterbox_device(img, imgsz, swap_rb=True)
        g = self.forward_letterboxed(lb)
        det, cnt, _ = self.detect(g, conf, (h0, w0), tuple(lb.shape[:2]), apply_reference_dedup=False)
        n = int(cnt[0].item())
        d = det[:n]
        boxes = Boxes(d[:, :4].contiguous(), d[:, 4].contiguous(), d[:, 5].contiguous()) if n else None
        masks = None
        if n and retina_masks:
            masks = Masks(self.retina_masks(g, det, None, n, (h0, w0), tuple(lb.shape[:2])).float())
        return [Results(boxes, masks, (h0, w0), self.names)]

    def retina_masks(self, g: dict, det: torch.Tensor, rows: Optional[torch.Tensor], n: int, orig_hw, lb_hw) -> torch.Tensor:
        """uint8 {0,1} masks [n][H][W] for rows of the detection table (process_mask_native)."""
        h0, w0 = orig_hw
        proto = g["proto"]
        mh, mw, nm = int(proto.shape[1]), int(proto.shape[2]), int(proto.shape[3])
        gain = min(mh / h0, mw / w0)
        pad_w, pad_h = (mw - w0 * gain) / 2, (mh - h0 * gain) / 2
        top, left = int(round(pad_h - 0.1)), int(round(pad_w - 0.1))
        bottom, right = mh - int(round(pad_h + 0.1)), mw - int(round(pad_w + 0.1))
        out = torch.empty((n, h0, w0), dtype=torch.uint8, device=self.device)
        mcs = (C.c_void_p * 3)(*[lv[2].data_ptr() for lv in g["levels"]])
        hw = (C.c_int * 6)(*[v for lv in g["levels"] for v in (lv[3], lv[4])])
        check(self.l.mtb_yolo_masks(ptr(proto), mh, mw, nm, mcs, hw, ptr(det), ptr(rows), n, top, left, bottom - top,
                                    right - left, h0, w0, ptr(out), stream_ptr()), "mtb_yolo_masks")
        return out
