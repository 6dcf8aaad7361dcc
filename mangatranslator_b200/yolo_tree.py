"""Ultralytics detection models executed from their own MODULE TREE: the panel detector (YOLO11-L,
`/root/reference/core/image/detection.py:1817-1921` `detect_panels`, loader `core/ml/model_manager.py:780-807`) and the
outside-speech-bubble text detector (YOLO12x, `detection.py:120-201` `_expand_boxes_with_osb_text`, loader
`model_manager.py:809-833`).  Both are called at imgsz 640 with boxes only (`Detect` head, no masks).

The speech-bubble detector (mangatranslator_b200/yolo.py) hard-wires the YOLOv8-seg graph because it is the hot path.
These two are secondary models of unknown exact size (the checkpoints are not available offline), so nothing about the
architecture is assumed: the `.pt` file's pickled model object is walked (weights.load_ultralytics_tree, inert classes,
nothing from the file is executed) into a plain tree of nodes

    Conv{w,b,k,s,p,g,act}  Bottleneck{cv1,cv2,add}  C2f{cv1,cv2,m[]}  C3{cv1,cv2,cv3,m[]}  SPPF{cv1,cv2,k}
    C2PSA{c,cv1,cv2,m[PSABlock{attn,ffn[],add}]}  Attention{num_heads,key_dim,head_dim,scale,qkv,proj,pe}
    A2C2f{cv1,cv2,gamma,m[ [ABlock{attn,mlp[]},..] | C3 ]}  AAttn{area,num_heads,head_dim,qkv,proj,pe}
    Concat{d}  Upsample{scale}  Detect{nc,reg_max,stride[],cv2[][],cv3[][]}            (+ "f": where a layer reads from)
    Segment = Detect + {nm, cv4[][], proto{cv1, upsample{w,b}, cv2, cv3}}

with BatchNorm folded (eps from the file), and `YoloTreeB200` turns that tree into a static plan of tcgen05 conv plans
(mtb_conv_plan_*), depthwise convs (mtb_dwconv), attention (mtb_attention, exact fp32 softmax), max-pool / upsample, and
the same decode + NMS kernels as the primary detector.  What is restated FROM MEMORY of ultralytics (absent here, so
"parity unpinned" like the YOLOv8 oracle) is only each block's forward rule; `oracle/yolo_tree_oracle.py` is the CPU
statement of the same rules on the same tree.  `synthetic_tree` builds seeded YOLO11 / YOLO12 trees of a given scale from
the published yaml layouts for tests and the opt-in synthetic mode."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import planes as P
from ._lib import check, lib, ptr, stream_ptr
from .ops import ConvPlan
from .weights import UnsupportedCheckpoint
from .yolo import Boxes, Masks, Results, YoloB200, _Slice, _declare as _declare_yolo


# ---- node helpers -----------------------------------------------------------------------------------------------------
def conv_node(w: torch.Tensor, b: torch.Tensor, k: int, s: int = 1, p: Optional[int] = None, g: int = 1, act: bool = True):
    return {"t": "Conv", "w": w.float().contiguous(), "b": b.float().contiguous(), "k": int(k), "s": int(s),
            "p": int(k // 2 if p is None else p), "g": int(g), "act": bool(act)}


def out_channels(node: dict, cin: Optional[int] = None) -> int:
    t = node["t"]
    if t == "Conv":
        return int(node["w"].shape[0])
    if t in ("C2f", "SPPF", "C2PSA", "A2C2f"):
        return int(node["cv2"]["w"].shape[0])
    if t == "C3":
        return int(node["cv3"]["w"].shape[0])
    if t == "Bottleneck":
        return int(node["cv2"]["w"].shape[0])
    raise KeyError(t)


# ---- seeded synthetic trees (tests, smoke, bench opt-in): the published yaml layouts, restated from memory ------------
SCALES = {"n": (0.50, 0.25, 1024), "s": (0.50, 0.50, 1024), "m": (0.50, 1.00, 512), "l": (1.00, 1.00, 512),
          "x": (1.00, 1.50, 512)}


class _Synth:
    def __init__(self, seed: int, scale: str, nc: int):
        self.g = torch.Generator().manual_seed(seed)
        self.scale = scale
        self.depth, self.width, self.max_ch = SCALES[scale]
        self.nc = nc

    def ch(self, c: int) -> int:                      # make_divisible(min(c, max_channels) * width, 8)
        return int(math.ceil(min(c, self.max_ch) * self.width / 8) * 8)

    def rep(self, n: int) -> int:
        return max(round(n * self.depth), 1) if n > 1 else n

    def conv(self, c1, c2, k=1, s=1, g=1, act=True, gain=1.0, bias_std=0.05):
        fan = (c1 // g) * k * k
        std = gain * math.sqrt((2.6 if act else 1.0) / fan)
        w = torch.randn((c2, c1 // g, k, k), generator=self.g) * std
        b = torch.randn((c2,), generator=self.g) * bias_std
        return conv_node(w, b, k, s, None, g, act)

    def bottleneck(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        c_ = int(c2 * e)
        return {"t": "Bottleneck", "cv1": self.conv(c1, c_, k[0]), "cv2": self.conv(c_, c2, k[1], gain=0.8),
                "add": bool(shortcut and c1 == c2)}

    def c3k(self, c1, c2, n=2, shortcut=True, e=0.5, k=3):
        c_ = int(c2 * e)
        return {"t": "C3", "cv1": self.conv(c1, c_), "cv2": self.conv(c1, c_), "cv3": self.conv(2 * c_, c2),
                "m": [self.bottleneck(c_, c_, shortcut, (k, k), 1.0) for _ in range(n)]}

    def c3k2(self, c1, c2, n, c3k=False, e=0.5, shortcut=True):
        c = int(c2 * e)
        m = [self.c3k(c, c, 2, shortcut) if c3k else self.bottleneck(c, c, shortcut) for _ in range(n)]
        return {"t": "C2f", "cv1": self.conv(c1, 2 * c), "cv2": self.conv((2 + n) * c, c2), "m": m}

    def sppf(self, c1, c2, k=5):
        c_ = c1 // 2
        return {"t": "SPPF", "cv1": self.conv(c1, c_), "cv2": self.conv(4 * c_, c2), "k": k}

    def attention(self, dim, num_heads, attn_ratio=0.5):
        head_dim = dim // num_heads
        key_dim = int(head_dim * attn_ratio)
        h = dim + key_dim * num_heads * 2
        return {"t": "Attention", "num_heads": num_heads, "head_dim": head_dim, "key_dim": key_dim, "scale": key_dim ** -0.5,
                "qkv": self.conv(dim, h, 1, act=False, gain=1.5), "proj": self.conv(dim, dim, 1, act=False, gain=0.5),
                "pe": self.conv(dim, dim, 3, g=dim, act=False)}

    def c2psa(self, c1, n):
        c = int(c1 * 0.5)
        blocks = [{"t": "PSABlock", "attn": self.attention(c, max(c // 64, 1)),
                   "ffn": [self.conv(c, 2 * c), self.conv(2 * c, c, act=False, gain=0.5)], "add": True} for _ in range(n)]
        return {"t": "C2PSA", "c": c, "cv1": self.conv(c1, 2 * c), "cv2": self.conv(2 * c, c1), "m": blocks}

    def aattn(self, dim, num_heads, area):
        hd = dim // num_heads
        return {"t": "AAttn", "area": area, "num_heads": num_heads, "head_dim": hd,
                "qkv": self.conv(dim, 3 * hd * num_heads, 1, act=False, gain=1.5),
                "proj": self.conv(hd * num_heads, dim, 1, act=False, gain=0.5),
                "pe": self.conv(hd * num_heads, dim, 7, g=dim, act=False)}

    def a2c2f(self, c1, c2, n, a2, area, residual=False, mlp_ratio=2.0, e=0.5, shortcut=True):
        c_ = int(c2 * e)
        m = []
        for _ in range(n):
            if a2:
                m.append([{"t": "ABlock", "attn": self.aattn(c_, c_ // 32, area),
                           "mlp": [self.conv(c_, int(c_ * mlp_ratio)), self.conv(int(c_ * mlp_ratio), c_, act=False, gain=0.5)]}
                          for _ in range(2)])
            else:
                m.append(self.c3k(c_, c_, 2, shortcut))
        gamma = (0.01 + 0.05 * torch.rand((c2,), generator=self.g)) if (a2 and residual) else None
        return {"t": "A2C2f", "cv1": self.conv(c1, c_), "cv2": self.conv((1 + n) * c_, c2), "gamma": gamma, "m": m}

    def detect(self, chs: List[int], legacy: bool = False, segment: bool = False, nm: int = 32, npr: int = 256):
        nc, reg_max = self.nc, 16
        c2, c3 = max(16, chs[0] // 4, reg_max * 4), max(chs[0], min(nc, 100))
        cv2, cv3 = [], []
        for i, x in enumerate(chs):
            cv2.append([self.conv(x, c2, 3), self.conv(c2, c2, 3), self.conv(c2, 4 * reg_max, 1, act=False, gain=2.0)])
            if legacy:
                br = [self.conv(x, c3, 3), self.conv(c3, c3, 3)]
            else:
                br = [self.conv(x, x, 3, g=x), self.conv(x, c3, 1), self.conv(c3, c3, 3, g=c3), self.conv(c3, c3, 1)]
            last = self.conv(c3, nc, 1, act=False, gain=4.0)
            # logits of about -3 +- 2: a few dozen anchors of a 640-pixel image clear the callers' thresholds (0.25 / 0.6)
            # by margins far above the numerical noise, the rest stay below
            last["b"] = torch.full((nc,), -3.0)
            cv3.append(br + [last])
        head = {"t": "Detect", "nc": nc, "reg_max": reg_max, "stride": [8, 16, 32], "cv2": cv2, "cv3": cv3}
        if segment:                                   # Segment(nc, nm, npr): mask-coefficient branch + prototype net
            c4, npr = max(chs[0] // 4, nm), self.ch(npr)
            head.update(t="Segment", nm=nm,
                        cv4=[[self.conv(x, c4, 3), self.conv(c4, c4, 3), self.conv(c4, nm, 1, act=False, gain=1.0)] for x in chs],
                        proto={"cv1": self.conv(chs[0], npr, 3),
                               "upsample": {"w": torch.randn((npr, npr, 2, 2), generator=self.g) * (1.6 / math.sqrt(npr)),
                                            "b": torch.randn((npr,), generator=self.g) * 0.1},
                               "cv2": self.conv(npr, npr, 3), "cv3": self.conv(npr, nm, 1)})
        return head


def synthetic_tree(family: str, scale: str = "s", nc: int = 1, seed: int = 0, names: Optional[dict] = None,
                   a2_residual: Optional[bool] = None, mlp_ratio: Optional[float] = None, segment: bool = False) -> dict:
    """YOLO11 (`family` "11": C3k2 + C2PSA, DWConv class branch) or YOLO12 ("12": A2C2f area attention) detection model
    of the given yaml scale with seeded weights.  Layouts as published in ultralytics cfg/models/11/yolo11.yaml and
    cfg/models/12/yolo12.yaml (from memory; real checkpoints bring their own tree)."""
    s = _Synth(seed, scale, nc)
    ch = s.ch
    layers: List[dict] = []
    width: List[int] = []

    def add(node, f=-1, c=None):
        node = dict(node)
        node["f"] = f
        layers.append(node)
        width.append(c if c is not None else out_channels(node))
        return len(layers) - 1

    c3k_all = scale in "mlx"
    add(s.conv(3, ch(64), 3, 2))
    add(s.conv(ch(64), ch(128), 3, 2))
    add(s.c3k2(ch(128), ch(256), s.rep(2), c3k_all, 0.25))
    add(s.conv(ch(256), ch(256), 3, 2))
    p3 = add(s.c3k2(ch(256), ch(512), s.rep(2), c3k_all, 0.25))
    add(s.conv(ch(512), ch(512), 3, 2))
    if family == "11":
        p4 = add(s.c3k2(ch(512), ch(512), s.rep(2), True))
        add(s.conv(ch(512), ch(1024), 3, 2))
        add(s.c3k2(ch(1024), ch(1024), s.rep(2), True))
        add(s.sppf(ch(1024), ch(1024)))
        p5 = add(s.c2psa(ch(1024), s.rep(2)))

        def neck(c1, c2, c3k):
            return s.c3k2(c1, c2, s.rep(2), c3k or c3k_all)
    elif family == "12":
        res, ratio = (True, 1.2) if scale in "lx" else (False, 2.0)
        res = res if a2_residual is None else a2_residual
        ratio = ratio if mlp_ratio is None else mlp_ratio
        p4 = add(s.a2c2f(ch(512), ch(512), s.rep(4), True, 4, res, ratio))
        add(s.conv(ch(512), ch(1024), 3, 2))
        p5 = add(s.a2c2f(ch(1024), ch(1024), s.rep(4), True, 1, res, ratio))

        def neck(c1, c2, c3k):
            return s.c3k2(c1, c2, s.rep(2), True) if c3k else s.a2c2f(c1, c2, s.rep(2), False, -1)
    else:
        raise ValueError(f"unknown family {family!r}")
    add({"t": "Upsample", "scale": 2}, c=width[p5])
    add({"t": "Concat", "d": 1}, f=[-1, p4], c=width[p5] + width[p4])
    n4 = add(neck(width[-1], ch(512), False))
    add({"t": "Upsample", "scale": 2}, c=width[n4])
    add({"t": "Concat", "d": 1}, f=[-1, p3], c=width[n4] + width[p3])
    n3 = add(neck(width[-1], ch(256), False))
    add(s.conv(ch(256), ch(256), 3, 2))
    add({"t": "Concat", "d": 1}, f=[-1, n4], c=ch(256) + width[n4])
    m4 = add(neck(width[-1], ch(512), False))
    add(s.conv(ch(512), ch(512), 3, 2))
    add({"t": "Concat", "d": 1}, f=[-1, p5], c=ch(512) + width[p5])
    m5 = add(neck(width[-1], ch(1024), True))
    add(s.detect([width[n3], width[m4], width[m5]], segment=segment), f=[n3, m4, m5], c=0)
    return {"layers": layers, "names": names or {i: f"class{i}" for i in range(nc)}, "family": family, "scale": scale}


# ---- device executor --------------------------------------------------------------------------------------------------
def _pad16(c: int) -> int:
    """The conv epilogue stores 16 output channels at a time: every activation slot is a multiple of 16 channels wide."""
    return (c + 15) // 16 * 16


def _declare(l) -> None:
    _declare_yolo(l)
    if getattr(l, "_tree_declared", False):
        return
    from .sam2 import _declare as _declare_sam
    _declare_sam(l)
    vp, i32 = C.c_void_p, C.c_int
    l.mtb_dwconv.argtypes = [vp, vp] + [i32] * 9 + [vp, vp, i32, i32, vp]
    l.mtb_dwconv.restype = i32
    l._tree_declared = True


class YoloTreeB200(YoloB200):
    """Callable with the reference's call shape (`model(image_bgr, conf=, device=, verbose=, imgsz=[, retina_masks=])` ->
    [Results]) and `.names`, like the object `ultralytics.YOLO(path)` gives the reference.  Shares the public methods of the
    hard-wired YOLOv8-seg detector (`forward_letterboxed`, `detect`, `retina_masks`): the plan it builds has the same
    head tensors, so a tree with a `Segment` head can also serve as the speech-bubble detector of the page path."""

    def __init__(self, tree: dict, device: torch.device, *, precision: str = "bf16x3"):
        self.l = lib()
        _declare(self.l)
        self.device = device
        self.planes = 2 if precision == "bf16x3" else 1
        self.tree = tree
        self.names = dict(tree.get("names") or {})
        head = tree["layers"][-1]
        if head["t"] not in ("Detect", "Segment"):
            raise UnsupportedCheckpoint(f"last layer is {head['t']}, expected a Detect or Segment head")
        self.has_masks = head["t"] == "Segment"
        self.nc = int(head["nc"])
        self.cfg = {"nc": self.nc, "nm": int(head.get("nm", 0))}
        if int(head["reg_max"]) != 16:
            raise UnsupportedCheckpoint("DFL with reg_max != 16 is not supported")
        if len(head["cv2"]) != 3:
            raise UnsupportedCheckpoint(f"{len(head['cv2'])} detection levels (3 supported)")
        self._w: Dict[int, tuple] = {}
        self._plans: Dict[tuple, dict] = {}

    # ---- weights ---------------------------------------------------------------------------------------------
    def _conv_w(self, node: dict, cin: Optional[int] = None):
        """Plane-format weights.  Widths that are not multiples of 16 (the 1.2x MLP of YOLO12-l/x: int(384 * 1.2) = 460) are
        padded with zero output rows / zero input columns: the extra channels hold act(0) = 0 and meet zero weights.
        (Only for tensors that own their buffer; a 8-mod-16 wide slice INSIDE a concatenation — the nano scales — is refused.)"""
        key = id(node)
        if key not in self._w:
            w, b = node["w"].to(self.device), node["b"].to(self.device)
            co, ci = int(w.shape[0]), int(w.shape[1])
            cin = ci if cin is None else cin
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin - ci, 0, _pad16(co) - co))
            b = torch.nn.functional.pad(b, (0, _pad16(co) - co))
            self._w[key] = (P.conv_weight_to_planes(w.contiguous(), self.planes), P.pad_bias(b, w.shape[0]))
        return self._w[key]

    def _detect_w(self, node: dict, cin: int, ocp: int):
        """Final 1x1 conv of a head branch: output rows as they are (the fp32 head tensor is `ocp` wide)."""
        key = id(node)
        if key not in self._w:
            w, b = node["w"].to(self.device), node["b"].to(self.device)
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin - int(w.shape[1])))
            self._w[key] = (P.conv_weight_to_planes(w.contiguous(), self.planes), P.pad_bias(b, w.shape[0]))
        return self._w[key]

    def _dw_w(self, node: dict):
        key = id(node)
        if key not in self._w:
            w = node["w"].to(self.device)                                  # [C][1][k][k]
            c, k = w.shape[0], w.shape[-1]
            self._w[key] = (w.reshape(c, k * k).t().contiguous(), node["b"].to(self.device).contiguous())   # [k*k][C]
        return self._w[key]

    def _deconv_tree_w(self, node: dict):
        """ConvTranspose2d(k=2, s=2) of the prototype branch as a 1x1 conv to 4 * Cout channels + pixel-shuffle store."""
        key = id(node)
        if key not in self._w:
            w = node["w"].to(self.device)                                  # [Cin][Cout][2][2]
            cin, cout = int(w.shape[0]), int(w.shape[1])
            w1 = w.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1).contiguous()
            self._w[key] = (P.conv_weight_to_planes(w1, self.planes), P.pad_bias(node["b"].to(self.device).repeat(4), 4 * cout))
        return self._w[key]

    # ---- plan ------------------------------------------------------------------------------------------------
    def _build(self, n: int, h: int, w: int) -> dict:
        dev, pl = self.device, self.planes
        steps: List[tuple] = []
        keep: List[torch.Tensor] = []

        def buf(hh, ww, ch):
            if ch % 16:
                raise UnsupportedCheckpoint(f"a tensor with {ch} channels: widths must be multiples of 16 (the s / m / l / x "
                                            "scales; the nano scale has 8-channel slices and is not supported)")
            t = torch.zeros((pl, n, hh, ww, ch), dtype=torch.bfloat16, device=dev)
            keep.append(t)
            return t

        def new(hh, ww, ch):
            return _Slice(buf(hh, ww, ch), 0, ch)

        def conv(node, src: _Slice, hh, ww, dst: Optional[_Slice] = None, residual: Optional[_Slice] = None, scale=None):
            """-> (dst slice, ho, wo)"""
            co, k, s, p, g = int(node["w"].shape[0]), node["k"], node["s"], node["p"], node["g"]
            ho, wo = (hh + 2 * p - k) // s + 1, (ww + 2 * p - k) // s + 1
            if g == 1:
                co = _pad16(co)
            if dst is None:
                dst = new(ho, wo, co)
            if dst.c != co:
                raise UnsupportedCheckpoint(f"a {co}-channel layer writes into a {dst.c}-channel slot")
            if g == 1:
                ci = int(node["w"].shape[1])
                if src.c not in (ci, _pad16(ci)):
                    raise UnsupportedCheckpoint(f"conv expects {ci} input channels, got {src.c}")
                wgt = self._conv_w(node, src.c)
                steps.append(("conv", ConvPlan(src.buf, wgt[0], wgt[1], dst.buf, k=k, stride=s, pad=p,
                                               act="silu" if node["act"] else None, x_coff=src.off, out_coff=dst.off,
                                               residual=None if residual is None else residual.buf,
                                               res_coff=0 if residual is None else residual.off, channel_scale=scale)))
            elif g == co == src.c and s == 1 and p == k // 2 and residual is None and scale is None:
                wgt = self._dw_w(node)
                steps.append(("dw", (src, dst, hh, ww, co, k, wgt[0], wgt[1], int(node["act"]))))
            else:
                raise UnsupportedCheckpoint(f"grouped convolution (groups {g}, {src.c} -> {co}, stride {s}) is not supported")
            return dst, ho, wo

        def seq(nodes, src, hh, ww, dst=None):
            for i, nd in enumerate(nodes):
                src, hh, ww = conv(nd, src, hh, ww, dst if i == len(nodes) - 1 else None)
            return src, hh, ww

        def bottleneck(node, src, hh, ww, dst=None):
            t, _, _ = conv(node["cv1"], src, hh, ww)
            return conv(node["cv2"], t, hh, ww, dst, residual=src if node["add"] else None)[0]

        def c3(node, src, hh, ww, dst=None):
            c_ = int(node["cv1"]["w"].shape[0])
            cat = buf(hh, ww, 2 * c_)
            cur, _, _ = conv(node["cv1"], src, hh, ww)
            for i, b in enumerate(node["m"]):
                cur = block(b, cur, hh, ww, _Slice(cat, 0, c_) if i == len(node["m"]) - 1 else None)
            if not node["m"]:
                steps.append(("copy", (cur, _Slice(cat, 0, c_))))
            conv(node["cv2"], src, hh, ww, _Slice(cat, c_, c_))
            return conv(node["cv3"], _Slice(cat, 0, 2 * c_), hh, ww, dst)[0]

        def c2f(node, src, hh, ww, dst=None):
            c = int(node["cv1"]["w"].shape[0]) // 2
            nb = len(node["m"])
            cat = buf(hh, ww, (2 + nb) * c)
            conv(node["cv1"], src, hh, ww, _Slice(cat, 0, 2 * c))
            for i, b in enumerate(node["m"]):
                block(b, _Slice(cat, (1 + i) * c, c), hh, ww, _Slice(cat, (2 + i) * c, c))
            return conv(node["cv2"], _Slice(cat, 0, (2 + nb) * c), hh, ww, dst)[0]

        def sppf(node, src, hh, ww, dst=None):
            c_ = int(node["cv1"]["w"].shape[0])
            sp = buf(hh, ww, 4 * c_)
            conv(node["cv1"], src, hh, ww, _Slice(sp, 0, c_))
            for i in range(3):
                steps.append(("maxpool", (sp, hh, ww, 4 * c_, i * c_, (i + 1) * c_, c_, int(node["k"]))))
            return conv(node["cv2"], _Slice(sp, 0, 4 * c_), hh, ww, dst)[0]

        def attention_core(qkv: _Slice, hh, ww, heads, kd, hd, scale, area, pe_node):
            """qkv: per token and head [q(kd) k(kd) v(hd)] -> softmax(q^T k * scale) applied to v, + pe(v) as an image."""
            if kd > hd or hd % 8 or kd % 8:
                raise UnsupportedCheckpoint(f"attention with key_dim {kd} / head_dim {hd}")
            ntok = hh * ww
            if area > 1 and ntok % area:
                raise UnsupportedCheckpoint(f"{hh}x{ww} tokens do not divide into {area} areas")
            per = 2 * kd + hd
            assert qkv.off == 0 and qkv.c == heads * per
            q, k, v = buf(hh, ww, heads * hd), buf(hh, ww, heads * hd), buf(hh, ww, heads * hd)   # q / k zero-padded to hd
            steps.append(("split_qkv", (qkv.buf, q, k, v, heads, kd, hd)))
            o = buf(hh, ww, heads * hd)
            a = max(area, 1)
            steps.append(("attn", dict(q=q, k=k, v=v, out=o, heads=heads, hd=hd, scale=float(scale), B=n * a, n=ntok // a)))
            pe, _, _ = conv(pe_node, _Slice(v, 0, heads * hd), hh, ww)
            s = buf(hh, ww, heads * hd)
            steps.append(("add", (o, pe.buf, s)))
            return _Slice(s, 0, heads * hd)

        def psablock(node, src, hh, ww, dst=None):
            a = node["attn"]
            qkv, _, _ = conv(a["qkv"], src, hh, ww)
            x = attention_core(qkv, hh, ww, a["num_heads"], a["key_dim"], a["head_dim"], a["scale"], 1, a["pe"])
            x1, _, _ = conv(a["proj"], x, hh, ww, residual=src if node["add"] else None)
            t, _, _ = conv(node["ffn"][0], x1, hh, ww)
            return conv(node["ffn"][1], t, hh, ww, dst, residual=x1 if node["add"] else None)[0]

        def c2psa(node, src, hh, ww, dst=None):
            c = int(node["c"])
            ab = buf(hh, ww, 2 * c)
            conv(node["cv1"], src, hh, ww, _Slice(ab, 0, 2 * c))
            cur = _Slice(ab, c, c)
            for i, b in enumerate(node["m"]):
                last = i == len(node["m"]) - 1
                nxt = psablock(b, cur, hh, ww, None)
                if last:
                    steps.append(("copy", (nxt, _Slice(ab, c, c))))
                cur = nxt
            return conv(node["cv2"], _Slice(ab, 0, 2 * c), hh, ww, dst)[0]

        def ablock(node, src, hh, ww, dst=None):
            a = node["attn"]
            hd, heads = a["head_dim"], a["num_heads"]
            qkv, _, _ = conv(a["qkv"], src, hh, ww)
            x = attention_core(qkv, hh, ww, heads, hd, hd, hd ** -0.5, int(a["area"]), a["pe"])
            x1, _, _ = conv(a["proj"], x, hh, ww, residual=src)
            t, _, _ = conv(node["mlp"][0], x1, hh, ww)
            return conv(node["mlp"][1], t, hh, ww, dst, residual=x1)[0]

        def a2c2f(node, src, hh, ww, dst=None):
            c_ = int(node["cv1"]["w"].shape[0])
            nb = len(node["m"])
            cat = buf(hh, ww, (1 + nb) * c_)
            conv(node["cv1"], src, hh, ww, _Slice(cat, 0, c_))
            for i, m in enumerate(node["m"]):
                cur, out = _Slice(cat, i * c_, c_), _Slice(cat, (i + 1) * c_, c_)
                if isinstance(m, list):
                    for j, ab in enumerate(m):
                        cur = ablock(ab, cur, hh, ww, out if j == len(m) - 1 else None)
                else:
                    block(m, cur, hh, ww, out)
            if node.get("gamma") is not None:                  # x + gamma * cv2(cat): the scale applies AFTER cv2's SiLU
                g = node["gamma"].to(dev).float().contiguous()
                co = int(node["cv2"]["w"].shape[0])
                if src.c != co:
                    raise UnsupportedCheckpoint("A2C2f residual with different input and output widths")
                y, _, _ = conv(node["cv2"], _Slice(cat, 0, (1 + nb) * c_), hh, ww)
                if dst is None:
                    dst = new(hh, ww, co)
                steps.append(("scale_add", (src, y, g, dst)))
                return dst
            return conv(node["cv2"], _Slice(cat, 0, (1 + nb) * c_), hh, ww, dst)[0]

        def block(node, src, hh, ww, dst=None) -> _Slice:
            t = node["t"]
            if t == "Conv":
                return conv(node, src, hh, ww, dst)[0]
            fn = {"Bottleneck": bottleneck, "C3": c3, "C2f": c2f, "SPPF": sppf, "C2PSA": c2psa, "A2C2f": a2c2f,
                  "PSABlock": psablock, "ABlock": ablock}.get(t)
            if fn is None:
                raise UnsupportedCheckpoint(f"module {t} is not supported")
            return fn(node, src, hh, ww, dst)

        x_in = torch.zeros((pl, n, h, w, 8), dtype=torch.bfloat16, device=dev)    # RGB + 5 zero channels (16-byte pixels)
        keep.append(x_in)
        proto = None
        outs: List[Tuple[_Slice, int, int]] = []
        cur: Tuple[_Slice, int, int] = (_Slice(x_in, 0, 3), h, w)
        levels = None
        for li, node in enumerate(self.tree["layers"]):
            f = node.get("f", -1)
            srcs = [cur if j == -1 else outs[j] for j in (f if isinstance(f, (list, tuple)) else [f])]
            t = node["t"]
            if t == "Concat":
                if int(node.get("d", 1)) != 1:
                    raise UnsupportedCheckpoint("Concat along a dimension other than channels")
                hh, ww = srcs[0][1], srcs[0][2]
                if any((s[1], s[2]) != (hh, ww) for s in srcs):
                    raise UnsupportedCheckpoint("Concat of maps with different sizes")
                cat = buf(hh, ww, sum(s[0].c for s in srcs))
                off = 0
                for s in srcs:
                    steps.append(("copy", (s[0], _Slice(cat, off, s[0].c))))
                    off += s[0].c
                res = (_Slice(cat, 0, off), hh, ww)
            elif t == "Upsample":
                if int(node.get("scale", 2)) != 2:
                    raise UnsupportedCheckpoint("Upsample by a factor other than 2")
                s, hh, ww = srcs[0]
                d = new(2 * hh, 2 * ww, s.c)
                steps.append(("up", (s, d, hh, ww)))
                res = (d, 2 * hh, 2 * ww)
            elif t in ("Detect", "Segment"):
                levels = []
                ncp = P.pad_to(self.nc, 16)
                nm = int(node.get("nm", 0))
                for i, (s, hh, ww) in enumerate(srcs):
                    outs_l = []
                    branches = [(node["cv2"][i], 64, 64), (node["cv3"][i], self.nc, ncp)]
                    if t == "Segment":
                        if nm % 16:
                            raise UnsupportedCheckpoint(f"{nm} mask coefficients (a multiple of 16 expected)")
                        branches.append((node["cv4"][i], nm, nm))
                    for br, oc, ocp in branches:
                        a, _, _ = seq(br[:-1], s, hh, ww)
                        o = torch.zeros((n, hh, ww, ocp), dtype=torch.float32, device=dev)
                        keep.append(o)
                        last = br[-1]
                        if last["g"] != 1 or last["k"] != 1 or int(last["w"].shape[0]) != oc:
                            raise UnsupportedCheckpoint("unexpected final layer in a Detect branch")
                        if a.c != _pad16(int(last["w"].shape[1])):
                            raise UnsupportedCheckpoint("Detect branch width mismatch")
                        wgt = self._detect_w(last, a.c, ocp)
                        steps.append(("conv", ConvPlan(a.buf, wgt[0], wgt[1], o, k=1, act="silu" if last["act"] else None,
                                                       x_coff=a.off)))
                        outs_l.append(o)
                    levels.append((outs_l[0], outs_l[1], outs_l[2] if t == "Segment" else None, hh, ww, int(node["stride"][i])))
                if t == "Segment":                              # Proto: cv1 3x3 -> ConvTranspose 2x -> cv2 3x3 -> cv3 1x1 (SiLU)
                    pr = node["proto"]
                    s0, hh, ww = srcs[0]
                    pa, _, _ = conv(pr["cv1"], s0, hh, ww)
                    npr = int(pr["upsample"]["w"].shape[1])
                    if (4 * npr) % 64 or int(pr["upsample"]["w"].shape[0]) != pa.c:
                        raise UnsupportedCheckpoint(f"prototype branch with {npr} channels")
                    pb = new(2 * hh, 2 * ww, npr)
                    wgt = self._deconv_tree_w(pr["upsample"])
                    steps.append(("conv", ConvPlan(pa.buf, wgt[0], wgt[1], pb.buf, k=1, act=None, pixel_shuffle=True)))
                    pc, _, _ = conv(pr["cv2"], pb, 2 * hh, 2 * ww)
                    proto = torch.zeros((n, 2 * hh, 2 * ww, nm), dtype=torch.float32, device=dev)
                    keep.append(proto)
                    last = pr["cv3"]
                    wgt = self._detect_w(last, pc.c, nm)
                    steps.append(("conv", ConvPlan(pc.buf, wgt[0], wgt[1], proto, k=1, act="silu" if last["act"] else None,
                                                   x_coff=pc.off)))
                res = (None, 0, 0)
            else:
                s, hh, ww = srcs[0]
                if t == "Conv":
                    res = conv(node, s, hh, ww)
                else:
                    res = (block(node, s, hh, ww), hh, ww)
            outs.append(res)
            cur = res
        if levels is None:
            raise UnsupportedCheckpoint("no Detect / Segment head")
        total_anchors = sum(fh * fw for (_, _, _, fh, fw, _) in levels)
        max_cand = min(total_anchors, 30000)
        return dict(steps=steps, keep=keep, x_in=x_in, levels=levels, proto=proto, layer_outputs=outs, ncp=P.pad_to(self.nc, 16),
                    max_cand=max_cand,
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
            if len(self._plans) >= 4:                       # page shapes vary: keep the plans of the last few
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = self._build(n, h, w)
        return self._plans[key]

    def _run_graph(self, g: dict) -> None:
        from .sam2 import AttnDesc
        l, st, pl = self.l, stream_ptr(), self.planes
        for kind, a in g["steps"]:
            if kind == "conv":
                a.run()
            elif kind == "dw":
                src, dst, hh, ww, c, k, wd, bd, act = a
                n = src.buf.shape[1]
                check(l.mtb_dwconv(ptr(src.buf), ptr(dst.buf), n, hh, ww, src.buf.shape[-1], src.off, dst.buf.shape[-1], dst.off,
                                   c, k, ptr(wd), ptr(bd), act, pl, st), "mtb_dwconv")
            elif kind == "copy":
                src, dst = a
                dst.buf[..., dst.off:dst.off + dst.c].copy_(src.buf[..., src.off:src.off + src.c])
            elif kind == "maxpool":
                t, hh, ww, ct, ci, co, c, k = a
                check(l.mtb_maxpool(ptr(t), ptr(t), t.shape[1], hh, ww, ct, ci, ct, co, c, k, pl, st), "mtb_maxpool")
            elif kind == "up":
                src, dst, hh, ww = a
                check(l.mtb_upsample2x(ptr(src.buf), ptr(dst.buf), src.buf.shape[1], hh, ww, src.buf.shape[-1], src.off,
                                       dst.buf.shape[-1], dst.off, src.c, pl, st), "mtb_upsample2x")
            elif kind == "split_qkv":
                qkv, q, k, v, heads, kd, hd = a
                lead = qkv.shape[:-1]
                t = qkv.view(*lead, heads, 2 * kd + hd)
                q.view(*lead, heads, hd)[..., :kd].copy_(t[..., :kd])
                k.view(*lead, heads, hd)[..., :kd].copy_(t[..., kd:2 * kd])
                v.view(*lead, heads, hd).copy_(t[..., 2 * kd:])
            elif kind == "scale_add":                           # dst = x + gamma[c] * y on plane tensors, in fp32
                x, y, gamma, dst = a
                xs, ys = x.buf[..., x.off:x.off + x.c], y.buf[..., y.off:y.off + y.c]
                r = xs.float().sum(0) + gamma * ys.float().sum(0)
                hi = r.to(torch.bfloat16)
                dst.buf[0, ..., dst.off:dst.off + dst.c].copy_(hi)
                if pl == 2:
                    dst.buf[1, ..., dst.off:dst.off + dst.c].copy_((r - hi.float()).to(torch.bfloat16))
            elif kind == "add":
                x, y, o = a
                rows = x[0].numel() // x.shape[-1]
                check(l.mtb_add_planes(ptr(x), ptr(y), ptr(o), rows, x.shape[-1], rows, pl, st), "mtb_add_planes")
            elif kind == "attn":
                d = AttnDesc()
                q, k, v, o = a["q"], a["k"], a["v"], a["out"]
                d.heads, d.hd, d.scale = a["heads"], a["hd"], a["scale"]
                d.q, d.k, d.v, d.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
                d.q_ct = d.k_ct = d.v_ct = d.o_ct = q.shape[-1]
                d.q_ps = d.k_ps = d.v_ps = d.o_ps = q[0].numel()
                d.planes, d.mode, d.B, d.nq, d.nk = pl, 0, a["B"], a["n"], a["n"]
                check(l.mtb_attention(C.byref(d), st), "mtb_attention")
            else:
                raise RuntimeError(kind)

    # ---- public (forward_letterboxed, detect and retina_masks are the primary detector's) ---------------------
    def __call__(self, image_bgr: np.ndarray, conf: float = 0.25, device=None, verbose: bool = False, imgsz: int = 640,
                 retina_masks: bool = True, **_):
        """Reference call shape (core/image/detection.py:1864-1870 panels, :140-146 OSB text, :1338-1345 bubbles)."""
        from .preproc import letterbox_device
        h0, w0 = image_bgr.shape[:2]
        img = torch.from_numpy(np.ascontiguousarray(image_bgr)).to(self.device)
        lb = letterbox_device(img, imgsz, swap_rb=True)
        g = self.forward_letterboxed(lb)
        det, cnt, _ = self.detect(g, conf, (h0, w0), tuple(lb.shape[:2]), apply_reference_dedup=False)
        n = int(cnt[0].item())
        d = det[:n].clone()
        boxes = Boxes(d[:, :4].contiguous(), d[:, 4].contiguous(), d[:, 5].contiguous()) if n else None
        masks = None
        if n and self.has_masks and retina_masks:
            masks = Masks(self.retina_masks(g, det, None, n, (h0, w0), tuple(lb.shape[:2])).float())
        return [Results(boxes, masks, (h0, w0), self.names)]
