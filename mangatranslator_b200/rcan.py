"""B200 RCAN upscaler (2x-AnimeSharpV4 architecture) — the object `ModelManager.load_upscale()` returns.

Mirrors what the reference gets from spandrel (core/ml/model_manager.py:617-657): a callable
`model(x: float32 (1,3,h,w) in [0,1]) -> (1,3,2h,2w)`; additionally `upscale_u8` keeps the page on the device as
uint8 (what the batch pipeline uses).  Every convolution is a tcgen05 plan (bf16x3 by default = fp32-grade); the
~400 RCAB body convs use the halo-tile kernels (conv_halo_cm.cu / conv_halo.cu).  Buffers are allocated once per
input size.

RCAB (spandrel RCAN: conv -> ReLU -> conv -> CALayer -> + x) runs as three launches:
    u   = relu(conv1(x))                       halo conv, epilogue also emits per-channel sums of u
    g   = gate(mean(conv2(u)))                 mtb_rcan_gate: the mean follows from sums of u (linearity), see the header
    x'  = x + g * conv2(u)                     halo conv with channel scale + residual in the epilogue
so the block's output is written once and the attention-weighted tensor is never materialised.

The lite model (`load_upscale_lite`, 2x-AnimeSharpV4_Fast_RCAN_PU) is the same network behind a PixelUnshuffle: the
head sees 3*d*d channels of the page reflect-padded to a multiple of d, the body runs on 1/d^2 of the pixels, the
upsampler has log2(2d) conv+PixelShuffle(2) stages and the output is cropped to 2H x 2W.  d, the stage count and the
depth are read from the state dict (keys/shapes), as spandrel's loader does.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
from typing import Dict, Tuple

import torch

from . import planes as P
from ._lib import check, lib, ptr, stream_ptr
from .ops import ConvPlan, RcanConvPlan


def _declare(l) -> None:
    if getattr(l, "_ew_declared", False):
        return
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    l.mtb_image_to_planes.argtypes = [vp, i32, i32, i32, i32, f32, C.POINTER(f32), vp, i32, i32, vp]
    l.mtb_ca_scale.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, vp, i32, vp, vp]
    l.mtb_scale_residual.argtypes = [vp, vp, vp, vp, C.c_longlong, i32, i32, i32, vp]
    l.mtb_rcan_gate.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    l.mtb_f32_to_u8.argtypes = [vp, C.c_longlong, i32, C.POINTER(f32), f32, vp, vp, vp]
    l.mtb_image_to_planes_unshuffle.argtypes = [vp, i32, i32, i32, i32, f32, C.POINTER(f32), i32, vp, i32, i32, vp]
    l.mtb_f32_to_u8_crop.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(f32), f32, vp, vp, vp]
    l.mtb_rcan_gate_fp16c.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    l.mtb_planes_bf16x2_to_fp16c.argtypes = [vp, C.c_longlong, vp, i32, vp]
    l.mtb_planes_fp16c_to_bf16x2.argtypes = [vp, C.c_longlong, vp, i32, vp]
    for n in ("mtb_image_to_planes", "mtb_ca_scale", "mtb_scale_residual", "mtb_f32_to_u8", "mtb_rcan_gate",
              "mtb_image_to_planes_unshuffle", "mtb_f32_to_u8_crop", "mtb_rcan_gate_fp16c", "mtb_planes_bf16x2_to_fp16c",
              "mtb_planes_fp16c_to_bf16x2"):
        getattr(l, n).restype = i32
    l._ew_declared = True


def infer_config(sd: Dict[str, torch.Tensor]) -> dict:
    """n_resgroups / n_resblocks / n_feats / reduction from state-dict keys and shapes (what spandrel's loader does)."""
    groups = {int(k.split(".")[1]) for k in sd if k.startswith("body.") and k.count(".") >= 4}
    blocks = {int(k.split(".")[3]) for k in sd if k.startswith("body.0.body.") and ".body." in k[12:]}
    f = sd["head.0.weight"].shape[0]
    red = f // sd["body.0.body.0.body.3.conv_du.0.weight"].shape[0]
    # "_PU" variants (spandrel RCAN unshuffle_mod): the head sees PixelUnshuffle(d) of the page = 3*d*d channels, and the
    # upsampler has one conv+PixelShuffle(2) stage more per factor of two (tail.0.0, tail.0.2, ...)
    in_ch = sd["head.0.weight"].shape[1]
    d = int(round((in_ch / 3) ** 0.5))
    if 3 * d * d != in_ch:
        raise ValueError(f"RCAN: head input channels {in_ch} are not 3*d*d")
    stages = sorted(int(k.split(".")[2]) for k in sd if k.startswith("tail.0.") and k.endswith(".weight"))
    for i in stages:
        if sd[f"tail.0.{i}.weight"].shape[0] != 4 * f:
            raise ValueError("RCAN: only PixelShuffle(2) upsampler stages are supported")
    scale = (2 ** len(stages)) // d
    if scale * d != 2 ** len(stages) or scale < 1:
        raise ValueError(f"RCAN: upsampler x{2 ** len(stages)} does not undo unshuffle x{d}")
    return dict(n_resgroups=max(groups) + 1, n_resblocks=max(blocks) + 1, n_feats=f, reduction=red, unshuffle=d,
                up_stages=stages, scale=scale)


class RcanB200:
    DIV2K_MEAN = (0.4488, 0.4371, 0.4040)

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: torch.device, *, precision: str | None = None,
                 rgb_range: float = 1.0, norm: bool = False, conv_mode: int = 0, lo_shift: int = 0):
        """precision: "fp16c" (default; MTB200_RCAN_PRECISION overrides) runs the 64->64 body convs as one fp16 product +
        an e5m2 correction product per tap (conv_halo_fp16c.cu; fp32-grade within the 1e-3 parity bound, 25 % fewer
        tensor-core cycles, 3 B per activation); "bf16x3" keeps them in bf16 hi/lo planes (the A/B partner, ~4e-5);
        "bf16" is a plain bf16 network (NOT parity grade).  Head, upsampler and tail run in bf16x3 for both."""
        precision = precision or os.environ.get("MTB200_RCAN_PRECISION", "fp16c")
        assert precision in ("fp16c", "bf16x3", "bf16")
        self.l = lib()
        _declare(self.l)
        self.device = device
        self.planes = 1 if precision == "bf16" else 2
        self.fp16c = precision == "fp16c"
        self.lo_shift = int(lo_shift)
        # channel sums for the RCAB gate as fixed-point integer atomics (default) or as per-CTA float rows (A/B partner)
        self.fixed_sums = os.environ.get("MTB200_RCAN_FIXED_SUMS", "1") != "0"
        # the gate computed in the prologue of the block's second conv (no launch of its own) or by mtb_rcan_gate_fp16c
        self.fused_gate = self.fixed_sums and os.environ.get("MTB200_RCAN_FUSED_GATE", "1") != "0"
        self.precision = precision
        self.cfg = infer_config(state_dict)
        self.rgb_range, self.norm, self.conv_mode = float(rgb_range), bool(norm), conv_mode
        f = self.cfg["n_feats"]
        if f != 64:
            raise ValueError(f"RcanB200: n_feats={f} not supported (kernels are built for 64 feature channels)")
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self.F = f
        pl = self.planes

        def conv_w(name, body=False):
            if body and self.fp16c:
                return (P.conv_weight_to_fp16c(sd[name + ".weight"], self.lo_shift),
                        P.pad_bias(sd.get(name + ".bias"), sd[name + ".weight"].shape[0]))
            return P.conv_weight_to_planes(sd[name + ".weight"], pl), P.pad_bias(sd.get(name + ".bias"), sd[name + ".weight"].shape[0])

        self.w_head = conv_w("head.0")
        G, R = self.cfg["n_resgroups"], self.cfg["n_resblocks"]
        self.blocks = []
        for g in range(G):
            grp = []
            for b in range(R):
                base = f"body.{g}.body.{b}.body"
                w1, w2 = conv_w(base + ".0", True), conv_w(base + ".2", True)
                cd1 = sd[base + ".3.conv_du.0.weight"].reshape(-1, f).contiguous()
                cb1 = sd[base + ".3.conv_du.0.bias"].contiguous()
                cd2 = sd[base + ".3.conv_du.2.weight"].reshape(f, -1).contiguous()
                cb2 = sd[base + ".3.conv_du.2.bias"].contiguous()
                w2f = sd[base + ".2.weight"].contiguous()                 # fp32 [64][64][3][3] for the gate's mean
                b2f = sd[base + ".2.bias"].contiguous() if (base + ".2.bias") in sd else None
                grp.append((w1, w2, cd1, cb1, cd2, cb2, w2f, b2f))
            self.blocks.append((grp, conv_w(f"body.{g}.body.{R}", True)))
        self.w_body_tail = conv_w(f"body.{G}", True)
        # upsampler conv: reorder output channels so the four PixelShuffle phases are contiguous 64-channel blocks
        idx = torch.arange(4 * f, device=device).view(f, 4).t().reshape(-1)   # new[(dy*2+dx)*f + c] = old[c*4 + dy*2+dx]
        self.w_up = []
        for i in self.cfg["up_stages"]:
            wu, bu = sd[f"tail.0.{i}.weight"], sd[f"tail.0.{i}.bias"]
            self.w_up.append((P.conv_weight_to_planes(wu[idx], pl), P.pad_bias(bu[idx], 4 * f)))
        self.w_tail = conv_w("tail.1")
        self.unshuffle, self.scale = self.cfg["unshuffle"], self.cfg["scale"]
        # plans (buffers + launch descriptors) per input size, least recently used first; crops make the sizes ragged
        self._plans: "collections.OrderedDict[Tuple[int, int], dict]" = collections.OrderedDict()
        self.max_plans = int(os.environ.get("MTB200_RCAN_MAX_PLANS", "12"))
        self.max_plan_bytes = int(float(os.environ.get("MTB200_RCAN_MAX_PLAN_GB", "48")) * 2 ** 30)

    # ---------------------------------------------------------------------------------------------------------
    def _build(self, h: int, w: int) -> dict:
        """h, w: size of the frame the conv body sees (the page, or the page / unshuffle factor rounded up)."""
        dev, pl, f = self.device, self.planes, self.F
        bf = torch.bfloat16
        n = 1
        nbytes = [0]

        def act(hh, ww, c=f):
            nbytes[0] += pl * n * hh * ww * c * 2
            return torch.zeros((pl, n, hh, ww, c), dtype=bf, device=dev)

        def cact(hh, ww):                      # fp16c byte planes of the body
            nbytes[0] += 3 * n * hh * ww * 64
            return torch.zeros((3, n, hh, ww, 64), dtype=torch.uint8, device=dev)

        bact = cact if self.fp16c else act
        b = dict(x_in=act(h, w), head=act(h, w), u=bact(h, w), pa=bact(h, w), pb=bact(h, w),
                 g0=bact(h, w), g1=bact(h, w), scale=torch.zeros((n, f), dtype=torch.float32, device=dev))
        steps = []
        mode = self.conv_mode

        def conv(x, wgt, out, **kw):
            return ConvPlan(x, wgt[0], wgt[1], out, k=3, pad=1, mode=kw.pop("mode", mode), **kw)

        def bconv(x, wgt, out, **kw):          # a 64 -> 64 body layer
            if self.fp16c:
                return RcanConvPlan(x, wgt[0], wgt[1], out, lo_shift=self.lo_shift, **kw)
            return conv(x, wgt, out, **kw)

        # head conv: only 3*d*d of the 64 padded input channels are non-zero (per-tap kernel)
        steps.append(("conv", conv(b["x_in"], self.w_head, b["head"], mode=1)))
        body_in = b["head"]
        if self.fp16c:
            b["head_c"] = cact(h, w)
            steps.append(("to_fp16c", (b["head"], b["head_c"], n * h * w)))
            body_in = b["head_c"]
        probe = bconv(body_in, self.blocks[0][0][0][0], b["u"], act="relu")
        parts = probe.num_sum_rows
        sums = torch.zeros((parts, f), dtype=torch.float32, device=dev)
        b["sums"] = sums
        # border-line sums of conv1's output come out of the same epilogue when the channel-major kernel runs the
        # layer (bf16x3); otherwise the gate kernel reads the four lines of u itself
        border = torch.zeros((parts, 4, f), dtype=torch.float32, device=dev)
        b["border"] = None
        # fp16c: fixed-point sums (integer atomics); two buffers alternate between consecutive blocks when the gate is fused
        b["fixed2"] = torch.zeros((2, 5, f), dtype=torch.int64, device=dev)
        b["fixed"] = b["fixed2"][0]
        src = body_in
        nblk = 0
        for gi, (grp, tailw) in enumerate(self.blocks):
            grp_in = src
            x = grp_in
            for (w1, w2, cd1, cb1, cd2, cb2, w2f, b2f) in grp:
                dst = b["pa"] if x is not b["pa"] else b["pb"]
                c1 = bconv(x, w1, b["u"], act="relu", tile_sums=sums)
                if c1.set_border_sums(border):
                    b["border"] = border
                fuse = self.fp16c and self.fused_gate and cd1.shape[0] <= 16
                fx = b["fixed2"][nblk % 2] if fuse else b["fixed"]
                if self.fp16c and self.fixed_sums:
                    c1.set_fixed_sums(fx)
                steps.append(("conv_body", c1))
                if fuse:
                    c2 = bconv(b["u"], w2, dst, residual=x)
                    c2.set_fused_gate(fx, b["fixed2"][(nblk + 1) % 2], b["u"], w2f, b2f, cd1, cb1, cd2, cb2)
                    steps.append(("conv_body", c2))
                else:
                    steps.append(("gate", (parts, w2f, b2f, cd1, cb1, cd2, cb2)))
                    steps.append(("conv_body", bconv(b["u"], w2, dst, residual=x, channel_scale=b["scale"])))
                nblk += 1
                x = dst
            gout = b["g0"] if grp_in is not b["g0"] else b["g1"]
            steps.append(("conv", bconv(x, tailw, gout, residual=grp_in)))   # group tail conv + group skip
            src = gout
        steps.append(("conv", bconv(src, self.w_body_tail, b["u"], residual=body_in)))
        # upsampler: conv F -> 4F with the PixelShuffle(2) folded into the store, once per stage
        cur, hh, ww = b["u"], h, w
        if self.fp16c:
            b["body_out"] = act(h, w)
            steps.append(("from_fp16c", (b["u"], b["body_out"], n * h * w)))
            cur = b["body_out"]
        for si, wu in enumerate(self.w_up):
            nxt = act(2 * hh, 2 * ww)
            b[f"up{si}"] = nxt
            steps.append(("conv", ConvPlan(cur, wu[0], wu[1], nxt, k=3, pad=1, pixel_shuffle=True, mode=1)))
            cur, hh, ww = nxt, 2 * hh, 2 * ww
        b["out32"] = torch.zeros((n, hh, ww, 16), dtype=torch.float32, device=dev)
        nbytes[0] += n * hh * ww * 16 * 4
        b["out_hw"] = (hh, ww)
        steps.append(("conv", ConvPlan(cur, self.w_tail[0], self.w_tail[1], b["out32"], k=3, pad=1, mode=1)))
        b["steps"] = steps
        b["nbytes"] = nbytes[0]
        b["uses"] = 0
        return b

    def body_hw(self, h: int, w: int) -> Tuple[int, int]:
        d = self.unshuffle
        return -(-h // d), -(-w // d)

    def _get(self, h: int, w: int) -> dict:
        """Plan for an input page of h x w (least-recently-used plans are dropped beyond max_plans / max_plan_bytes)."""
        key = (h, w)
        if key in self._plans:
            self._plans.move_to_end(key)
            return self._plans[key]
        plan = self._build(*self.body_hw(h, w))
        self._plans[key] = plan
        while len(self._plans) > 1 and (len(self._plans) > self.max_plans or
                                         sum(p["nbytes"] for p in self._plans.values()) > self.max_plan_bytes):
            self._plans.popitem(last=False)
        return plan

    def _run_body(self, b: dict, h: int, w: int) -> None:
        l, st = self.l, stream_ptr()
        if self.fp16c and self.fixed_sums:
            b["fixed2"].zero_()                # the gate leaves them zeroed; this only matters after an aborted pass
        for kind, arg in b["steps"]:
            self._run_step(b, kind, arg, h, w, st)

    def _run_step(self, b: dict, kind: str, arg, h: int, w: int, st) -> None:
        if kind in ("conv", "conv_body"):
            arg.run()
        elif kind == "gate":
            self._gate(b, arg, h, w, st)
        elif kind == "to_fp16c":
            check(self.l.mtb_planes_bf16x2_to_fp16c(ptr(arg[0]), arg[2], ptr(arg[1]), self.lo_shift, st), "mtb_planes_bf16x2_to_fp16c")
        elif kind == "from_fp16c":
            check(self.l.mtb_planes_fp16c_to_bf16x2(ptr(arg[0]), arg[2], ptr(arg[1]), self.lo_shift, st), "mtb_planes_fp16c_to_bf16x2")
        else:
            raise ValueError(kind)

    def _gate(self, b: dict, arg, h: int, w: int, st) -> None:
        parts, w2f, b2f, cd1, cb1, cd2, cb2 = arg
        if self.fp16c:
            if b["border"] is None:
                raise RuntimeError("RcanB200: the fp16c body needs the border sums of conv1's epilogue (one image per launch)")
            check(self.l.mtb_rcan_gate_fp16c(ptr(b["sums"]), parts, ptr(b["border"]), ptr(b["fixed"]) if self.fixed_sums else None,
                                             ptr(b["u"]), self.lo_shift, h, w, ptr(w2f),
                                             ptr(b2f), ptr(cd1), ptr(cb1), ptr(cd2), ptr(cb2), cd1.shape[0], ptr(b["scale"]), st),
                  "mtb_rcan_gate_fp16c")
            return
        check(self.l.mtb_rcan_gate(ptr(b["sums"]), parts, ptr(b["border"]), ptr(b["u"]), self.planes, h, w, ptr(w2f), ptr(b2f), ptr(cd1),
                                   ptr(cb1), ptr(cd2), ptr(cb2), cd1.shape[0], ptr(b["scale"]), st), "mtb_rcan_gate")

    def time_steps(self, img: torch.Tensor):
        """CUDA-event duration (ms) of every launch of one pass over `img`, as (kind, ms) in launch order; kinds:
        "conv_body" (RCAB 3x3 convs), "gate", "conv" (head / group tails / upsampler / tail)."""
        h, w, _ = img.shape
        b = self._get(h, w)
        hb, wb = self.body_hw(h, w)
        evs = []
        st = stream_ptr()
        if self.fp16c and self.fixed_sums:
            b["fixed2"].zero_()
        for kind, arg in b["steps"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._run_step(b, kind, arg, hb, wb, st)
            e1.record()
            evs.append((kind, e0, e1))
        torch.cuda.synchronize()
        return [(k, a.elapsed_time(c)) for k, a, c in evs]

    def time_body_convs(self, img: torch.Tensor):
        """CUDA-event duration (ms) of every RCAB body conv launch during one pass over `img` (bench roofline)."""
        return [ms for k, ms in self.time_steps(img) if k == "conv_body"]

    def _upscale_static(self, b: dict, h: int, w: int, c: int, swap_rb: bool, want_float: bool) -> None:
        """page in b['in_u8'] -> b['out_u8'] (and b['out_f']); only static buffers, so the sequence can be graph-captured."""
        l, st = self.l, stream_ptr()
        mean = [m * self.rgb_range for m in self.DIV2K_MEAN] if self.norm else [0.0, 0.0, 0.0]
        sub = (C.c_float * 3)(*mean)
        d, sc = self.unshuffle, self.scale
        hb, wb = self.body_hw(h, w)
        if d == 1:
            check(l.mtb_image_to_planes(ptr(b["in_u8"]), h, w, c, int(swap_rb), self.rgb_range / 255.0, sub, ptr(b["x_in"]),
                                        64, self.planes, st), "mtb_image_to_planes")
        else:
            check(l.mtb_image_to_planes_unshuffle(ptr(b["in_u8"]), h, w, c, int(swap_rb), self.rgb_range / 255.0, sub, d,
                                                  ptr(b["x_in"]), 64, self.planes, st), "mtb_image_to_planes_unshuffle")
        self._run_body(b, hb, wb)
        add = (C.c_float * 3)(*mean)
        oh, ow = b["out_hw"]
        if (oh, ow) == (sc * h, sc * w):
            check(l.mtb_f32_to_u8(ptr(b["out32"]), oh * ow, 16, add, 1.0 / self.rgb_range, ptr(b["out_u8"]),
                                  ptr(b["out_f"]) if want_float else None, st), "mtb_f32_to_u8")
        else:
            check(l.mtb_f32_to_u8_crop(ptr(b["out32"]), oh, ow, 16, sc * h, sc * w, add, 1.0 / self.rgb_range,
                                       ptr(b["out_u8"]), ptr(b["out_f"]) if want_float else None, st), "mtb_f32_to_u8_crop")

    def upscale_u8(self, img: torch.Tensor, *, swap_rb: bool = False, want_float: bool = False):
        """img: device uint8 HxWx(3|4).  Returns uint8 sH x sW x 3 (s = model scale, 2 for both AnimeSharp models; RGB
        model output, `swap_rb` feeds a BGR page), and optionally the float output before quantisation.  The returned
        tensors are the plan's static output buffers: consume (copy) them before the next call of the same size."""
        from . import graphs
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
        h, w, c = img.shape
        sc = self.scale
        b = self._get(h, w)
        key = ("io", c)
        if key not in b:
            b[key] = True
            b["in_u8"] = torch.empty((h, w, c), dtype=torch.uint8, device=self.device)
            b["out_u8"] = torch.empty((sc * h, sc * w, 3), dtype=torch.uint8, device=self.device)
            b["out_f"] = torch.empty((sc * h, sc * w, 3), dtype=torch.float32, device=self.device)
        b["in_u8"].copy_(img)
        b["uses"] += 1
        gkey = ("graph", c, bool(swap_rb), bool(want_float))
        # a size seen once (ragged bubble crops) runs eagerly; from its second use on the ~600 launches replay as a graph
        if graphs.ENABLED and (gkey in b or b["uses"] >= 2):
            if gkey not in b:
                b[gkey] = graphs.CapturedGraph(lambda: self._upscale_static(b, h, w, c, swap_rb, want_float))
            b[gkey].replay()
        else:
            self._upscale_static(b, h, w, c, swap_rb, want_float)
        return (b["out_u8"], b["out_f"]) if want_float else b["out_u8"]

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Reference call shape (core/image/image_utils.py:369-374): float32 (1,3,h,w) in [0,1] -> (1,3,sh,sw)."""
        assert x.dim() == 4 and x.shape[0] == 1 and x.shape[1] == 3
        h, w = x.shape[2], x.shape[3]
        d, sc = self.unshuffle, self.scale
        hb, wb = self.body_hw(h, w)
        b = self._get(h, w)
        xin = x.to(self.device, torch.float32) * self.rgb_range
        if self.norm:
            xin = xin - torch.tensor(self.DIV2K_MEAN, device=self.device).view(1, 3, 1, 1) * self.rgb_range
        if d > 1:
            xin = torch.nn.functional.pad(xin, (0, wb * d - w, 0, hb * d - h), mode="reflect")
            xin = torch.nn.functional.pixel_unshuffle(xin, d)
        b["x_in"].copy_(P.nchw_to_planes(xin, self.planes))
        self._run_body(b, hb, wb)
        y = b["out32"][:, :sc * h, :sc * w, :3].permute(0, 3, 1, 2)
        if self.norm:
            y = y + torch.tensor(self.DIV2K_MEAN, device=self.device).view(1, 3, 1, 1) * self.rgb_range
        return (y / self.rgb_range).contiguous()
