"""Thin Python wrappers over the C-ABI plan objects (device pointers + stream in, nothing computed here)."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import ConvDesc, check, lib, ptr, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "silu": 2, "gelu": 3, "sigmoid": 4}


class ConvPlan:
    """One tcgen05 implicit-GEMM conv / linear layer bound to fixed device buffers.

    x: bf16 [planes_in][N][H][W][Cin]; w: bf16 [planes_in][KH*KW][Cout][Cin]; out: bf16 [planes_out][N][Ho][Wo][Cout]
    (or fp32 [N][Ho][Wo][Cout] when ``out`` is float32).
    """

    def __init__(self, x, w, bias, out, *, k=1, stride=1, pad=0, act=None, residual=None, tile_sums=None,
                 tile=None, mode=0, pixel_shuffle=False, x_coff=0, out_coff=0, res_coff=0, res_bcast=False, act_after_res=False,
                 channel_scale=None):
        """Channel slices: `x` / `out` / `residual` may be wider (concat) tensors; the layer reads channels
        [x_coff, x_coff + Cin) and writes [out_coff, out_coff + Cout).  The weights' Cin is padded to a multiple of
        64 with zeros, so whatever lies beyond the slice (or beyond the tensor: TMA zero-fills) contributes 0.
        `channel_scale` (fp32 [Cout], device): out = act((conv + bias) * scale) + residual, read when the plan runs."""
        planes_in, n, h, wd, x_ctotal = x.shape
        assert w.shape[0] == planes_in and w.shape[1] == k * k, (w.shape, x.shape)
        cout, cin = w.shape[2], w.shape[3]
        d = ConvDesc()
        d.N, d.H, d.W, d.Cin, d.Cout = n, h, wd, cin, cout
        d.KH = d.KW = k
        d.stride, d.pad = stride, pad
        d.planes_in = planes_in
        d.planes_out = 4 if out.dtype == torch.float32 else out.shape[0]
        d.act = ACT[act]
        d.res_planes = 0 if residual is None else residual.shape[0]
        d.tile_w, d.tile_h = tile if tile else (0, 0)
        d.x_ctotal, d.x_coff = x_ctotal, x_coff
        if pixel_shuffle:
            d.out_ctotal, d.out_coff = 0, 0
        else:
            d.out_ctotal, d.out_coff = out.shape[-1], out_coff
        if residual is not None:
            d.res_ctotal, d.res_coff = residual.shape[-1], res_coff
        d.res_bcast, d.act_after_res = int(bool(res_bcast)), int(bool(act_after_res))
        d.mode = mode
        d.pixel_shuffle = int(bool(pixel_shuffle))
        self._keep = (x, w, bias, out, residual, tile_sums)
        h = C.c_void_p()
        check(lib().mtb_conv_plan_create(C.byref(d), ptr(x), ptr(w), ptr(bias), ptr(out), ptr(residual),
                                         ptr(tile_sums), C.byref(h)), "mtb_conv_plan_create")
        self._h = h
        self.out = out
        if channel_scale is not None:
            assert channel_scale.dtype == torch.float32 and channel_scale.numel() >= cout
            self._keep = self._keep + (channel_scale,)
            check(lib().mtb_conv_plan_set_channel_scale(self._h, ptr(channel_scale)), "mtb_conv_plan_set_channel_scale")

    @property
    def num_mtiles(self) -> int:
        return lib().mtb_conv_plan_num_mtiles(self._h)

    @property
    def num_sum_rows(self) -> int:
        """Rows of the tile_sums buffer ([rows][Cout] fp32) this plan writes."""
        return lib().mtb_conv_plan_num_sum_rows(self._h)

    def set_border_sums(self, border) -> bool:
        """Ask the layer to also emit per-channel sums of its output over the four image border lines
        ([num_sum_rows][4][Cout] fp32).  Returns False when the kernel that runs this layer cannot provide them."""
        rc = lib().mtb_conv_plan_set_border_sums(self._h, ptr(border))
        if rc == 0:
            self._keep = self._keep + (border,)
        return rc == 0

    def set_fixed_sums(self, fixed) -> None:
        """fp16c plans: accumulate the channel / border sums into `fixed` (int64 [5][64], 2^-20 units) with integer atomics
        instead of per-CTA rows (set_border_sums must have been called)."""
        assert fixed.dtype == torch.int64 and fixed.numel() >= 320
        l = lib()
        l.mtb_conv_plan_set_fixed_sums.argtypes = [C.c_void_p, C.c_void_p]
        l.mtb_conv_plan_set_fixed_sums.restype = C.c_int
        check(l.mtb_conv_plan_set_fixed_sums(self._h, ptr(fixed)), "mtb_conv_plan_set_fixed_sums")
        self._keep = self._keep + (fixed,)

    def set_fused_gate(self, fixed_in, fixed_zero, u, conv_w, conv_b, w1, b1, w2, b2) -> None:
        """fp16c plans, an RCAB's second conv: compute the CALayer gate in this launch's prologue from the fixed-point sums
        the first conv accumulated (`fixed_in`), zero `fixed_zero` for the next block (include/mtb200.h)."""
        l = lib()
        l.mtb_conv_plan_set_fused_gate.argtypes = [C.c_void_p] * 10 + [C.c_int]
        l.mtb_conv_plan_set_fused_gate.restype = C.c_int
        check(l.mtb_conv_plan_set_fused_gate(self._h, ptr(fixed_in), ptr(fixed_zero), ptr(u), ptr(conv_w), ptr(conv_b), ptr(w1),
                                             ptr(b1), ptr(w2), ptr(b2), int(w1.shape[0])), "mtb_conv_plan_set_fused_gate")
        self._keep = self._keep + (fixed_in, fixed_zero, u, conv_w, conv_b, w1, b1, w2, b2)

    def run(self) -> None:
        check(lib().mtb_conv_plan_run(self._h, stream_ptr()), "mtb_conv_plan_run")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().mtb_conv_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


class RcanConvPlan(ConvPlan):
    """RCAN body layer (3x3, 64 -> 64) in the fp16c format (one fp16 product + an e5m2 correction product per tap).

    x / out / residual: uint8 [3][N][H][W][64] (planes.nhwc_to_fp16c); w: uint8 [2880][64] (planes.conv_weight_to_fp16c).
    Shares run / channel scale / border sums / sum rows with ConvPlan (same plan object underneath)."""

    def __init__(self, x, w, bias, out, *, act=None, residual=None, tile_sums=None, channel_scale=None, lo_shift=0):
        for t in (x, out) + ((residual,) if residual is not None else ()):
            assert t.dtype == torch.uint8 and t.dim() == 5 and t.shape[0] == 3 and t.shape[-1] == 64 and t.is_contiguous()
        assert out.shape == x.shape and (residual is None or residual.shape == x.shape)
        assert w.dtype == torch.uint8 and tuple(w.shape) == (2880, 64) and w.is_contiguous()
        _, n, h, wd, _ = x.shape
        l = lib()
        l.mtb_rcan_conv_plan_create.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int] * 2 + [C.POINTER(C.c_void_p)]
        l.mtb_rcan_conv_plan_create.restype = C.c_int
        self._keep = (x, w, bias, out, residual, tile_sums)
        hd = C.c_void_p()
        check(l.mtb_rcan_conv_plan_create(n, h, wd, ptr(x), ptr(w), ptr(bias), ptr(out), ptr(residual), ptr(tile_sums),
                                          ACT[act], lo_shift, C.byref(hd)), "mtb_rcan_conv_plan_create")
        self._h = hd
        self.out = out
        if channel_scale is not None:
            assert channel_scale.dtype == torch.float32 and channel_scale.numel() >= 64
            self._keep = self._keep + (channel_scale,)
            check(l.mtb_conv_plan_set_channel_scale(self._h, ptr(channel_scale)), "mtb_conv_plan_set_channel_scale")
