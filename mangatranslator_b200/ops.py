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

    def run(self) -> None:
        check(lib().mtb_conv_plan_run(self._h, stream_ptr()), "mtb_conv_plan_run")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().mtb_conv_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
