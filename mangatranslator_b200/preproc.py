"""Device pre-/post-processing wrappers (C ABI: mtb_letterbox_u8, mtb_resize_aa_u8, mtb_resize_lanczos_u8)."""
from __future__ import annotations

import collections
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr


def _declare(l) -> None:
    if getattr(l, "_pre_declared", False):
        return
    vp, i32 = C.c_void_p, C.c_int
    l.mtb_letterbox_u8.argtypes = [vp, i32, i32, i32, vp] + [i32] * 8 + [vp, vp]
    l.mtb_letterbox_u8.restype = i32
    l.mtb_resize_aa_u8.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, vp, C.c_longlong, vp]
    l.mtb_resize_aa_u8.restype = i32
    l.mtb_resize_lanczos_table_ints.argtypes = [i32] * 4
    l.mtb_resize_lanczos_table_ints.restype = C.c_longlong
    l.mtb_resize_lanczos_u8.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, vp, C.c_longlong, i32, vp]
    l.mtb_resize_lanczos_u8.restype = i32
    l.mtb_flatten_alpha_u8.argtypes = [vp, i32, i32, C.POINTER(i32), vp, vp]
    l.mtb_flatten_alpha_u8.restype = i32
    l._pre_declared = True


def letterbox_geometry(h0: int, w0: int, imgsz: int, stride: int = 32):
    """ultralytics LetterBox(auto=True, scaleup=True): ((nw, nh), (top, bottom, left, right), (H, W), gain)."""
    r = min(imgsz / h0, imgsz / w0)
    nw, nh = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = ((imgsz - nw) % stride) / 2, ((imgsz - nh) % stride) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return (nw, nh), (top, bottom, left, right), (nh + top + bottom, nw + left + right), r


def letterbox_device(img: torch.Tensor, imgsz: int, *, swap_rb: bool = True) -> torch.Tensor:
    """img: device uint8 HxWx(3|4) (BGR[A]) -> letterboxed uint8 H'xW'x3 (RGB when swap_rb)."""
    l = lib()
    _declare(l)
    h0, w0, c = img.shape
    (nw, nh), (top, _, left, _), (oh, ow), _ = letterbox_geometry(h0, w0, imgsz)
    out = torch.empty((oh, ow, 3), dtype=torch.uint8, device=img.device)
    tables = torch.empty(3 * (nw + nh) + 16, dtype=torch.int32, device=img.device)
    check(l.mtb_letterbox_u8(ptr(img), h0, w0, c, ptr(out), oh, ow, top, left, nh, nw, 114, int(swap_rb), ptr(tables),
                             stream_ptr()), "mtb_letterbox_u8")
    return out


def resize_aa_device(img: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """torchvision bilinear antialias resize of a device uint8 HxWx(3|4) image -> uint8 oh x ow x 3 (same channel order)."""
    l = lib()
    _declare(l)
    h0, w0, c = img.shape
    tmp = torch.empty((h0, ow, 3), dtype=torch.uint8, device=img.device)
    out = torch.empty((oh, ow, 3), dtype=torch.uint8, device=img.device)
    kx = (int(-(-max(w0 / ow, 1.0) // 1)) * 2 + 1)
    ky = (int(-(-max(h0 / oh, 1.0) // 1)) * 2 + 1)
    n_ints = 2 * ow + 2 * oh + (ow * kx + oh * ky + 3) // 2 + 64
    tables = torch.empty(n_ints, dtype=torch.int32, device=img.device)
    check(l.mtb_resize_aa_u8(ptr(img), h0, w0, c, ptr(tmp), ptr(out), oh, ow, ptr(tables), n_ints, stream_ptr()),
          "mtb_resize_aa_u8")
    return out


_LANCZOS_TABLES: "collections.OrderedDict" = collections.OrderedDict()     # (device, sh, sw, oh, ow) -> filled device scratch
_LANCZOS_TABLES_MAX = 32


def resize_lanczos_device(img: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """PIL `Image.resize((ow, oh), Image.LANCZOS)` of a device uint8 HxWx(3|4) image -> uint8 oh x ow x 3, bit-exact with
    Pillow (reference: core/image/image_utils.py:545,551-595).  The coefficient tables of the last few geometries stay on
    the device (a page geometry repeats for a whole batch; bubble crops are ragged and simply cycle through the LRU)."""
    l = lib()
    _declare(l)
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    h0, w0, c = img.shape
    if (oh, ow) == (h0, w0) and c == 3:
        return img                                   # PIL returns a copy of the same pixels
    tmp = torch.empty((h0, ow, 3), dtype=torch.uint8, device=img.device) if (ow != w0 and oh != h0) else None
    out = torch.empty((oh, ow, 3), dtype=torch.uint8, device=img.device)
    key = (str(img.device), h0, w0, oh, ow)
    ready = key in _LANCZOS_TABLES
    if ready:
        _LANCZOS_TABLES.move_to_end(key)
        tables = _LANCZOS_TABLES[key]
    else:
        n_ints = int(l.mtb_resize_lanczos_table_ints(h0, w0, oh, ow))
        tables = torch.empty(n_ints, dtype=torch.int32, device=img.device)
        _LANCZOS_TABLES[key] = tables
        while len(_LANCZOS_TABLES) > _LANCZOS_TABLES_MAX:
            _LANCZOS_TABLES.popitem(last=False)
    check(l.mtb_resize_lanczos_u8(ptr(img), h0, w0, c, ptr(tmp), ptr(out), oh, ow, ptr(tables), tables.numel(), int(ready),
                                  stream_ptr()), "mtb_resize_lanczos_u8")
    return out


def flatten_alpha_device(img: torch.Tensor, background=(255, 255, 255)) -> torch.Tensor:
    """Device uint8 HxWx4 (alpha last) -> HxWx3 over a constant background with Pillow's paste arithmetic
    (reference: convert_image_to_target_mode, core/image/image_utils.py:598-675)."""
    l = lib()
    _declare(l)
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 4 and img.is_contiguous()
    h, w, _ = img.shape
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=img.device)
    bg = (C.c_int * 3)(*[int(v) for v in background])
    check(l.mtb_flatten_alpha_u8(ptr(img), h, w, bg, ptr(out), stream_ptr()), "mtb_flatten_alpha_u8")
    return out
