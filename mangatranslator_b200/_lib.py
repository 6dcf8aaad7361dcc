"""ctypes binding of libmtb200.so (the C ABI declared in include/mtb200.h).

The product path has no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MTB200_LIB", _HERE / "lib" / "libmtb200.so"))


class MtbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad",
        "planes_in", "planes_out", "act", "res_planes", "tile_w", "tile_h", "x_ctotal", "x_coff", "out_ctotal", "out_coff", "res_ctotal", "res_coff",
        "res_bcast", "act_after_res", "pixel_shuffle", "mode")]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MtbError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    vp, i32, f32p = C.c_void_p, C.c_int, C.c_void_p
    l.mtb_last_error.restype = C.c_char_p
    l.mtb_version.restype = C.c_int
    l.mtb_launch_count.restype = C.c_longlong
    l.mtb_conv_plan_create.argtypes = [C.POINTER(ConvDesc), vp, vp, f32p, vp, vp, f32p, C.POINTER(vp)]
    l.mtb_conv_plan_create.restype = i32
    l.mtb_conv_plan_run.argtypes = [vp, vp]
    l.mtb_conv_plan_run.restype = i32
    l.mtb_conv_plan_num_mtiles.argtypes = [vp]
    l.mtb_conv_plan_num_mtiles.restype = i32
    l.mtb_conv_plan_num_sum_rows.argtypes = [vp]
    l.mtb_conv_plan_num_sum_rows.restype = i32
    l.mtb_conv_plan_set_channel_scale.argtypes = [vp, vp]
    l.mtb_conv_plan_set_channel_scale.restype = i32
    l.mtb_conv_plan_set_border_sums.argtypes = [vp, vp]
    l.mtb_conv_plan_set_border_sums.restype = i32
    l.mtb_conv_plan_destroy.argtypes = [vp]
    l.mtb_conv_plan_destroy.restype = None


_exp = None


def exp_lib() -> C.CDLL:
    """libmtb200_exp.so: hardware-behaviour probes (csrc/experiments.cu).  Tools only; nothing in the product imports it."""
    global _exp
    if _exp is None:
        lib()   # the probes link against the product library (TMA-descriptor and error helpers)
        path = LIB_PATH.parent / "libmtb200_exp.so"
        if not path.exists():
            raise MtbError(f"{path} not found: run `make -C mangatranslator_b200/csrc`")
        _exp = C.CDLL(str(path))
        vp, i32 = C.c_void_p, C.c_int
        _exp.mtb_exp_shifted_desc.argtypes = [vp, vp, vp, i32, i32, i32, vp]
        _exp.mtb_exp_shifted_desc.restype = i32
    return _exp


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mtb_last_error().decode("utf-8", "replace")
        raise MtbError(f"{what} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().mtb_launch_count())


# ---- one device section at a time --------------------------------------------------------------------------------
# The reference's batch mode runs several page threads against the same model objects (core/pipeline.py:2470); the
# B200 models keep static activation buffers and CUDA graphs per input size, so their stage functions serialise on
# this re-entrant lock (the GPU executes one page at a time anyway; host-side work of other threads still overlaps).
import functools
import threading

device_section = threading.RLock()


def serialized(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with device_section:
            return fn(*args, **kwargs)
    return wrapper
