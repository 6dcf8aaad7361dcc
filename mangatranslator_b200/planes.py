"""bf16 hi/lo "plane" tensors: the on-device activation/weight format of the DNN kernels.

A float tensor v is carried as planes[0] = bf16(v), planes[1] = bf16(v - planes[0]); the tensor-core
contraction Ahi*Bhi + Ahi*Blo + Alo*Bhi then reproduces the fp32 product to ~2^-16 relative error, which is what
keeps float outputs within the 1e-3 parity bound against the reference's fp32 CPU path.
Layout is NHWC with channels zero-padded to a multiple of 64 (one 128-byte swizzle row per 64 channels).
"""
from __future__ import annotations

import torch


def pad_to(c: int, m: int) -> int:
    return (c + m - 1) // m * m


def split_planes(v: torch.Tensor, planes: int = 2) -> torch.Tensor:
    """float tensor [...] -> bf16 [planes, ...]"""
    v = v.float()
    hi = v.to(torch.bfloat16)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    lo = (v - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def merge_planes(p: torch.Tensor) -> torch.Tensor:
    """bf16 [planes, ...] -> float [...]"""
    out = p[0].float()
    for i in range(1, p.shape[0]):
        out = out + p[i].float()
    return out


def nchw_to_planes(x: torch.Tensor, planes: int = 2, cpad: int = 64) -> torch.Tensor:
    """float NCHW -> bf16 [planes][N][H][W][Cpad]"""
    n, c, h, w = x.shape
    cp = pad_to(c, cpad)
    y = torch.zeros((n, h, w, cp), dtype=torch.float32, device=x.device)
    y[..., :c] = x.permute(0, 2, 3, 1)
    return split_planes(y, planes)


def planes_to_nchw(p: torch.Tensor, c: int) -> torch.Tensor:
    """bf16 [planes][N][H][W][Cpad] -> float NCHW (first c channels)"""
    return merge_planes(p)[..., :c].permute(0, 3, 1, 2).contiguous()


def conv_weight_to_planes(w: torch.Tensor, planes: int = 2, cin_pad: int = 64, cout_pad: int = 16) -> torch.Tensor:
    """torch conv weight [Cout][Cin][KH][KW] -> bf16 [planes][KH*KW][CoutP][CinP] (K-major rows of Cin)."""
    co, ci, kh, kw = w.shape
    cop, cip = pad_to(co, cout_pad), pad_to(ci, cin_pad)
    y = torch.zeros((kh * kw, cop, cip), dtype=torch.float32, device=w.device)
    y[:, :co, :ci] = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    return split_planes(y, planes)


def pad_bias(b: torch.Tensor | None, cout: int, cout_pad: int = 16) -> torch.Tensor | None:
    if b is None:
        return None
    cop = pad_to(cout, cout_pad)
    y = torch.zeros(cop, dtype=torch.float32, device=b.device)
    y[:cout] = b.float()
    return y
