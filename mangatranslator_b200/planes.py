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


# ---- "fp16c" format of the RCAN body (csrc/conv_halo_fp16c.cu): fp16 value + e5m2 rounding residual -----------------
# activations: uint8 [3][N][H][W][64]: plane 0 = fp16 of channels 0-31, plane 1 = fp16 of channels 32-63 (64 B per pixel
# each), plane 2 = e5m2 of (v - fp16(v)) * 2^lo_shift for the 64 channels.
FP16C_ROWS = 9 * 2 * 128 + 9 * 64


def nhwc_to_fp16c(v: torch.Tensor, lo_shift: int = 0) -> torch.Tensor:
    """float [N][H][W][64] -> uint8 [3][N][H][W][64]"""
    assert v.shape[-1] == 64
    v = v.float()
    h = v.to(torch.float16)
    lo = ((v - h.float()) * (2.0 ** lo_shift)).to(torch.float8_e5m2)
    out = torch.empty((3,) + tuple(v.shape[:-1]) + (64,), dtype=torch.uint8, device=v.device)
    out[0] = h[..., :32].contiguous().view(torch.uint8).reshape(*v.shape[:-1], 64)
    out[1] = h[..., 32:].contiguous().view(torch.uint8).reshape(*v.shape[:-1], 64)
    out[2] = lo.view(torch.uint8)
    return out


def fp16c_to_nhwc(p: torch.Tensor, lo_shift: int = 0) -> torch.Tensor:
    """uint8 [3][N][H][W][64] -> float [N][H][W][64]"""
    h = torch.cat([p[0].contiguous().view(torch.float16), p[1].contiguous().view(torch.float16)], dim=-1).float()
    lo = p[2].contiguous().view(torch.float8_e5m2).float() * (2.0 ** -lo_shift)
    return h + lo


def nchw_to_fp16c(x: torch.Tensor, lo_shift: int = 0) -> torch.Tensor:
    return nhwc_to_fp16c(x.permute(0, 2, 3, 1).contiguous(), lo_shift)


def fp16c_to_nchw(p: torch.Tensor, lo_shift: int = 0) -> torch.Tensor:
    return fp16c_to_nhwc(p, lo_shift).permute(0, 3, 1, 2).contiguous()


def fp16c_row_channels() -> torch.Tensor:
    """Output channel held by accumulator row i of a 16-row group: the epilogue reads TMEM with tcgen05.ld.16x256b, which
    hands a thread rows r and r + 8, so those carry the adjacent channels 2r and 2r + 1 (one 4-byte fp16 pair store)."""
    i = torch.arange(16)
    return torch.where(i < 8, 2 * i, 2 * (i - 8) + 1)


def conv_weight_to_fp16c(w: torch.Tensor, lo_shift: int = 0) -> torch.Tensor:
    """torch conv weight [64][64][3][3] -> uint8 [2880][64], the kernel's shared-memory order:
    per tap (ky*3+kx) and input-channel half (0-31, 32-63) 128 rows of 32 halves — rows 32g..32g+15 = fp16(w) of the
    output channels 16g + fp16c_row_channels(), rows 32g+16..32g+31 = fp16(w - fp16(w)) of the same channels (a channel's
    two partial sums then sit in one warp's TMEM lane quarter) —, then per tap 64 rows in the same channel order (an M = 64
    instruction writes row i to TMEM lane 32*(i/16) + i%16, where the fp16 hi row accumulates) of 64 e5m2 values
    e5m2(w * 2^-lo_shift)."""
    co, ci, kh, kw = w.shape
    assert (co, ci, kh, kw) == (64, 64, 3, 3), w.shape
    wt = w.float().permute(2, 3, 0, 1).reshape(9, 64, 64)                 # [tap][cout][cin]
    rows = (16 * torch.arange(4).view(4, 1) + fp16c_row_channels().view(1, 16)).reshape(64).to(w.device)
    wt = wt[:, rows, :]                                                   # accumulator row order (see fp16c_row_channels)
    hi = wt.to(torch.float16)
    lo = (wt - hi.float()).to(torch.float16)
    # [tap][half][g][plane][i][32 cin]
    hi6 = hi.reshape(9, 4, 16, 2, 32).permute(0, 3, 1, 2, 4)             # [tap][half][g][i][32]
    lo6 = lo.reshape(9, 4, 16, 2, 32).permute(0, 3, 1, 2, 4)
    w16 = torch.stack([hi6, lo6], dim=3).contiguous()                    # [tap][half][g][plane][i][32]
    w16 = w16.view(torch.uint8).reshape(9 * 2 * 128, 64)
    w8 = (wt * (2.0 ** -lo_shift)).to(torch.float8_e5m2).view(torch.uint8).reshape(9 * 64, 64)
    return torch.cat([w16, w8], dim=0).contiguous()
