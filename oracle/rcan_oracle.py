"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle of the 2x-AnimeSharpV4 upscaler's architecture (RCAN).

The reference runs it through a third-party package that is absent here: `spandrel>=0.3.0` (requirements.txt:17),
call sites core/ml/model_manager.py:652-654 (ModelLoader().load_from_state_dict) and
core/image/image_utils.py:369-374 (`model(x)` on a 1x3xHxW float32 [0,1] tensor).  This file restates the published
RCAN architecture (Zhang et al., ECCV 2018; spandrel `architectures/RCAN`): head conv -> G residual groups of R
residual channel-attention blocks (conv3x3-ReLU-conv3x3-CALayer, +skip) each closed by a conv (+skip) -> conv (+global
skip) -> Upsampler(conv3x3 -> PixelShuffle(2)) -> conv3x3, with optional DIV2K mean shift.  State-dict keys follow the
original RCAN layout (head.0 / body.G.body.R.body.{0,2,3.conv_du.{0,2}} / tail.0.0 / tail.1) so a real checkpoint maps
1:1.  PARITY UNPINNED: neither spandrel nor the checkpoint is available offline, so this oracle is checked only for
self-consistency (shape/dtype contract of image_utils.py, key-driven hyper-parameter inference), not against spandrel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class CALayer(nn.Module):
    def __init__(self, f: int, reduction: int):
        super().__init__()
        self.conv_du = nn.Sequential(nn.Conv2d(f, f // reduction, 1), nn.ReLU(inplace=True),
                                     nn.Conv2d(f // reduction, f, 1), nn.Sigmoid())

    def forward(self, x):
        return x * self.conv_du(x.mean((2, 3), keepdim=True))


class RCAB(nn.Module):
    def __init__(self, f: int, reduction: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(f, f, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(f, f, 3, padding=1),
                                  CALayer(f, reduction))

    def forward(self, x):
        return x + self.body(x)


class ResidualGroup(nn.Module):
    def __init__(self, f: int, reduction: int, n_blocks: int):
        super().__init__()
        self.body = nn.Sequential(*[RCAB(f, reduction) for _ in range(n_blocks)], nn.Conv2d(f, f, 3, padding=1))

    def forward(self, x):
        return x + self.body(x)


class RCAN(nn.Module):
    DIV2K_MEAN = (0.4488, 0.4371, 0.4040)

    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, scale=2, rgb_range=1.0, norm=False):
        super().__init__()
        assert scale == 2, "the hot path only needs the 2x model"
        self.cfg = dict(n_resgroups=n_resgroups, n_resblocks=n_resblocks, n_feats=n_feats, reduction=reduction,
                        scale=scale, rgb_range=rgb_range, norm=norm)
        self.rgb_range, self.norm = rgb_range, norm
        self.head = nn.Sequential(nn.Conv2d(3, n_feats, 3, padding=1))
        self.body = nn.Sequential(*[ResidualGroup(n_feats, reduction, n_resblocks) for _ in range(n_resgroups)],
                                  nn.Conv2d(n_feats, n_feats, 3, padding=1))
        self.tail = nn.Sequential(nn.Sequential(nn.Conv2d(n_feats, 4 * n_feats, 3, padding=1), nn.PixelShuffle(2)),
                                  nn.Conv2d(n_feats, 3, 3, padding=1))

    def forward(self, x):
        mean = torch.tensor(self.DIV2K_MEAN, dtype=x.dtype).view(1, 3, 1, 1) * self.rgb_range
        x = x * self.rgb_range
        if self.norm:
            x = x - mean
        h = self.head(x)
        y = self.body(h) + h
        y = self.tail(y)
        if self.norm:
            y = y + mean
        return y / self.rgb_range


def infer_config(state_dict) -> dict:
    """Hyper-parameters from the key/shape structure, like spandrel's loader does."""
    groups = {int(k.split(".")[1]) for k in state_dict if k.startswith("body.") and k.count(".") >= 4}
    g = max(groups) + 1
    blocks = {int(k.split(".")[3]) for k in state_dict if k.startswith("body.0.body.") and ".body." in k[12:]}
    r = max(blocks) + 1
    f = state_dict["head.0.weight"].shape[0]
    red = f // state_dict["body.0.body.0.body.3.conv_du.0.weight"].shape[0]
    return dict(n_resgroups=g, n_resblocks=r, n_feats=f, reduction=red, scale=2)


def make_model(seed: int = 0, **cfg) -> RCAN:
    """Seeded random weights (no checkpoint offline)."""
    torch.manual_seed(seed)
    m = RCAN(**cfg).eval()
    return m


def upscale_u8(model: RCAN, rgb_u8):
    """image_to_tensor -> model -> tensor_to_image as the reference does (core/image/image_utils.py:351-374):
    /255 float32 NCHW, clamp [0,1], *255, truncate to uint8.  Returns (float output NCHW, uint8 HxWx3)."""
    import numpy as np
    x = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1).float().div(255.0).unsqueeze(0)
    with torch.no_grad():
        y = model(x)
    out = (y.squeeze(0).clamp(0, 1).permute(1, 2, 0).numpy() * 255.0).astype(np.uint8)
    return y, out
