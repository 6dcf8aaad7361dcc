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

    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, scale=2, rgb_range=1.0, norm=False,
                 unshuffle=1):
        """`unshuffle` = d > 1 restates the "_PU" variant (spandrel RCAN `unshuffle_mod`, the reference's lite model
        2x-AnimeSharpV4_Fast_RCAN_PU, core/ml/model_manager.py:660-700): the input is reflect-padded to a multiple of d
        and pixel-unshuffled (3*d*d channels), the upsampler scales by scale*d in PixelShuffle(2) stages and the output
        is cropped to scale*H x scale*W.  UNVERIFIED against spandrel (absent): pad mode and stage layout from memory."""
        super().__init__()
        assert scale == 2, "the hot path only needs the 2x models"
        self.cfg = dict(n_resgroups=n_resgroups, n_resblocks=n_resblocks, n_feats=n_feats, reduction=reduction,
                        scale=scale, rgb_range=rgb_range, norm=norm, unshuffle=unshuffle)
        self.rgb_range, self.norm, self.unshuffle, self.scale = rgb_range, norm, unshuffle, scale
        self.head = nn.Sequential(nn.Conv2d(3 * unshuffle * unshuffle, n_feats, 3, padding=1))
        self.body = nn.Sequential(*[ResidualGroup(n_feats, reduction, n_resblocks) for _ in range(n_resgroups)],
                                  nn.Conv2d(n_feats, n_feats, 3, padding=1))
        ups = []
        total = scale * unshuffle
        while total > 1:
            assert total % 2 == 0
            ups += [nn.Conv2d(n_feats, 4 * n_feats, 3, padding=1), nn.PixelShuffle(2)]     # tail.0.0, tail.0.2, ...
            total //= 2
        self.tail = nn.Sequential(nn.Sequential(*ups), nn.Conv2d(n_feats, 3, 3, padding=1))

    def forward(self, x):
        mean = torch.tensor(self.DIV2K_MEAN, dtype=x.dtype).view(1, 3, 1, 1) * self.rgb_range
        x = x * self.rgb_range
        if self.norm:
            x = x - mean
        hh, ww = x.shape[2], x.shape[3]
        d = self.unshuffle
        if d > 1:
            x = F.pad(x, (0, -ww % d, 0, -hh % d), mode="reflect")
            x = F.pixel_unshuffle(x, d)
        h = self.head(x)
        y = self.body(h) + h
        y = self.tail(y)[:, :, :self.scale * hh, :self.scale * ww]
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
    d = int(round((state_dict["head.0.weight"].shape[1] / 3) ** 0.5))
    stages = len([k for k in state_dict if k.startswith("tail.0.") and k.endswith(".weight")])
    return dict(n_resgroups=g, n_resblocks=r, n_feats=f, reduction=red, scale=(2 ** stages) // d, unshuffle=d)


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


# ---- the reference's wrappers around the model (core/image/image_utils.py) ------------------------------------------
def _met(w: int, h: int, target: int, mode: str) -> bool:
    return (max(w, h) >= target) if mode == "max" else (min(w, h) >= target)


def upscale_to_dimension(model: RCAN, rgb_u8, target: int, mode: str):
    """upscale_image_to_dimension (:377-500): 2x passes until the max / min side reaches `target`."""
    cur = rgb_u8
    while not _met(cur.shape[1], cur.shape[0], target, mode):
        cur = upscale_u8(model, cur)[1]
    return cur


def upscale_image(model: RCAN, rgb_u8, factor: float):
    """upscale_image (:503-548): passes until the larger side reaches int(side*factor), then PIL LANCZOS to the exact size."""
    from PIL import Image
    if factor == 1.0:
        return rgb_u8
    h, w = rgb_u8.shape[:2]
    tw, th = int(w * factor), int(h * factor)
    up = upscale_to_dimension(model, rgb_u8, max(tw, th), "max")
    import numpy as np
    return np.asarray(Image.fromarray(up).resize((tw, th), Image.LANCZOS))


def process_bubble(model: RCAN, rgb_u8, target_min_side: int = 200, mode: str = "min"):
    """process_bubble_image_cached (:678-746): passes until the `mode` side reaches the target, then resize_to_min_side
    (:569-595: LANCZOS so that the smaller side equals the target)."""
    import numpy as np
    from PIL import Image
    up = upscale_to_dimension(model, rgb_u8, target_min_side, mode)
    h, w = up.shape[:2]
    cur = min(w, h)
    if cur == target_min_side:
        return up
    s = target_min_side / cur
    nw, nh = max(1, int(round(w * s))), max(1, int(round(h * s)))
    return np.asarray(Image.fromarray(up).resize((nw, nh), Image.LANCZOS))
