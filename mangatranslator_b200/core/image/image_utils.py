"""Upscale stage + tensor/image helpers — drop-in for the hot-path part of the reference's core/image/image_utils.py
(image_to_tensor :351, tensor_to_image :361, _upscale_image :369, upscale_image_to_dimension :377, upscale_image :503).

With the B200 upscaler (mangatranslator_b200.rcan.RcanB200) the page stays uint8 on the device: u8 -> planes, the
RCAN conv stack and the clamp/*255/truncate back to u8 are all kernels; only the final exact-size LANCZOS resample
(a no-op at exactly 2x, where the reference still calls it) is PIL on the host, as in the reference (:545).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200.core.caching import get_cache
from mangatranslator_b200.core.ml.model_manager import get_model_manager
from mangatranslator_b200.utils.logging import log_message


def pil_to_cv2(pil_image: Image.Image) -> np.ndarray:
    if pil_image.mode == "RGBA":
        return np.ascontiguousarray(np.asarray(pil_image)[:, :, [2, 1, 0, 3]])
    if pil_image.mode != "RGB":
        pil_image = pil_image.convert("RGB")
    return np.ascontiguousarray(np.asarray(pil_image)[:, :, ::-1])


def cv2_to_pil(cv2_image: np.ndarray) -> Image.Image:
    if cv2_image.ndim == 3 and cv2_image.shape[2] == 4:
        return Image.fromarray(np.ascontiguousarray(cv2_image[:, :, [2, 1, 0, 3]]), "RGBA")
    return Image.fromarray(np.ascontiguousarray(cv2_image[:, :, ::-1]), "RGB")


def image_to_tensor(image: Image.Image, device: torch.device) -> torch.Tensor:
    if image.mode != "RGB":
        image = image.convert("RGB")
    a = np.array(image).astype(np.float32) / 255.0
    return torch.from_numpy(a).permute(2, 0, 1).unsqueeze(0).to(device)


def tensor_to_image(tensor: torch.Tensor) -> Image.Image:
    a = (tensor.squeeze(0).permute(1, 2, 0).clamp(0, 1).cpu().numpy() * 255).astype(np.uint8)   # truncation, not rounding
    return Image.fromarray(a)


def _upscale_image(model, image: Image.Image, device: torch.device) -> Image.Image:
    """One 2x pass.  B200 models take the uint8 page directly; any other callable gets the reference's tensor contract."""
    if hasattr(model, "upscale_u8"):
        if image.mode != "RGB":
            image = image.convert("RGB")
        dev = model.device
        page = torch.from_numpy(np.ascontiguousarray(np.asarray(image))).to(dev)
        return Image.fromarray(model.upscale_u8(page).cpu().numpy())
    with torch.no_grad():
        return tensor_to_image(model(image_to_tensor(image, device)))


def upscale_image_to_dimension(model, image: Image.Image, target: int, device: torch.device, mode: str = "max",
                               model_type: str = "model", verbose: bool = False) -> Image.Image:
    """Repeat 2x passes until max(w,h) (or min for mode='min') reaches `target` (reference :377-500; the reference's
    PNG temp-file round trip between passes is lossless and therefore dropped)."""
    pick = max if mode == "max" else min
    out = image
    guard = 0
    while pick(out.width, out.height) < target and guard < 6:
        out = _upscale_image(model, out, device)
        guard += 1
    return out


def upscale_image(image: Image.Image, factor: float, model_type: str = "model", verbose: bool = False) -> Image.Image:
    if factor == 1.0:
        return image
    cache = get_cache()
    key = cache.get_upscale_cache_key(image, factor, model_type)
    hit = cache.get_upscaled_image(key)
    if hit is not None:
        log_message("  - Using cached upscaled image", verbose=verbose)
        return hit
    mm = get_model_manager()
    model = mm.load_upscale_lite() if model_type == "model_lite" else mm.load_upscale()
    log_message(f"Upscaling image by {factor}x{' with lite model' if model_type == 'model_lite' else ''}...", verbose=verbose)
    tw, th = int(image.width * factor), int(image.height * factor)
    up = upscale_image_to_dimension(model, image, max(tw, th), mm.device, "max", model_type, verbose)
    result = up.resize((tw, th), Image.LANCZOS)
    cache.set_upscaled_image(key, result)
    return result


def convert_image_to_target_mode(image: Image.Image, target_mode: str, verbose: bool = False) -> Image.Image:
    if image.mode == target_mode:
        return image
    if target_mode == "RGB" and image.mode in ("RGBA", "LA", "P"):
        rgba = image.convert("RGBA")
        bg = Image.new("RGB", rgba.size, (255, 255, 255))
        bg.paste(rgba, mask=rgba.split()[3])
        return bg
    return image.convert(target_mode)
