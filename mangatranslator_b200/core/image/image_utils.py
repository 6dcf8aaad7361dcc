"""Upscale stage + tensor/image helpers — drop-in for the hot-path part of the reference's core/image/image_utils.py
(image_to_tensor :351, tensor_to_image :361, _upscale_image :369, upscale_image_to_dimension :377, upscale_image :503).

With the B200 upscaler (mangatranslator_b200.rcan.RcanB200) the page stays uint8 on the device: u8 -> planes, the
RCAN conv stack, the clamp/*255/truncate back to u8 and the final exact-size LANCZOS resample (:545; bit-exact with
Pillow, mtb_resize_lanczos_u8) are all kernels.  resize_to_min_side / resize_to_max_side (:551-595) and
process_bubble_image_cached (:678-746, the per-bubble crops of core/services/translation.py:2097-2258) go the same way.
calculate_centroid_expansion_box (:173-348, the renderer's per-bubble safe text box on the cleaned masks) is one
integer kernel launch for all bubbles of a page (csrc/safebox_core.cuh).
"""
from __future__ import annotations



from pathlib import Path

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200.core.caching import get_cache
from mangatranslator_b200.core.ml.model_manager import get_model_manager
from mangatranslator_b200.utils.exceptions import ImageProcessingError
from mangatranslator_b200.utils.logging import log_message
from mangatranslator_b200._lib import device_section, serialized


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


def _page_to_device(image: Image.Image, device: torch.device) -> torch.Tensor:
    if image.mode != "RGB":
        image = image.convert("RGB")                  # alpha is dropped like image_to_tensor does (:353-354)
    return torch.from_numpy(np.array(image)).to(device)


def _device_to_pil(page: torch.Tensor) -> Image.Image:
    return Image.fromarray(page.cpu().numpy())


def _met(w: int, h: int, target: int, mode: str) -> bool:
    return (max(w, h) >= target) if mode == "max" else (min(w, h) >= target)


def upscale_to_dimension_device(model, page: torch.Tensor, target: int, mode: str = "max") -> torch.Tensor:
    """2x passes on a device uint8 HxWx3 RGB page until max(w,h) (mode 'max') or min(w,h) (mode 'min') reaches `target`
    (reference loop :417-497; its PNG temp-file round trip between passes is lossless and therefore dropped).  The page
    never leaves the device; the result may alias the model's static output buffer."""
    out = page
    while not _met(out.shape[1], out.shape[0], target, mode):
        out = model.upscale_u8(out)
    return out


def resize_device(page: torch.Tensor, width: int, height: int) -> torch.Tensor:
    """PIL `image.resize((width, height), Image.LANCZOS)` on a device uint8 page (bit-exact, mtb_resize_lanczos_u8)."""
    from mangatranslator_b200.preproc import resize_lanczos_device
    return resize_lanczos_device(page.contiguous(), height, width)


@serialized
def _upscale_image(model, image: Image.Image, device: torch.device) -> Image.Image:
    """One 2x pass.  B200 models take the uint8 page directly; any other callable gets the reference's tensor contract."""
    if hasattr(model, "upscale_u8"):
        return _device_to_pil(model.upscale_u8(_page_to_device(image, model.device)))
    with torch.no_grad():
        return tensor_to_image(model(image_to_tensor(image, device)))


@serialized
def upscale_image_to_dimension(model, image: Image.Image, target: int, device: torch.device, mode: str = "max",
                               model_type: str = "model", verbose: bool = False) -> Image.Image:
    """Repeat 2x passes until max(w,h) (or min for mode='min') reaches `target` (reference :377-500)."""
    if mode not in {"max", "min"}:
        raise ImageProcessingError("mode must be 'max' or 'min'")
    if image.width <= 0 or image.height <= 0:
        raise ImageProcessingError(f"Invalid image dimensions: {image.width}x{image.height}. Cannot upscale 0x0 images.")
    cache = get_cache()
    key = cache.get_upscale_dimension_cache_key(image, target, mode, model_type)
    hit = cache.get_upscaled_image(key)
    if hit is not None:
        log_message("  - Using cached upscaled image", verbose=verbose)
        return hit
    if _met(image.width, image.height, target, mode):
        cache.set_upscaled_image(key, image)
        return image
    if hasattr(model, "upscale_u8"):
        out = _device_to_pil(upscale_to_dimension_device(model, _page_to_device(image, model.device), target, mode))
    else:
        out = image
        while not _met(out.width, out.height, target, mode):
            out = _upscale_image(model, out, device)
    cache.set_upscaled_image(key, out)
    return out


@serialized
def upscale_image(image: Image.Image, factor: float, model_type: str = "model", verbose: bool = False) -> Image.Image:
    """Reference :503-548: 2x passes until the larger side reaches the target, then an exact-size LANCZOS resample.  With
    the B200 upscaler the page is uploaded once and both steps run on the device."""
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
    if hasattr(model, "upscale_u8"):
        up = upscale_to_dimension_device(model, _page_to_device(image, model.device), max(tw, th), "max")
        result = _device_to_pil(resize_device(up, tw, th))
    else:
        up = upscale_image_to_dimension(model, image, max(tw, th), mm.device, "max", model_type, verbose)
        result = up.resize((tw, th), Image.LANCZOS)
    cache.set_upscaled_image(key, result)
    return result


@serialized
def _resize_pil(image: Image.Image, new_width: int, new_height: int) -> Image.Image:
    """`image.resize(..., Image.LANCZOS)` through the device kernel.  The hot path only resamples RGB pages and crops
    (alpha is dropped before the upscaler, :353-354); other modes are refused rather than handed to a CPU library."""
    if image.mode != "RGB":
        raise ImageProcessingError(f"resize: mode {image.mode} is not on the B200 hot path (convert to RGB first)")
    dev = get_model_manager()._require_cuda()
    return _device_to_pil(resize_device(_page_to_device(image, dev), new_width, new_height))


def resize_to_max_side(image: Image.Image, max_side: int, verbose: bool = False) -> Image.Image:
    """Largest side becomes `max_side`, aspect preserved (reference :551-566)."""
    width, height = image.size
    current_max = max(width, height)
    if current_max == max_side:
        return image
    scale = max_side / current_max
    return _resize_pil(image, max(1, int(round(width * scale))), max(1, int(round(height * scale))))


def resize_side_geometry(width: int, height: int, min_side: int):
    """(new_width, new_height) of resize_to_min_side (reference :583-587), or None when nothing changes."""
    current_min = min(width, height)
    if current_min == min_side:
        return None
    scale = min_side / current_min
    return max(1, int(round(width * scale))), max(1, int(round(height * scale)))


def resize_to_min_side(image: Image.Image, min_side: int, verbose: bool = False) -> Image.Image:
    """Smallest side becomes `min_side`, aspect preserved (reference :569-595)."""
    width, height = image.size
    if width <= 0 or height <= 0:
        raise ImageProcessingError(f"Invalid image dimensions: {width}x{height}. Cannot resize 0x0 images.")
    g = resize_side_geometry(width, height, min_side)
    return image if g is None else _resize_pil(image, *g)


@serialized
def process_bubble_image_cached(bubble_image_pil: Image.Image, upscale_model, device: torch.device,
                                target_min_side: int = 200, mode: str = "min", model_type: str = "model",
                                verbose: bool = False) -> Image.Image:
    """Reference :678-746 — the per-bubble crop the translation service sends to the LLM: 2x passes until the crop's
    `mode` side reaches `target_min_side`, then LANCZOS so that the smaller side is exactly `target_min_side`."""
    cache = get_cache()
    key = cache.get_bubble_processing_cache_key(bubble_image_pil, target_min_side, mode, model_type)
    hit = cache.get_upscaled_image(key)
    if hit is not None:
        log_message("  - Using cached bubble processing result", verbose=verbose)
        return hit
    if hasattr(upscale_model, "upscale_u8"):
        if mode not in {"max", "min"}:
            raise ImageProcessingError("mode must be 'max' or 'min'")
        if bubble_image_pil.width <= 0 or bubble_image_pil.height <= 0:
            raise ImageProcessingError(f"Invalid image dimensions: {bubble_image_pil.width}x{bubble_image_pil.height}. "
                                       "Cannot upscale 0x0 images.")
        out = _device_to_pil(process_bubble_crop_device(_page_to_device(bubble_image_pil, upscale_model.device),
                                                        upscale_model, target_min_side, mode))
    else:
        up = upscale_image_to_dimension(upscale_model, bubble_image_pil, target_min_side, device, mode, model_type, verbose)
        out = resize_to_min_side(up, target_min_side, verbose)
    cache.set_upscaled_image(key, out)
    return out


def process_bubble_crop_device(crop: torch.Tensor, upscale_model, target_min_side: int = 200,
                               mode: str = "min") -> torch.Tensor:
    """Device part of process_bubble_image_cached: crop (uint8 hxwx3 RGB on the device) -> upscaled + resized crop."""
    up = upscale_to_dimension_device(upscale_model, crop.contiguous(), target_min_side, mode)
    g = resize_side_geometry(up.shape[1], up.shape[0], target_min_side)
    return up.clone() if g is None else resize_device(up, *g)


@serialized
def process_page_bubbles_device(page_rgb: torch.Tensor, bboxes, upscale_model, target_min_side: int = 200,
                                mode: str = "min") -> list:
    """All bubble crops of one device-resident page (the loop of core/services/translation.py:2097-2258 around
    process_bubble_image_cached): crop -> RCAN passes -> exact LANCZOS, nothing leaves the device."""
    h, w = page_rgb.shape[:2]
    out = []
    for (x0, y0, x1, y1) in bboxes:
        x0, y0, x1, y1 = max(0, int(x0)), max(0, int(y0)), min(w, int(x1)), min(h, int(y1))
        if x1 <= x0 or y1 <= y0:
            raise ImageProcessingError(f"Invalid image dimensions: {x1 - x0}x{y1 - y0}. Cannot upscale 0x0 images.")
        out.append(process_bubble_crop_device(page_rgb[y0:y1, x0:x1, :3].contiguous(), upscale_model, target_min_side,
                                              mode))
    return out


def convert_image_to_target_mode(image: Image.Image, target_mode: str, verbose: bool = False) -> Image.Image:
    """Reference :598-675.  RGBA / LA / P-with-transparency pages headed for RGB are flattened onto white with Pillow's
    paste arithmetic — on the device for RGBA pages (mtb_flatten_alpha_u8); mode bookkeeping (LA / P expansion, plain
    conversions) stays with Pillow."""
    if image.mode == target_mode:
        return image
    if target_mode == "RGB" and (image.mode in ("RGBA", "LA") or (image.mode == "P" and "transparency" in image.info)):
        log_message(f"Converting {image.mode} to RGB (flattening transparency)", verbose=verbose)
        rgba = image if image.mode == "RGBA" else image.convert("RGBA")
        with device_section:
            dev = get_model_manager()._require_cuda()
            from mangatranslator_b200.preproc import flatten_alpha_device
            return _device_to_pil(flatten_alpha_device(torch.from_numpy(np.array(rgba)).to(dev)))
    return image.convert(target_mode)


def save_image_with_compression(image: Image.Image, output_path, jpeg_quality: int = 95, png_compression: int = 2,
                                verbose: bool = False):
    """What reaches the file in the reference's function of the same name (image_utils.py:59-170): JPEG on a white
    background with the clamped quality, lossless PNG / WEBP, unknown extensions become .png.  The reference re-packs the
    PNG stream with oxipng; that is lossless, so the decoded pixels are identical while the file bytes differ (image
    encoding is outside this build).  Returns the path written."""
    path = Path(output_path)
    ext = path.suffix.lower()
    try:
        path.parent.mkdir(parents=True, exist_ok=True)
        if ext in (".jpg", ".jpeg"):
            if image.mode in ("RGBA", "LA"):
                flat = Image.new("RGB", image.size, (255, 255, 255))
                flat.paste(image, mask=image.split()[-1])
                image = flat
            elif image.mode != "RGB":
                image = image.convert("RGB")
            image.save(path, format="JPEG", quality=max(1, min(int(jpeg_quality), 100)))
        elif ext == ".webp":
            image.save(path, format="WEBP", lossless=True)
        else:
            if ext != ".png":
                log_message(f"Warning: Unknown output extension '{ext}'. Saving as PNG.", verbose=verbose, always_print=True)
                path = path.with_suffix(".png")
            image.save(path, format="PNG", compress_level=min(6, max(0, int(png_compression))))
    except Exception as e:
        raise ImageProcessingError(f"Failed to save image to {output_path}") from e
    return path


def safe_boxes_for_masks(masks, padding_pixels: float = 4.0, bboxes=None):
    """All bubbles of a page in one launch: `masks` are device uint8 HxW tensors (e.g. the cleaned masks of
    clean_pages_device).  Returns a list with, per mask, ((x, y, w, h), (cx, cy)) or the ImageProcessingError the
    reference would have raised for it (returned, not raised, so one bad bubble does not hide the others)."""
    from mangatranslator_b200 import safebox_host as S
    with device_section:
        recs = S.safe_boxes_device(masks, padding_pixels, bboxes)
    out = []
    for r in recs:
        try:
            out.append(S.decode(r))
        except ValueError as e:
            out.append(ImageProcessingError(e.args[0]))
    return out


def calculate_centroid_expansion_box(cleaned_mask, padding_pixels: float = 4.0, verbose: bool = False):
    """Reference signature (:173-175): mask (numpy uint8 HxW, or a device tensor) -> ((x, y, width, height), (cx, cy));
    raises ImageProcessingError with the reference's messages (:204-205, :348)."""
    if cleaned_mask is None:
        raise ImageProcessingError("Invalid or empty mask provided")
    dev = get_model_manager()._require_cuda()
    if isinstance(cleaned_mask, torch.Tensor):
        m = cleaned_mask.to(dev)
    else:
        a = np.asarray(cleaned_mask)
        if a.ndim != 2:
            raise ImageProcessingError("Safe area calculation failed")
        if a.dtype != np.uint8:
            a = a.astype(np.uint8)                      # what assigning into the reference's uint8 frame does (:213)
        m = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if m.numel() == 0:
        raise ImageProcessingError("Invalid or empty mask provided")
    res = safe_boxes_for_masks([m.contiguous()], padding_pixels)[0]
    if isinstance(res, Exception):
        if res.args[0] != "Invalid or empty mask provided":
            log_message(f"Safe area calculation failed: {res}", verbose=verbose, always_print=True)
        raise res
    box, centroid = res
    log_message(f"Safe area: {box[2]:.0f}x{box[3]:.0f} at ({centroid[0]:.0f}, {centroid[1]:.0f})", verbose=verbose)
    return box, centroid
