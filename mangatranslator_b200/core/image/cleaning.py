"""Speech-bubble cleaning stage — drop-in for the reference's core/image/cleaning.py.

Same public functions and result dictionaries (`clean_speech_bubbles` :524-553, `retry_cleaning_with_otsu` :1051,
`process_single_bubble` :210), but every pixel operation runs in the sm_100a kernels behind
mtb_clean_bubbles / mtb_clean_paint (include/mtb200.h); this module only converts arguments, plans crop windows and
rebuilds the reference's return values.  Not supported (out of scope): the Flux inpainting branch
for "coloured" bubbles (`inpaint_colored_bubbles=True`), which the reference only takes with a diffusion backend.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200 import clean_host as H
from mangatranslator_b200.clean_engine import STATUS_NAMES, clean_batch
from mangatranslator_b200.utils.exceptions import CleaningError, ImageProcessingError, ValidationError
from mangatranslator_b200.utils.logging import log_message
from mangatranslator_b200._lib import serialized

GRAYSCALE_MIDPOINT = 128
MIN_CONTOUR_AREA = H.MIN_CONTOUR_AREA
DILATION_KERNEL_SIZE = H.DILATION_KERNEL_SIZE
EROSION_KERNEL_SIZE = H.EROSION_KERNEL_SIZE
DISTANCE_TRANSFORM_MASK_SIZE = 5


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise CleaningError("CUDA device required: the B200 cleaning path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _pil_to_bgr(pil_image: Image.Image) -> np.ndarray:
    """PIL -> BGR / BGRA uint8 array, like the reference's pil_to_cv2 (core/image/image_utils.py:20-56)."""
    if pil_image.mode == "RGBA":
        a = np.asarray(pil_image)
        return np.ascontiguousarray(a[:, :, [2, 1, 0, 3]])
    if pil_image.mode != "RGB":
        pil_image = pil_image.convert("RGB")
    return np.ascontiguousarray(np.asarray(pil_image)[:, :, ::-1])


def _masks_from_detections(detections: List[Dict[str, Any]], h: int, w: int):
    """sam_mask entries are used as they are; polygon (`mask_points`) detections are rasterised on the host exactly
    like the reference does with cv2.fillPoly (cleaning.py:743-760) — a few hundred vertices, not a pixel pass."""
    out = []
    for det in detections:
        d = dict(det)
        if d.get("sam_mask") is None and d.get("mask_points"):
            import cv2  # host-side polygon rasterisation only
            pts = np.array(d["mask_points"], dtype=np.float32)
            if pts.ndim == 3 and pts.shape[1] == 1:
                pts_i = np.round(pts).astype(int)
            elif pts.ndim == 2 and pts.shape[1] == 2:
                pts_i = np.round(pts).astype(int).reshape((-1, 1, 2))
            else:
                d["sam_mask"] = None
                out.append((d, False))
                continue
            m = np.zeros((h, w), np.uint8)
            cv2.fillPoly(m, [pts_i], 255)
            d["sam_mask"] = m
            out.append((d, False))
        else:
            out.append((d, d.get("sam_mask") is not None))
    return out


def _result_to_dict(res: H.CleanResult, det: Dict[str, Any], base_mask, mask, is_sam: bool, n_channels: int):
    tc = None
    if res.has_text_color:
        if res.text_color[3] == -1 or n_channels == 3:
            tc = (int(res.text_color[0]), int(res.text_color[1]), int(res.text_color[2]))
        else:
            tc = tuple(int(res.text_color[i]) for i in range(n_channels))
    color = (int(res.fill_bgr[0]), int(res.fill_bgr[1]), int(res.fill_bgr[2]))
    return {
        "mask": mask,
        "base_mask": base_mask,
        "color": color,
        "bbox": det.get("bbox"),
        "is_colored": False,
        "text_bbox": tuple(int(v) for v in res.text_bbox),
        "text_color_bgr": tc,
        "is_sam": is_sam,
        "inpainted": False,
    }


@serialized
def clean_pages_device(pages: List[torch.Tensor], detections: List[List[Dict[str, Any]]], *,
                       thresholding_value: int = 200, use_otsu_threshold: bool = False, roi_shrink_px: float = 5,
                       processing_scale: float = 1.0, in_place: bool = False):
    """Batch entry used by the pipeline / bench: device pages in, device pages + per-bubble results out
    (no full-frame masks cross PCIe)."""
    params = H.build_params(thresholding_value, use_otsu_threshold, roi_shrink_px, processing_scale, retry_otsu=True)
    return clean_batch(pages, detections, params, in_place=in_place)


@serialized
def clean_speech_bubbles(
    image_input: Union[str, Path, Image.Image],
    model_path=None,
    confidence=0.6,
    pre_computed_detections=None,
    device=None,
    thresholding_value: int = 200,
    use_otsu_threshold: bool = False,
    roi_shrink_px: int = 5,
    verbose: bool = False,
    processing_scale: float = 1.0,
    conjoined_confidence=0.35,
    inpaint_colored_bubbles: bool = False,
    flux_hf_token: str = "",
    flux_num_inference_steps: int = 8,
    flux_residual_diff_threshold: float = 0.15,
    flux_seed: int = 1,
    osb_text_verification: bool = False,
    osb_text_hf_token: str = "",
    inpaint_method: str = "flux_kontext",
    flux_backend: str = "sdnq",
    flux_low_vram: bool = False,
    flux_sdcpp_cache_mode: str = "none",
    flux_sdcpp_diffusion_quant: str = "Q4_K_M",
    flux_sdcpp_text_encoder_quant: str = "",
    flux_luminance_correction: bool = True,
    flux_upscale_small_crops: bool = True,
    bubble_detector_model: str = "yolo_2",
    request_coordinator: Optional[Any] = None,
):
    """Same contract as the reference (core/image/cleaning.py:524-588): returns (cleaned BGR(A) ndarray, list of
    per-bubble dicts with keys mask/base_mask/color/bbox/is_colored/text_bbox/text_color_bgr/is_sam/inpainted)."""
    try:
        if inpaint_colored_bubbles:
            raise CleaningError("inpaint_colored_bubbles (Flux) is outside the B200 hot path of this build")
        if isinstance(image_input, (str, Path)):
            pil_image = Image.open(image_input)
            image_path = image_input
        else:
            pil_image = image_input
            image_path = None
        image = _pil_to_bgr(pil_image)
        img_h, img_w = image.shape[:2]

        if pre_computed_detections is not None:
            detections = pre_computed_detections
        elif image_path is not None:
            from mangatranslator_b200.core.image.detection import detect_speech_bubbles
            res = detect_speech_bubbles(image_path, model_path, confidence, device=device,
                                        conjoined_confidence=conjoined_confidence,
                                        osb_text_verification=osb_text_verification,
                                        osb_text_hf_token=osb_text_hf_token,
                                        bubble_detector_model=bubble_detector_model)
            detections = res[0] if isinstance(res, tuple) else res
        else:
            raise ValidationError("Bubble detection requires an image path, but an image object was provided "
                                  "without pre-computed detections.")

        dev = _require_cuda()
        page = torch.from_numpy(image).to(dev)
        prepared = _masks_from_detections(list(detections), img_h, img_w)
        submit = []
        for d, is_sam in prepared:
            if d.get("sam_mask") is None:
                log_message(f"Skipping detection {d.get('bbox')}: no mask points", verbose=verbose)
            submit.append(d)
        batch = clean_pages_device([page], [submit], thresholding_value=thresholding_value,
                                   use_otsu_threshold=use_otsu_threshold, roi_shrink_px=roi_shrink_px,
                                   processing_scale=processing_scale)
        processed = []
        for k, ((d, is_sam), res) in enumerate(zip(prepared, batch.results[0])):
            if res is None:
                continue
            if res.status != 0:
                log_message(f"Error processing {'SAM' if is_sam else 'YOLO'} mask for detection {d.get('bbox')}: "
                            f"{STATUS_NAMES.get(res.status, res.status)}", always_print=True)
                continue
            mask = batch.export_mask(0, k).cpu().numpy()
            m = d["sam_mask"]
            m_np = m if isinstance(m, np.ndarray) else m.cpu().numpy()
            base = np.where(m_np > 0, 255, 0).astype(np.uint8)
            processed.append(_result_to_dict(res, d, base, mask, is_sam, image.shape[2]))
            log_message(f"Detection {d.get('bbox')}: processed successfully", verbose=verbose)
        cleaned = batch.pages_out[0].cpu().numpy()
        log_message(f"Cleaned {len(processed)} speech bubbles", always_print=True)
        return cleaned, processed
    except IOError as e:
        raise ImageProcessingError(f"Error loading image {image_input}: {str(e)}")
    except (CleaningError, ValidationError) as e:
        raise CleaningError(f"Error cleaning speech bubbles: {str(e)}")
    except Exception as e:
        raise CleaningError(f"Error cleaning speech bubbles: {str(e)}")


@serialized
def retry_cleaning_with_otsu(image_bgr: np.ndarray, bubble_info: dict, thresholding_value: int, roi_shrink_px: int,
                             processing_scale: float = 1.0, verbose: bool = False,
                             classify_colored: bool = False) -> Optional[dict]:
    """Single-bubble Otsu retry with the reference's contract (core/image/cleaning.py:1051-1170)."""
    base_mask = bubble_info.get("base_mask")
    if base_mask is None:
        log_message(f"Otsu retry skipped for {bubble_info.get('bbox')}: missing base_mask", verbose=verbose)
        return None
    try:
        dev = _require_cuda()
        page = torch.from_numpy(np.ascontiguousarray(image_bgr)).to(dev)
        det = {"bbox": bubble_info.get("bbox"), "sam_mask": base_mask,
               "conjoined_neighbor_bboxes": bubble_info.get("neighbor_bboxes")}
        params = H.build_params(thresholding_value, True, roi_shrink_px, processing_scale, retry_otsu=False)
        batch = clean_batch([page], [[det]], params)
        res = batch.results[0][0]
        if res is None or res.status != 0:
            log_message(f"Otsu retry cleaning failed for {bubble_info.get('bbox')}", always_print=True)
            return None
        mask = batch.export_mask(0, 0).cpu().numpy()
        out = _result_to_dict(res, det, np.where(np.asarray(base_mask) > 0, 255, 0).astype(np.uint8), mask,
                              bubble_info.get("is_sam", False), image_bgr.shape[2])
        out.pop("inpainted", None)
        return out
    except Exception as e:  # the reference swallows unexpected errors here too (cleaning.py:1135-1140)
        log_message(f"Otsu retry cleaning unexpected error for {bubble_info.get('bbox')}: {e}", always_print=True)
        return None


@serialized
def process_single_bubble(base_mask, img_gray, img_height, img_width, thresholding_value, use_otsu_threshold,
                          roi_shrink_px, verbose, detection_bbox=None, is_sam=False, dilation_kernel=None,
                          constraint_erosion_kernel=None, min_contour_area: float = MIN_CONTOUR_AREA,
                          classify_colored: bool = False, neighbor_bboxes: Optional[list] = None,
                          processing_scale: float = 1.0, image_bgr: Optional[np.ndarray] = None):
    """Reference signature (core/image/cleaning.py:210-228).  `roi_shrink_px` is the EFFECTIVE (already scaled)
    shrink, kernels are given as cv2 arrays whose size is what matters.  Returns the reference's 6-tuple."""
    if image_bgr is None:
        image_bgr = np.repeat(np.asarray(img_gray)[:, :, None], 3, axis=2)
    dev = _require_cuda()
    params = H.build_params(thresholding_value, use_otsu_threshold, 5, processing_scale, retry_otsu=False)
    kd = int(dilation_kernel.shape[0]) if dilation_kernel is not None else H.DILATION_KERNEL_SIZE[0]
    ke = int(constraint_erosion_kernel.shape[0]) if constraint_erosion_kernel is not None else H.EROSION_KERNEL_SIZE[0]
    params.kd, params.ke = kd, ke
    for i in range(H.MAX_SE):
        params.sed_hw[i] = params.see_hw[i] = -1
    for i, wv in enumerate(H.ellipse_rows(kd)):
        params.sed_hw[i] = wv
    for i, wv in enumerate(H.ellipse_rows(ke)):
        params.see_hw[i] = wv
    r, rows = H.chamfer_ball_rows(float(roi_shrink_px))
    params.ball_r = r
    for i in range(2 * H.MAX_BALL + 1):
        params.ball_hw[i] = -1
    for i, wv in enumerate(rows):
        params.ball_hw[i] = wv
    params.min_area = float(min_contour_area)
    params.margin = kd // 2 + max(r, params.jball_r) + 3
    page = torch.from_numpy(np.ascontiguousarray(image_bgr)).to(dev)
    det = {"bbox": detection_bbox, "sam_mask": np.asarray(base_mask), "conjoined_neighbor_bboxes": neighbor_bboxes}
    batch = clean_batch([page], [[det]], params)
    res = batch.results[0][0]
    if res is None or res.status != 0:
        log_message(f"Failed to process {'SAM' if is_sam else 'YOLO'} mask for {detection_bbox}", always_print=True)
        raise CleaningError("Failed to process bubble mask")
    d = _result_to_dict(res, det, None, batch.export_mask(0, 0).cpu().numpy(), is_sam, image_bgr.shape[2])
    return d["mask"], d["color"], False, d["color"], d["text_bbox"], d["text_color_bgr"]
