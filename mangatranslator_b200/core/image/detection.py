"""Speech-bubble detection + segmentation stage — drop-in for the reference's core/image/detection.py on the hot path.

`detect_speech_bubbles` keeps the reference signature and result dictionaries (:1263-1277, :1119-1131): YOLO call
(:1338-1345) -> IoU de-duplication (:219-254) -> containment filter (:257-295) -> [SAM 2.1 box-prompted masks
(:475-511) -> floor/ceil box clip (:1732-1750)] -> detection dicts.  The model objects come from the ModelManager and
are duck-typed exactly like the reference's, so injected ultralytics/transformers objects work as well; with the
B200-native objects everything between the page upload and the final masks stays on the device
(`detect_pages_device`).

Out of scope here: the RT-DETRv2 conjoined-bubble branch (its load failure is swallowed exactly like
the reference does at :1541-1548, so all bubbles are treated as simple), OSB-text verification, SAM3, panel detection.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200.core.caching import get_cache
from mangatranslator_b200.core.device import get_best_device
from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
from mangatranslator_b200.utils.exceptions import ImageProcessingError, ModelError
from mangatranslator_b200.utils.logging import log_message

IOA_THRESHOLD = 0.50
SAM_MASK_THRESHOLD = 0.5
IOU_DUPLICATE_THRESHOLD = 0.7


# ---- box geometry (host floats, like the reference's Python helpers :44-60,204-216) ------------------------------
def _box_intersection_area(a, b) -> float:
    return max(0.0, min(a[2], b[2]) - max(a[0], b[0])) * max(0.0, min(a[3], b[3]) - max(a[1], b[1]))


def _box_area(b) -> float:
    return max(0.0, b[2] - b[0]) * max(0.0, b[3] - b[1])


def _calculate_ioa(box_inner, box_outer) -> float:
    a = _box_area(box_inner)
    return 0.0 if a <= 0 else _box_intersection_area(box_inner, box_outer) / a


def _calculate_iou(box_a, box_b) -> float:
    inter = _box_intersection_area(box_a, box_b)
    union = _box_area(box_a) + _box_area(box_b) - inter
    return inter / union if union > 0 else 0.0


def _deduplicate_primary_boxes(boxes: torch.Tensor, confidences: torch.Tensor, threshold: float):
    """Greedy IoU de-duplication in descending-confidence order; returns (kept boxes, kept indices)."""
    if len(boxes) <= 1:
        return boxes, list(range(len(boxes)))
    bl, cl = boxes.tolist(), confidences.tolist()
    keep: List[int] = []
    for i in sorted(range(len(bl)), key=lambda k: cl[k], reverse=True):
        if all(_calculate_iou(bl[i], bl[k]) <= threshold for k in keep):
            keep.append(i)
    return boxes[keep], keep


def _remove_contained_boxes(boxes: torch.Tensor, indices=None, threshold: float = 0.9):
    """Drop box i when IoA(i in j) > threshold for a still-kept j (order dependent, like the reference)."""
    if indices is None:
        indices = [("primary", i) for i in range(len(boxes))]
    if len(boxes) <= 1:
        return boxes, indices
    bl = boxes.tolist()
    n = len(bl)
    alive = [True] * n
    for i in range(n):
        if not alive[i]:
            continue
        for j in range(n):
            if i != j and alive[j] and _calculate_ioa(bl[i], bl[j]) > threshold:
                alive[i] = False
                break
    return boxes[alive], [indices[i] for i in range(n) if alive[i]]


def _build_rect_mask_from_box(box, img_h: int, img_w: int) -> np.ndarray:
    x0f, y0f, x1f, y1f = box.tolist() if hasattr(box, "tolist") else box
    x0, y0 = int(np.floor(max(0, min(x0f, img_w)))), int(np.floor(max(0, min(y0f, img_h))))
    x1, y1 = int(np.ceil(max(0, min(x1f, img_w)))), int(np.ceil(max(0, min(y1f, img_h))))
    m = np.zeros((img_h, img_w), np.uint8)
    if x1 > x0 and y1 > y0:
        m[y0:y1, x0:x1] = 255
    return m


def _fallback_to_yolo_mask(primary_results, i, mask_type="binary"):
    masks = getattr(primary_results, "masks", None)
    if masks is None:
        return None
    try:
        if len(masks) <= i:
            return None
        if mask_type == "points":
            pts = masks[i].xy[0]
            return pts.tolist() if hasattr(pts, "tolist") else pts
        t = masks.data[i]
        oh, ow = primary_results.orig_shape
        if tuple(t.shape[-2:]) != (oh, ow):
            t = torch.nn.functional.interpolate(t.float()[None, None], size=(oh, ow), mode="bilinear",
                                                align_corners=False)[0, 0]
        return (t > SAM_MASK_THRESHOLD).cpu().numpy().astype(np.uint8) * 255
    except (IndexError, AttributeError) as e:
        log_message(f"Could not extract YOLO mask for detection {i}: {e}", always_print=True)
        return None


def _process_simple_bubbles(image, primary_boxes, simple_indices, processor, sam_model, device):
    """Same call sequence as the reference (:475-511) on whatever (processor, model) pair the ModelManager returned."""
    if not simple_indices:
        return []
    boxes = primary_boxes[simple_indices].unsqueeze(0).cpu()
    inputs = processor(image, input_boxes=boxes, return_tensors="pt")
    for key in inputs:
        if isinstance(inputs[key], torch.Tensor) and inputs[key].is_floating_point():
            inputs[key] = inputs[key].to(sam_model.dtype)
    inputs = inputs.to(device)
    with torch.no_grad():
        outputs = sam_model(multimask_output=False, **inputs)
    masks_tensor = processor.post_process_masks(outputs.pred_masks, inputs["original_sizes"])[0][:, 0]
    return [m for m in (masks_tensor > SAM_MASK_THRESHOLD).cpu().numpy()]


def _class_name(model, results, idx) -> str:
    try:
        names = getattr(model, "names", None) or getattr(results, "names", {})
        return names[int(results.boxes.cls[idx])]
    except Exception:
        return "speech_bubble"


def detect_speech_bubbles(image_path: Path, model_path, confidence=0.6, verbose=False, device=None,
                          seg_model: str = "yolo", conjoined_detection: bool = True, conjoined_confidence=0.35,
                          image_override: Optional[Image.Image] = None, osb_enabled: bool = False,
                          osb_text_verification: bool = False, osb_text_hf_token: str = "",
                          bubble_detector_model: str = "yolo_2") -> Tuple[List[dict], List[List[float]]]:
    detections: List[dict] = []
    text_free_boxes: List[List[float]] = []
    _device = device if device is not None else get_best_device()
    try:
        if image_override is not None:
            image_pil = image_override if image_override.mode == "RGB" else image_override.convert("RGB")
        else:
            image_pil = Image.open(str(image_path)).convert("RGB")
        image_cv = np.ascontiguousarray(np.asarray(image_pil)[:, :, ::-1])       # RGB -> BGR like cv2.cvtColor
    except Exception as e:
        raise ImageProcessingError(f"Error loading image: {e}")
    img_h, img_w = image_cv.shape[:2]
    mm = get_model_manager()
    cache = get_cache()
    try:
        primary_model = mm.load_yolo_speech_bubble(model_path)
    except Exception as e:
        raise ModelError(f"Error loading primary model: {e}")

    key = cache.get_yolo_cache_key(image_pil, model_path, confidence)
    cached = cache.get_yolo_detection(key)
    if cached is not None:
        primary_results, primary_boxes = cached
    else:
        imgsz = 1600 if bubble_detector_model == "yolo_2" else 640
        primary_results = primary_model(image_cv, conf=confidence, device=_device, verbose=False, imgsz=imgsz,
                                        retina_masks=True)[0]
        primary_boxes = primary_results.boxes.xyxy if primary_results.boxes is not None else torch.tensor([])
        cache.set_yolo_detection(key, (primary_results, primary_boxes))
    sources = [("primary", i) for i in range(len(primary_boxes))]
    if len(primary_boxes) > 1:
        n0 = len(primary_boxes)
        primary_boxes, keep = _deduplicate_primary_boxes(primary_boxes, primary_results.boxes.conf, IOU_DUPLICATE_THRESHOLD)
        sources = [sources[i] for i in keep]
        if len(primary_boxes) < n0:
            log_message(f"Removed {n0 - len(primary_boxes)} duplicate detections", verbose=verbose)
    if len(primary_boxes) > 1:
        n0 = len(primary_boxes)
        primary_boxes, sources = _remove_contained_boxes(primary_boxes, sources)
        if len(primary_boxes) < n0:
            log_message(f"Removed {n0 - len(primary_boxes)} contained detections", verbose=verbose)
    if len(primary_boxes) == 0:
        log_message("No detections found", verbose=verbose)
        return detections, text_free_boxes
    log_message(f"Detected {len(primary_boxes)} speech bubbles with YOLO", always_print=True)

    if conjoined_detection:
        try:
            mm.load_rtdetr_conjoined_bubble()
            log_message("Secondary conjoined-bubble detector injected but its merge logic is not part of this build",
                        verbose=verbose)
        except Exception as e:   # the reference swallows a secondary-model failure and keeps the primaries (:1541-1548)
            log_message(f"Secondary detection skipped: {e}", verbose=verbose)

    simple = list(range(len(primary_boxes)))
    sam_masks: Optional[list] = None
    if seg_model in ("sam2", "sam3"):
        if seg_model == "sam3":
            raise ModelError("SAM3 is outside the B200 hot path of this build")
        try:
            processor, sam_model = mm.load_sam2(verbose=verbose)
            raw = _process_simple_bubbles(image_pil, primary_boxes, simple, processor, sam_model, _device)
            sam_masks = []
            for m, box in zip(raw, primary_boxes):
                clip = _build_rect_mask_from_box(box, img_h, img_w) > 0        # floor/ceil box clip (:1732-1750)
                sam_masks.append(np.where(np.logical_and(m, clip), 255, 0).astype(np.uint8))
        except ModelError:
            raise
        except Exception as e:    # SAM failure -> YOLO masks -> rectangles (:1783-1813)
            log_message(f"SAM2 segmentation failed: {e}. Falling back to YOLO segmentation masks.", always_print=True)
            sam_masks = None
    for k in simple:
        source, orig = sources[k]
        box = primary_boxes[k]
        mask = None
        if sam_masks is not None and sam_masks[k] is not None:
            mask = sam_masks[k]
        else:
            mask = _fallback_to_yolo_mask(primary_results, orig, "binary")
        if mask is None:
            mask = _build_rect_mask_from_box(box, img_h, img_w)
        x0, y0, x1, y1 = box.tolist()
        detections.append({"bbox": (int(round(x0)), int(round(y0)), int(round(x1)), int(round(y1))),
                           "confidence": float(primary_results.boxes.conf[orig]),
                           "class": _class_name(primary_model, primary_results, orig), "sam_mask": mask})
    return detections, text_free_boxes


# ---- device-resident fast path used by the batch pipeline / bench -------------------------------------------------------
def detect_pages_device(pages_bgr: List[torch.Tensor], *, confidence: float = 0.6, imgsz: int = 1600,
                        seg_model: str = "sam2", injected_boxes: Optional[List[np.ndarray]] = None,
                        own_masks: bool = False):
    """pages_bgr: device uint8 HxWx3 (BGR, like the cleaning stage wants).  Per page: letterbox -> YOLO graph -> decode
    + NMS + scale_boxes + reference dedup/containment (all on device, one small D2H of the box table) -> SAM 2.1
    masks on device.  Returns per page a list of detection dicts whose `sam_mask` is a DEVICE uint8 tensor and which
    carry `mask_bbox` so the cleaning stage needs no further host work.  `injected_boxes` (per page [P,4] float32
    original-pixel boxes) bypasses the detector output (ground-truth boxes for stage-level parity runs).  The SAM masks
    live in the segmenter's static output buffer, which the next page overwrites: pass `own_masks=True` to get a copy
    when detections of several pages must stay alive together."""
    from mangatranslator_b200.preproc import letterbox_device
    mm = get_model_manager()
    yolo = mm.load_yolo_speech_bubble(None)
    sam = None
    if seg_model == "sam2":
        sam = mm.load_sam2()[1].net
    out = []
    for pi, page in enumerate(pages_bgr):
        h, w = int(page.shape[0]), int(page.shape[1])
        lb = letterbox_device(page, imgsz, swap_rb=True)
        g = yolo.forward_letterboxed(lb)
        det, cnt, final_idx = yolo.detect(g, confidence, (h, w), tuple(lb.shape[:2]), apply_reference_dedup=True)
        n_final = int(cnt[1].item())                                   # the one host sync of the detect stage
        rows = det[final_idx[:n_final].long()].cpu().numpy() if n_final else np.zeros((0, 8), np.float32)
        boxes = rows[:, :4].astype(np.float32)
        confs = rows[:, 4]
        if injected_boxes is not None:
            boxes = np.asarray(injected_boxes[pi], np.float32).reshape(-1, 4)
            confs = np.full((boxes.shape[0],), 0.9, np.float32)
        dets: List[Dict[str, Any]] = []
        masks = None
        if sam is not None and boxes.shape[0]:
            rgb = page[:, :, [2, 1, 0]].contiguous() if page.shape[2] == 3 else page[:, :, [2, 1, 0]].contiguous()
            enc = sam.encode(rgb)
            masks = sam.decode(enc, torch.from_numpy(boxes), (h, w))
            if own_masks:
                masks = masks.clone()
        for k in range(boxes.shape[0]):
            x0, y0, x1, y1 = [float(v) for v in boxes[k]]
            bx0, by0 = int(np.floor(max(0, min(x0, w)))), int(np.floor(max(0, min(y0, h))))
            bx1, by1 = int(np.ceil(max(0, min(x1, w)))), int(np.ceil(max(0, min(y1, h))))
            if masks is not None:
                m = masks[k]
            else:
                m = torch.zeros((h, w), dtype=torch.uint8, device=page.device)
                m[by0:by1, bx0:bx1] = 255
            dets.append({"bbox": (int(round(x0)), int(round(y0)), int(round(x1)), int(round(y1))),
                         "confidence": float(confs[k]), "class": "speech_bubble", "sam_mask": m,
                         "mask_bbox": (bx0, by0, max(bx1, bx0 + 1), max(by1, by0 + 1))})
        out.append(dets)
    return out
