"""Speech-bubble detection + segmentation stage — drop-in for the reference's core/image/detection.py on the hot path.

`detect_speech_bubbles` keeps the reference signature and result dictionaries (:1263-1277, :1119-1131): YOLO call
(:1338-1345) -> IoU de-duplication (:219-254) -> containment filter (:257-295) -> [SAM 2.1 box-prompted masks
(:475-511) -> floor/ceil box clip (:1732-1750)] -> detection dicts.  The model objects come from the ModelManager and
are duck-typed exactly like the reference's, so injected ultralytics/transformers objects work as well; with the
B200-native objects everything between the page upload and the final masks stays on the device
(`detect_pages_device`).

Conjoined bubbles (:345-472, :582-1035, :1075-1260): primaries that overlap each other are grouped into synthetic
conjoined bubbles on every page (:1596-1619), SAM segments the group's union box, and the parent mask is divided between
the members by `mtb_split_conjoined` (mangatranslator_b200/conjoined.py holds the box geometry).  The same code serves
conjoined parents found by the secondary RT-DETRv2 detector (mangatranslator_b200/rtdetr.py, loaded by
`load_rtdetr_conjoined_bubble` when its checkpoint exists; otherwise the load failure is swallowed exactly like the
reference does at :1541-1548 and the primaries are kept).

Secondary ultralytics detectors, executed from their checkpoints' module trees (mangatranslator_b200/yolo_tree.py):
`detect_panels` (:1817-1921, YOLO11-L at imgsz 640, class "frame") and OSB-text verification
`_expand_boxes_with_osb_text` (:120-201, YOLO12x at imgsz 640: a bubble box grows to hold the text box that belongs to
it; the same text boxes steer the cut of a conjoined split away from the text, :700-783, mangatranslator_b200/conjoined.py).
Not restated: SAM3.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200.core.caching import get_cache
from mangatranslator_b200.core.device import get_best_device
from mangatranslator_b200.core.ml.model_manager import get_model_manager
from mangatranslator_b200.utils.exceptions import ImageProcessingError, ModelError
from mangatranslator_b200.utils.logging import log_message
from mangatranslator_b200._lib import serialized

IOA_THRESHOLD = 0.50
SAM_MASK_THRESHOLD = 0.5
IOA_OVERLAP_THRESHOLD = 0.5
IOU_DUPLICATE_THRESHOLD = 0.7
OSB_TEXT_MATCH_IOA_THRESHOLD = 0.2      # :20-22


# ---- box geometry (host floats, like the reference's Python helpers :44-60,204-216) ------------------------------
def _box_intersection_area(a, b) -> float:
    return max(0.0, min(a[2], b[2]) - max(a[0], b[0])) * max(0.0, min(a[3], b[3]) - max(a[1], b[1]))


def _box_area(b) -> float:
    return max(0.0, b[2] - b[0]) * max(0.0, b[3] - b[1])


def _box_contains(inner, outer) -> bool:
    return inner[0] >= outer[0] and inner[1] >= outer[1] and inner[2] <= outer[2] and inner[3] <= outer[3]


def _point_in_box(px: float, py: float, box) -> bool:
    return box[0] <= px <= box[2] and box[1] <= py <= box[3]


def _text_box_meaningfully_matches_box(t_box, b_box) -> bool:
    """:91-106 — a text box belongs to a bubble when a fifth of it lies inside, or its centre does."""
    inter = _box_intersection_area(t_box, b_box)
    area = _box_area(t_box)
    if inter <= 0.0 or area <= 0.0:
        return False
    return inter / area >= OSB_TEXT_MATCH_IOA_THRESHOLD or _point_in_box((t_box[0] + t_box[2]) / 2.0,
                                                                          (t_box[1] + t_box[3]) / 2.0, b_box)


def _expand_boxes_with_osb_text(image_cv, image_pil, primary_boxes: torch.Tensor, cache, model_manager, device,
                                confidence: float, hf_token: str, verbose: bool):
    """:120-201 — every text box found by the OSB-text detector (imgsz 640) is assigned to the bubble it intersects most;
    if it meaningfully belongs there and sticks out, the bubble box grows to the union.  Boxes are updated one text box
    after the other (a later text box sees the grown boxes); any failure leaves the primaries as they were."""
    if primary_boxes is None or len(primary_boxes) == 0:
        return primary_boxes
    try:
        from mangatranslator_b200.core.ml.model_manager import ModelType
        key = cache.get_yolo_cache_key(image_pil, str(model_manager.model_paths[ModelType.YOLO_OSBTEXT]), confidence)
        hit = cache.get_yolo_detection(key)
        if hit is not None:
            _, osb_boxes, _ = hit
        else:
            res = model_manager.load_yolo_osbtext(token=hf_token)(image_cv, conf=confidence, device=device, verbose=False,
                                                                  imgsz=640)[0]
            osb_boxes = res.boxes.xyxy if res.boxes is not None else torch.tensor([])
            osb_confs = res.boxes.conf if res.boxes is not None else torch.tensor([])
            cache.set_yolo_detection(key, (res, osb_boxes, osb_confs))
        if osb_boxes is None or len(osb_boxes) == 0:
            return primary_boxes
        pb = primary_boxes.detach().cpu().numpy()
        for t in osb_boxes.detach().cpu().numpy():
            best, best_inter = None, 0.0
            for i, b in enumerate(pb):
                inter = _box_intersection_area(t, b)
                if inter > best_inter:
                    best, best_inter = i, inter
            if best is None or best_inter <= 0.0:
                continue
            if not _text_box_meaningfully_matches_box(t, pb[best]) or _box_contains(t, pb[best]):
                continue
            b = pb[best]
            pb[best] = [min(b[0], t[0]), min(b[1], t[1]), max(b[2], t[2]), max(b[3], t[3])]
        return torch.tensor(pb, device=primary_boxes.device, dtype=primary_boxes.dtype)
    except Exception as e:
        log_message(f"OSB text verification skipped: {e}", verbose=verbose)
        return primary_boxes


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


def _secondary_detections(mm, image_cv, primary_boxes, sources, conjoined_confidence, device, osb_enabled, verbose):
    """The RT-DETR branch of the reference (:1392-1548): run the secondary detector (mangatranslator_b200.rtdetr.RtDetrB200,
    or whatever duck-typed detector sits in the ModelManager slot), drop nested secondary boxes, route `text_free` boxes
    aside, add bubbles the primary detector missed, remove primaries that are really free text.  A load failure propagates
    to the caller, which keeps the primaries like the reference does.  Returns (primary_boxes, sources, secondary_boxes,
    secondary_sources, secondary_results, text_free_boxes)."""
    text_free: List[List[float]] = []
    model = mm.load_rtdetr_conjoined_bubble()
    results = model(image_cv, conf=conjoined_confidence, device=device, verbose=False, imgsz=640)[0]
    sec = results.boxes.xyxy if results.boxes is not None else torch.tensor([])
    sec_src = [("secondary", i) for i in range(len(sec))]
    if len(sec) > 1:
        sec, sec_src = _remove_contained_boxes(sec, sec_src)
    if len(sec) > 0 and hasattr(model, "names"):
        bubble_id = next((c for c, n in model.names.items() if n == "bubble"), None)
        text_free_id = next((c for c, n in model.names.items() if n == "text_free"), None)
        keep_b, keep_s = [], []
        for i, sb in enumerate(sec):
            cid = int(results.boxes.cls[sec_src[i][1]])
            if text_free_id is not None and cid == text_free_id:
                text_free.append(sb.tolist())
            elif bubble_id is None or cid == bubble_id:
                keep_b.append(sb)
                keep_s.append(sec_src[i])
        sec = torch.stack(keep_b) if keep_b else sec[:0]
        sec_src = keep_s
        if len(sec) > 0:                                   # bubbles the primary detector missed (:1456-1497)
            plist = primary_boxes.tolist() if len(primary_boxes) > 0 else []
            missed = [(sb, sec_src[i]) for i, sb in enumerate(sec)
                      if not any(_calculate_ioa(sb.tolist(), pb) > IOA_OVERLAP_THRESHOLD or
                                 _calculate_ioa(pb, sb.tolist()) > IOA_OVERLAP_THRESHOLD for pb in plist)]
            if missed:
                log_message(f"Found {len(missed)} missed bubbles from secondary model", always_print=True)
                extra = torch.stack([m[0] for m in missed])
                primary_boxes = torch.cat((primary_boxes, extra.to(primary_boxes.device)), dim=0) if len(primary_boxes) else extra
                sources = sources + [m[1] for m in missed]
    if text_free and len(primary_boxes) > 0:               # primaries that are really free text (:1499-1538)
        plist = primary_boxes.tolist()
        keep = [i for i, pb in enumerate(plist)
                if not any(_calculate_ioa(pb, tf) > IOA_OVERLAP_THRESHOLD or _calculate_ioa(tf, pb) > IOA_OVERLAP_THRESHOLD
                           for tf in text_free)]
        if len(keep) < len(plist):
            action = "routing to OSB pipeline" if osb_enabled else "discarding (OSB disabled)"
            log_message(f"Removing {len(plist) - len(keep)} bubbles marked text_free ({action})", always_print=True)
            primary_boxes = primary_boxes[keep] if keep else torch.tensor([])
            sources = [sources[i] for i in keep]
    return primary_boxes, sources, sec, sec_src, results, text_free


def _detection_metadata(source, orig, primary_results, primary_model, secondary_results, conjoined_confidence):
    """(:1038-1072) confidence / class name of a box by where it came from."""
    if source == "primary" and primary_results is not None and len(primary_results.boxes) > 0:
        k = min(orig, len(primary_results.boxes.conf) - 1)
        return float(primary_results.boxes.conf[k]), _class_name(primary_model, primary_results, k)
    if source == "secondary" and secondary_results is not None and len(secondary_results.boxes) > 0:
        k = min(orig, len(secondary_results.boxes.conf) - 1)
        names = getattr(secondary_results, "names", None)
        cid = int(secondary_results.boxes.cls[k])
        return float(secondary_results.boxes.conf[k]), (names.get(cid, "speech_bubble") if names is not None else "speech_bubble")
    return conjoined_confidence, "speech_bubble"


def _round_bbox(box) -> Tuple[int, int, int, int]:
    x0, y0, x1, y1 = box.tolist() if hasattr(box, "tolist") else box
    return (int(round(x0)), int(round(y0)), int(round(x1)), int(round(y1)))


def _split_group_on_device(parent_mask: np.ndarray, group_boxes, device, text_boxes=None) -> List[np.ndarray]:
    """`_split_conjoined_mask` of the parent (with the child rectangles ORed in, :1161-1164) through the CUDA kernel;
    `text_boxes` = the group's OSB text boxes (`_get_group_osb_text_boxes`), which move the cut off the text."""
    from mangatranslator_b200.conjoined import split_conjoined_device
    dev = get_model_manager()._require_cuda()
    kw = {} if text_boxes is None else {"text_boxes": text_boxes}
    out, _ = split_conjoined_device(torch.from_numpy(np.ascontiguousarray(parent_mask)).to(dev), group_boxes, **kw)
    return [m for m in out.cpu().numpy()]


def _get_cached_osb_text_boxes(cache, model_manager, image_pil, confidence):
    """:298-314 — the OSB text boxes `_expand_boxes_with_osb_text` left in the detection cache, as an array, or None."""
    try:
        from mangatranslator_b200.core.ml.model_manager import ModelType
        key = cache.get_yolo_cache_key(image_pil, str(model_manager.model_paths[ModelType.YOLO_OSBTEXT]), confidence)
        hit = cache.get_yolo_detection(key)
        if hit is not None:
            _, osb_boxes, _ = hit
            if osb_boxes is not None and len(osb_boxes) > 0:
                return osb_boxes.detach().cpu().numpy() if hasattr(osb_boxes, "detach") else np.asarray(osb_boxes)
    except Exception:
        pass
    return None


def _build_segmentation_detections(primary_boxes, grouping_boxes, sources, primary_results, primary_model, secondary_boxes,
                                   secondary_sources, secondary_results, simple_indices, conjoined_indices, img_h, img_w,
                                   conjoined_confidence, sam_masks=None, synthetic_groups=None, device=None,
                                   osb_text_boxes_np=None) -> List[dict]:
    """(:1075-1260) simple boxes first, then the children of every conjoined parent, then the synthetic groups."""
    from mangatranslator_b200.conjoined import group_osb_text_boxes, union_box
    dets: List[dict] = []

    def meta(src):
        return _detection_metadata(src[0], src[1], primary_results, primary_model, secondary_results, conjoined_confidence)

    def own_or_yolo_mask(idx):
        if sam_masks is not None and sam_masks[idx] is not None:
            return sam_masks[idx]
        if sources[idx][0] == "primary":
            return _fallback_to_yolo_mask(primary_results, sources[idx][1], "binary")
        return None

    for idx in simple_indices:
        conf, cls = meta(sources[idx])
        mask = own_or_yolo_mask(idx)
        if mask is None:
            mask = _build_rect_mask_from_box(primary_boxes[idx], img_h, img_w)
        dets.append({"bbox": _round_bbox(primary_boxes[idx]), "confidence": conf, "class": cls, "sam_mask": mask})

    def emit_group(parent_mask, group_boxes, group_sources, parent_box):
        group_osb = group_osb_text_boxes(osb_text_boxes_np, parent_box)          # :1164, :1218
        if group_osb is None:
            masks = _split_group_on_device(parent_mask, group_boxes, device)
        else:
            masks = _split_group_on_device(parent_mask, group_boxes, device, group_osb)
        bboxes = [_round_bbox(b) for b in group_boxes]
        for k, src in enumerate(group_sources):
            conf, cls = meta(src)
            dets.append({"bbox": bboxes[k], "confidence": conf, "class": cls, "sam_mask": masks[k],
                         "conjoined_neighbor_bboxes": [b for n, b in enumerate(bboxes) if n != k]})

    for p_idx, s_indices in conjoined_indices:
        parent_box = union_box(torch.cat([primary_boxes[p_idx].unsqueeze(0)] +
                                         [secondary_boxes[s].unsqueeze(0) for s in s_indices], dim=0))
        parent_mask = own_or_yolo_mask(p_idx)
        if parent_mask is None:
            parent_mask = _build_rect_mask_from_box(parent_box, img_h, img_w)
        emit_group(parent_mask, [secondary_boxes[s] for s in s_indices], [secondary_sources[s] for s in s_indices], parent_box)
    for sg in synthetic_groups or []:
        parent_mask = sg.get("parent_mask")
        if parent_mask is None:
            parent_mask = _build_rect_mask_from_box(sg["parent_box"], img_h, img_w)
        members = sg["member_indices"]
        emit_group(parent_mask, [grouping_boxes[m] for m in members], [sources[m] for m in members], sg["parent_box"])
    return dets


@serialized
def detect_speech_bubbles(image_path: Path, model_path, confidence=0.6, verbose=False, device=None,
                          seg_model: str = "yolo", conjoined_detection: bool = True, conjoined_confidence=0.35,
                          image_override: Optional[Image.Image] = None, osb_enabled: bool = False,
                          osb_text_verification: bool = False, osb_text_hf_token: str = "",
                          bubble_detector_model: str = "yolo_2") -> Tuple[List[dict], List[List[float]]]:
    from mangatranslator_b200.conjoined import categorize_detections, detect_overlapping_primaries, union_box
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

    secondary_boxes, secondary_sources, secondary_results = torch.tensor([]), [], None
    if conjoined_detection:
        try:
            (primary_boxes, sources, secondary_boxes, secondary_sources, secondary_results,
             text_free_boxes) = _secondary_detections(mm, image_cv, primary_boxes, sources, conjoined_confidence, _device,
                                                      osb_enabled, verbose)
        except Exception as e:   # the reference swallows a secondary-model failure and keeps the primaries (:1541-1548)
            log_message(f"Warning: Could not load/run secondary RT-DETR model: {e}. "
                        "Proceeding without conjoined/fallback detection.", verbose=verbose)
            secondary_boxes, secondary_sources, secondary_results = torch.tensor([]), [], None
    if len(primary_boxes) == 0:
        return detections, text_free_boxes
    primary_boxes = primary_boxes.detach().float().cpu()
    grouping_boxes = primary_boxes.clone()
    osb_text_boxes_np = None
    if osb_text_verification and len(primary_boxes) > 0:       # :1555-1571 (grouping keeps the un-expanded boxes)
        primary_boxes = _expand_boxes_with_osb_text(image_cv, image_pil, primary_boxes, cache, mm, _device, confidence,
                                                    osb_text_hf_token, verbose)
        osb_text_boxes_np = _get_cached_osb_text_boxes(cache, mm, image_pil, confidence)   # they also steer the split cuts

    conjoined_indices: list = []
    simple = list(range(len(primary_boxes)))
    if len(secondary_boxes) > 0 and conjoined_detection:
        secondary_boxes = secondary_boxes.detach().float().cpu()
        conjoined_indices, simple = categorize_detections(grouping_boxes, secondary_boxes, IOA_THRESHOLD)
        if conjoined_indices:
            log_message(f"Detected {len(conjoined_indices)} conjoined speech bubbles with RT-DETR", always_print=True)
    # primaries that overlap each other = one bubble cut into sections by the detector (:1596-1619); no secondary model needed
    synthetic_groups: List[dict] = []
    if len(simple) > 1:
        groups, simple = detect_overlapping_primaries(grouping_boxes, simple)
        from mangatranslator_b200.conjoined import MAX_CHILDREN
        groups, simple = _ungroup_oversized(groups, simple, MAX_CHILDREN, verbose)
        if groups:
            log_message(f"Detected {len(groups)} synthetic conjoined group(s) from "
                        f"{sum(len(g) for g in groups)} overlapping primary detections", always_print=True)
        for members in groups:
            synthetic_groups.append({"member_indices": members, "parent_box": union_box(grouping_boxes[members]),
                                     "parent_mask": None})

    def assemble(sam_masks):
        return _build_segmentation_detections(primary_boxes, grouping_boxes, sources, primary_results, primary_model,
                                              secondary_boxes, secondary_sources, secondary_results, simple,
                                              conjoined_indices, img_h, img_w, conjoined_confidence, sam_masks=sam_masks,
                                              synthetic_groups=synthetic_groups, device=_device,
                                              osb_text_boxes_np=osb_text_boxes_np)

    if seg_model not in ("sam2", "sam3"):
        return assemble(None), text_free_boxes
    if seg_model == "sam3":
        raise ModelError("SAM3 is outside the B200 hot path of this build")
    try:
        sam_key = cache.get_sam_cache_key(image_pil, primary_boxes, seg_model, conjoined_detection, conjoined_confidence)
        hit = cache.get_sam_masks(sam_key)
        if hit is not None:
            return hit, text_free_boxes
        processor, sam_model = mm.load_sam2(verbose=verbose)
        # one SAM batch: simple boxes, the combined parent box of every conjoined group, the synthetic parents (:1660-1693)
        prompts, owner = [primary_boxes[i] for i in simple], list(simple)
        for p_idx, s_indices in conjoined_indices:
            prompts.append(union_box(torch.cat([primary_boxes[p_idx].unsqueeze(0)] +
                                               [secondary_boxes[s].unsqueeze(0) for s in s_indices], dim=0)))
            owner.append(p_idx)
        synth_start = len(prompts)
        prompts += [sg["parent_box"] for sg in synthetic_groups]
        sam_masks: List[Optional[np.ndarray]] = [None] * len(primary_boxes)
        if prompts:
            raw = _process_simple_bubbles(image_pil, torch.stack(prompts), list(range(len(prompts))), processor, sam_model,
                                          _device)
            for i, (m, box) in enumerate(zip(raw, prompts)):
                clip = _build_rect_mask_from_box(box, img_h, img_w) > 0        # floor/ceil box clip (:1732-1750)
                clipped = np.logical_and(m, clip).astype(np.uint8) * 255 if clip.any() else np.asarray(m).astype(np.uint8) * 255
                if i < synth_start:
                    sam_masks[owner[i]] = clipped
                else:
                    synthetic_groups[i - synth_start]["parent_mask"] = clipped
            log_message(f"Generated {len(raw)} primary masks with SAM 2.1", always_print=True)
        detections = assemble(sam_masks)
        cache.set_sam_masks(sam_key, detections)
        return detections, text_free_boxes
    except ModelError:
        raise
    except Exception as e:    # SAM failure -> YOLO masks -> rectangles (:1783-1813)
        log_message(f"SAM 2.1 segmentation failed: {e}. Falling back to YOLO segmentation masks.", always_print=True)
        for sg in synthetic_groups:
            sg["parent_mask"] = None
        return assemble(None), text_free_boxes


@serialized
def detect_panels(image_path: Path, confidence: float = 0.25, device=None, verbose=False,
                  image_override: Optional[Image.Image] = None) -> List[Tuple[int, int, int, int]]:
    """:1817-1921 — manga panels: the panel detector at imgsz 640, detections of class "frame" (all classes when the
    model has no such class) as rounded xyxy tuples.  Image loading errors raise ImageProcessingError and a model that
    cannot be loaded ModelError; a failure while detecting is logged and gives [] like the reference."""
    _device = device if device is not None else get_best_device()
    try:
        if image_override is not None:
            image_pil = image_override if image_override.mode == "RGB" else image_override.convert("RGB")
        else:
            image_pil = Image.open(str(image_path)).convert("RGB")
        image_cv = np.ascontiguousarray(np.asarray(image_pil)[:, :, ::-1])
    except Exception as e:
        raise ImageProcessingError(f"Error loading image: {e}")
    try:
        panel_model = get_model_manager().load_yolo_panel(verbose=verbose)
    except Exception as e:
        raise ModelError(f"Error loading panel model: {e}")
    try:
        res = panel_model(image_cv, conf=confidence, device=_device, verbose=False, imgsz=640)[0]
        boxes = res.boxes.xyxy if res.boxes is not None else torch.tensor([])
        classes = res.boxes.cls if res.boxes is not None else torch.tensor([])
        if len(boxes) == 0:
            log_message("No panels detected", verbose=verbose)
            return []
        frame_id = next((cid for cid, name in getattr(panel_model, "names", {}).items() if str(name).lower() == "frame"), None)
        boxes_l, classes_l = boxes.detach().cpu().tolist(), classes.detach().cpu().tolist()
        panels = [tuple(int(round(v)) for v in b) for b, c in zip(boxes_l, classes_l) if frame_id is None or int(c) == frame_id]
        log_message(f"Detected {len(panels)} panels", verbose=verbose)
        return panels
    except Exception as e:
        log_message(f"Panel detection failed: {e}. Proceeding without panel information.", always_print=True)
        return []


def _ungroup_oversized(groups, simple, limit: int, verbose: bool = False):
    """The device splitter divides a parent mask between at most MTB_SPLIT_MAX_CHILDREN members.  A larger synthetic group
    (dozens of mutually overlapping boxes: a detector gone wrong, not a page layout) is not split: its members are
    segmented one by one like simple bubbles — the page keeps its other bubbles instead of failing as a whole."""
    keep = [g for g in groups if len(g) <= limit]
    big = [g for g in groups if len(g) > limit]
    if big:
        log_message(f"Warning: {len(big)} conjoined group(s) with more than {limit} members are not split; their "
                    f"{sum(len(g) for g in big)} boxes are segmented individually", always_print=True)
        simple = sorted(list(simple) + [i for g in big for i in g])
    return keep, simple



# ---- device-resident fast path used by the batch pipeline / bench -------------------------------------------------------
@serialized
def detect_pages_device(pages_bgr: List[torch.Tensor], *, confidence: float = 0.6, imgsz: int = 1600,
                        seg_model: str = "sam2", injected_boxes: Optional[List[np.ndarray]] = None,
                        own_masks: bool = False, conjoined_detection: bool = False, conjoined_confidence: float = 0.35):
    """pages_bgr: device uint8 HxWx3 (BGR, like the cleaning stage wants).  Per page: letterbox -> YOLO graph -> decode
    + NMS + scale_boxes + reference dedup/containment (all on device, one small D2H of the box table) -> [secondary
    RT-DETR detections merged like :1392-1548 when `conjoined_detection`] -> grouping (conjoined parents, synthetic
    groups of overlapping primaries) -> SAM 2.1 masks on device -> parent masks divided on device.  Returns per page a
    list of detection dicts in the reference's order (simple, conjoined children, synthetic children) whose `sam_mask`
    is a DEVICE uint8 tensor and which carry `mask_bbox` so the cleaning stage needs no further host work.
    `injected_boxes` (per page [P,4] float32 original-pixel boxes) bypasses the primary detector's output (ground-truth
    boxes for stage-level parity runs).  The SAM masks live in the segmenter's static output buffer, which the next
    page overwrites: pass `own_masks=True` to get a copy when detections of several pages must stay alive together."""
    from mangatranslator_b200.conjoined import (categorize_detections, detect_overlapping_primaries, split_conjoined_device,
                                               union_box)
    from mangatranslator_b200.preproc import letterbox_device
    mm = get_model_manager()
    yolo = mm.load_yolo_speech_bubble(None)
    sam = None
    if seg_model == "sam2":
        sam = mm.load_sam2()[1].net
    out = []
    # Phase 1 (several pages): the detector of EVERY page is enqueued first — letterbox, YOLO graph, decode, NMS, the
    # reference's dedup, final rows gathered into a per-page slot (the graph's own output buffers are overwritten by the next
    # page) — and the host reads all box tables with ONE synchronisation; phase 2 then runs segmentation page by page with no
    # further host wait, so the grouping of page i + 1 on the host overlaps the device work of page i.  A single page keeps
    # the older order: its SAM encoder is enqueued before the host waits for the box table.
    npages = len(pages_bgr)
    tables = torch.zeros((npages, 300, 8), dtype=torch.float32, device=pages_bgr[0].device) if npages else None
    counts = torch.zeros((npages,), dtype=torch.int32, device=pages_bgr[0].device) if npages else None
    enc0 = None
    import os as _os
    # (the hard-wired YOLOv8-seg plan only: the module-tree executor has been verified page by page)
    batched = (npages > 1 and _os.environ.get("MTB200_YOLO_BATCH", "1") != "0" and type(yolo).__name__ == "YoloB200"
               and len({tuple(p.shape) for p in pages_bgr}) == 1)
    if batched:
        # pages of one size share ONE detector plan with a batch dimension (the deep layers of a single page leave most
        # SMs idle); decode / NMS / dedup run for the whole batch in their own batched launches
        h, w = int(pages_bgr[0].shape[0]), int(pages_bgr[0].shape[1])
        lbs = [letterbox_device(page, imgsz, swap_rb=True) for page in pages_bgr]
        g = yolo.forward_letterboxed_batch(lbs)
        det, cnt, final_idx = yolo.detect_batch(g, confidence, (h, w), tuple(lbs[0].shape[:2]), apply_reference_dedup=True)
        idx = final_idx.long().clamp_(0, det.shape[1] - 1)
        tables.copy_(torch.gather(det, 1, idx.unsqueeze(-1).expand(-1, -1, det.shape[2])))
        counts.copy_(cnt[:, 1])
    for pi, page in enumerate(pages_bgr if not batched else []):
        h, w = int(page.shape[0]), int(page.shape[1])
        lb = letterbox_device(page, imgsz, swap_rb=True)
        g = yolo.forward_letterboxed(lb)
        det, cnt, final_idx = yolo.detect(g, confidence, (h, w), tuple(lb.shape[:2]), apply_reference_dedup=True)
        tables[pi].copy_(det[final_idx.long().clamp_(0, det.shape[0] - 1)])       # rows beyond the count are ignored below
        counts[pi].copy_(cnt[1])
        if npages == 1 and sam is not None:
            # The SAM image encoder does not depend on the boxes: enqueue it before the host waits for the detector, so
            # the box table's round trip and the host-side grouping run under ~5 ms of encoder work instead of an idle GPU.
            enc0 = sam.encode(page[:, :, [2, 1, 0]].contiguous())
    if npages:
        counts_h = counts.cpu().numpy()                                            # the one host sync of the detect stage
        tables_h = tables.cpu().numpy()
    for pi, page in enumerate(pages_bgr):
        h, w = int(page.shape[0]), int(page.shape[1])
        enc = enc0 if npages == 1 else (sam.encode(page[:, :, [2, 1, 0]].contiguous()) if sam is not None else None)
        n_final = int(counts_h[pi])
        rows = tables_h[pi, :n_final] if n_final else np.zeros((0, 8), np.float32)
        boxes = rows[:, :4].astype(np.float32)
        confs = rows[:, 4].astype(np.float32)
        if injected_boxes is not None:
            boxes = np.asarray(injected_boxes[pi], np.float32).reshape(-1, 4)
            confs = np.full((boxes.shape[0],), 0.9, np.float32)
        tb = torch.from_numpy(boxes)
        sources = [("primary", i) for i in range(len(tb))]
        sec, sec_src, sec_conf, sec_names, sec_cls = torch.zeros((0, 4)), [], None, None, None
        if conjoined_detection and len(tb):
            try:
                tb, sources, sec, sec_src, sec_res, _free = _secondary_detections(mm, page, tb, sources, conjoined_confidence,
                                                                                 page.device, False, False)
                tb, sec = tb.detach().float().cpu(), sec.detach().float().cpu()
                sec_conf, sec_cls = sec_res.boxes.conf.detach().float().cpu(), sec_res.boxes.cls.detach().cpu()
                sec_names = getattr(sec_res, "names", None)
            except Exception as e:                                     # like the reference: keep the primaries (:1541-1548)
                log_message(f"Warning: Could not load/run secondary RT-DETR model: {e}. "
                            "Proceeding without conjoined/fallback detection.", verbose=False)
                sec, sec_src = torch.zeros((0, 4)), []

        def meta(src):
            if src[0] == "primary":
                return float(confs[src[1]]), "speech_bubble"
            cid = int(sec_cls[src[1]])
            return float(sec_conf[src[1]]), (sec_names.get(cid, "speech_bubble") if sec_names is not None else "speech_bubble")

        dets: List[Dict[str, Any]] = []
        if len(tb) == 0:
            out.append(dets)
            continue
        # grouping: conjoined parents from the secondary boxes, then synthetic groups of overlapping primaries
        # (:1596-1619, formed with or without a secondary detector); SAM is prompted with each group's union box
        conjoined, simple = [], list(range(len(tb)))
        if len(sec) > 0:
            conjoined, simple = categorize_detections(tb, sec, IOA_THRESHOLD)
        groups = []
        if len(simple) > 1:
            groups, simple = detect_overlapping_primaries(tb, simple)
            from mangatranslator_b200.conjoined import MAX_CHILDREN
            groups, simple = _ungroup_oversized(groups, simple, MAX_CHILDREN)
        group_boxes = [sec[s_idx] for _, s_idx in conjoined] + [tb[m] for m in groups]
        group_srcs = [[sec_src[s] for s in s_idx] for _, s_idx in conjoined] + [[sources[m] for m in ms] for ms in groups]
        parents = [union_box(torch.cat([tb[p].unsqueeze(0), sec[s_idx]], 0)) for p, s_idx in conjoined] + \
                  [union_box(tb[m]) for m in groups]
        prompts = [tb[i] for i in simple] + parents
        masks = None
        if sam is not None and prompts:
            masks = sam.decode(enc, torch.stack(prompts), (h, w))
            if own_masks:
                masks = masks.clone()

        def clip_rect(b):
            x0, y0, x1, y1 = [float(v) for v in b]
            bx0, by0 = int(np.floor(max(0, min(x0, w)))), int(np.floor(max(0, min(y0, h))))
            bx1, by1 = int(np.ceil(max(0, min(x1, w)))), int(np.ceil(max(0, min(y1, h))))
            return bx0, by0, bx1, by1

        def prompt_mask(n):
            if masks is not None:
                return masks[n]
            bx0, by0, bx1, by1 = clip_rect(prompts[n])
            m = torch.zeros((h, w), dtype=torch.uint8, device=page.device)
            m[by0:by1, bx0:bx1] = 255
            return m

        for n, k in enumerate(simple):
            bx0, by0, bx1, by1 = clip_rect(tb[k])
            conf, cls = meta(sources[k])
            dets.append({"bbox": tuple(int(round(float(v))) for v in tb[k]), "confidence": conf, "class": cls,
                         "sam_mask": prompt_mask(n), "mask_bbox": (bx0, by0, max(bx1, bx0 + 1), max(by1, by0 + 1))})
        for gi, (gb, gs) in enumerate(zip(group_boxes, group_srcs)):
            px0, py0, px1, py1 = clip_rect(prompts[len(simple) + gi])
            window = (px0, py0, max(px1, px0 + 1), max(py1, py0 + 1))
            child_masks, plan = split_conjoined_device(prompt_mask(len(simple) + gi), gb, window=window)
            for n, src in enumerate(gs):
                conf, cls = meta(src)
                dets.append({"bbox": plan.bboxes[n], "confidence": conf, "class": cls, "sam_mask": child_masks[n],
                             "mask_bbox": window,
                             "conjoined_neighbor_bboxes": [b for m, b in enumerate(plan.bboxes) if m != n]})
        out.append(dets)
    return out
