"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle of the SAM 2.1 segmentation call.

The arithmetic lives in a third-party package that IS installed here and on the GPU box: `transformers` (reference pins
>=5.0.0, requirements.txt:19; this image has 5.5.0).  The oracle therefore runs the real `Sam2Model` / `Sam2Processor`
exactly the way the reference does (core/ml/model_manager.py:996-1005, core/image/detection.py:475-511, :1732-1750):
processor(image, input_boxes) -> model(multimask_output=False) -> post_process_masks -> [:, 0] -> > 0.5 -> AND with
the floor/ceil box rectangle -> uint8 {0,255}.  Weights are seeded random (no checkpoint offline), dtype fp32 (the
reference's CPU dtype; on CUDA it would run bf16).  PARITY UNPINNED by reference tests (the reference has none).
"""
from __future__ import annotations

import numpy as np
import torch


def make_model(seed: int = 0, config=None, spread: float = 1.0):
    from transformers import Sam2Config, Sam2Model
    torch.manual_seed(seed)
    cfg = config or Sam2Config()
    m = Sam2Model(cfg).eval()
    if spread != 1.0:
        # the default init (std 0.02) gives near-zero mask logits; widen the hyper-network output so masks are non-trivial
        with torch.no_grad():
            for mlp in m.mask_decoder.output_hypernetworks_mlps:
                mlp.proj_out.weight.mul_(spread)
    return m


def make_processor():
    from transformers import Sam2ImageProcessorFast, Sam2Processor
    return Sam2Processor(Sam2ImageProcessorFast())


@torch.no_grad()
def segment(model, processor, pil_image, boxes_xyxy: np.ndarray):
    """Reference flow.  Returns dict(masks uint8 [P,H,W] {0,255}, pred_masks [P,256,256] (selected low-res logits),
    full_logits [P,H,W] (interpolated), image_embeddings list)."""
    boxes = torch.as_tensor(boxes_xyxy, dtype=torch.float32).unsqueeze(0)
    inputs = processor(pil_image, input_boxes=boxes, return_tensors="pt")
    for k in inputs:
        if isinstance(inputs[k], torch.Tensor) and inputs[k].is_floating_point():
            inputs[k] = inputs[k].to(model.dtype)
    out = model(multimask_output=False, **inputs)
    h, w = [int(v) for v in inputs["original_sizes"][0]]
    full = torch.nn.functional.interpolate(out.pred_masks[0], (h, w), mode="bilinear", align_corners=False)[:, 0]
    masks_t = processor.post_process_masks(out.pred_masks, inputs["original_sizes"])[0][:, 0]
    m = (masks_t > 0.5).cpu().numpy()
    res = []
    for mk, b in zip(m, boxes_xyxy):
        x0 = int(np.floor(max(0, min(b[0], w))))
        y0 = int(np.floor(max(0, min(b[1], h))))
        x1 = int(np.ceil(max(0, min(b[2], w))))
        y1 = int(np.ceil(max(0, min(b[3], h))))
        clip = np.zeros_like(mk)
        if x1 > x0 and y1 > y0:
            clip[y0:y1, x0:x1] = True
        res.append(np.where(mk & clip, 255, 0).astype(np.uint8))
    return dict(masks=np.stack(res), pred_masks=out.pred_masks[0, :, 0], full_logits=full, iou=out.iou_scores[0],
                image_embeddings=out.image_embeddings, pixel_values=inputs["pixel_values"])
