"""Duck-typed `(processor, model)` pair for SAM 2.1 with the call shapes the reference uses on the transformers objects
(core/image/detection.py:494-511):

    inputs = processor(image_pil, input_boxes=boxes_cpu (1,P,4), return_tensors="pt")
    for k in inputs: inputs[k] = inputs[k].to(model.dtype) ...; inputs = inputs.to(device)
    outputs = model(multimask_output=False, **inputs)
    masks = processor.post_process_masks(outputs.pred_masks, inputs["original_sizes"])[0][:, 0]   # bool (P,H,W)

The processor only records the page and boxes (the resize/normalise runs on the device inside the encoder); the model
runs encoder + decoder and keeps the low-res logits of the selected mask; post_process_masks runs the fused
bilinear-upsample + threshold kernel.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class _Inputs(dict):
    """Mapping that also supports `.to(device)` like transformers' BatchFeature."""

    def to(self, *args, **kwargs):
        for k, v in list(self.items()):
            if isinstance(v, torch.Tensor):
                self[k] = v.to(*args, **kwargs)
        return self


class Sam2ProcessorB200:
    def __init__(self, net):
        self.net = net

    def __call__(self, images, input_boxes=None, return_tensors="pt", **_):
        arr = np.asarray(images.convert("RGB") if hasattr(images, "convert") else images)
        h, w = arr.shape[:2]
        img = torch.from_numpy(np.ascontiguousarray(arr))
        boxes = torch.as_tensor(input_boxes, dtype=torch.float32)
        if boxes.dim() == 2:
            boxes = boxes.unsqueeze(0)
        return _Inputs(pixel_values=img, original_sizes=torch.tensor([[h, w]], dtype=torch.int64), input_boxes=boxes)

    def post_process_masks(self, pred_masks, original_sizes, **_):
        """-> [bool tensor (P, 1, H, W)] like Sam2Processor.post_process_masks (binarised at 0)."""
        h, w = [int(v) for v in original_sizes[0]]
        full = self.net._last_full_masks(h, w)      # uint8 (P,H,W) from bilinear(>0), no box clip
        return [(full > 0).unsqueeze(1)]


class Sam2ModelB200:
    def __init__(self, net):
        self.net = net
        self.dtype = torch.float32          # the reference casts its float inputs to model.dtype (detection.py:498-500)
        self.device = net.device

    def __call__(self, multimask_output=False, pixel_values=None, original_sizes=None, input_boxes=None, **_):
        img = pixel_values.to(self.net.device)
        if img.dtype != torch.uint8:        # the reference casts float tensors; the page itself stays uint8
            img = img.to(torch.uint8)
        enc = self.net.encode(img.contiguous())
        boxes = input_boxes[0].to(torch.float32)
        h, w = img.shape[:2]
        pred = self.net.decode_lowres(enc, boxes, (h, w))
        return SimpleNamespace(pred_masks=pred.unsqueeze(0).unsqueeze(2))     # (1, P, 1, 256, 256)
