"""TEST INFRASTRUCTURE ONLY — CPU fp32 oracle of the reference's secondary (conjoined / fallback) bubble detector:
the real `transformers.RTDetrV2ForObjectDetection` + `RTDetrImageProcessor`, driven exactly like the reference's
`RTDetrYOLOAdapter.__call__` does (core/ml/rtdetr_adapter.py:61-113; built at core/ml/model_manager.py:745-778, called at
core/image/detection.py:1401-1407 with conf = conjoined_confidence, imgsz = 640).

The arithmetic lives in the third-party library (reference pins transformers >= 5.0.0, requirements.txt:19; this image
has 5.5.0), which is present, so the oracle IS the library.  WEIGHTS UNPINNED: the checkpoint
(ogkalu/comic-text-and-bubble-detector) is not available offline; `make_model` seeds a random initialisation, gives the
frozen batch norms non-trivial statistics and lifts the class biases so that a few dozen queries pass conf = 0.35
(mangatranslator_b200.weights.rtdetr_model_and_state, shared with the CUDA path).
"""
from __future__ import annotations

import numpy as np
import torch

NAMES = {0: "bubble", 1: "text_bubble", 2: "text_free"}


def make_model(seed: int = 0, **overrides):
    """The library model carrying the seeded synthetic weights the CUDA path uses (mangatranslator_b200.weights)."""
    from transformers import RTDetrV2ForObjectDetection
    from mangatranslator_b200 import weights as W
    cfg, sd = W.rtdetr_model_and_state(seed, **overrides)
    model = RTDetrV2ForObjectDetection(cfg).eval()
    model.load_state_dict(sd)
    return cfg, model


def make_processor():
    from transformers import RTDetrImageProcessor
    return RTDetrImageProcessor()


@torch.no_grad()
def predict(model, processor, rgb_u8: np.ndarray, conf: float = 0.35, imgsz: int = 640):
    """The adapter's call: processor(images, size) -> model -> post_process_object_detection(threshold, target size)."""
    from PIL import Image
    h, w = rgb_u8.shape[:2]
    inputs = processor(images=Image.fromarray(rgb_u8), return_tensors="pt", size={"height": imgsz, "width": imgsz})
    out = model(**inputs)
    res = processor.post_process_object_detection(out, threshold=float(conf), target_sizes=[(h, w)],
                                                  use_focal_loss=bool(getattr(model.config, "use_focal_loss", True)))[0]
    return dict(xyxy=res["boxes"].float(), conf=res["scores"].float(), cls=res["labels"].float(), pixel_values=inputs["pixel_values"],
                logits=out.logits[0], pred_boxes=out.pred_boxes[0], enc_cls=out.enc_outputs_class[0],
                enc_box=out.enc_outputs_coord_logits[0], pan=[t[0] for t in out.encoder_last_hidden_state],
                init_ref=out.init_reference_points[0])
