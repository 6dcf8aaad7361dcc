"""TEST INFRASTRUCTURE ONLY — the reference's CPU pipeline for the hot path, stage by stage, on the host cores:
YOLOv8-seg predict (oracle/yolo_oracle.py) -> SAM 2.1 box-prompted masks through the real transformers classes
(oracle/sam2_oracle.py) -> OpenCV bubble cleaning (oracle/clean_oracle.py) -> RCAN 2x upscale (oracle/rcan_oracle.py),
with the same seeded weights the CUDA path uses (mangatranslator_b200.weights).  Used by bench.py's cpu_baseline /
`--impl reference` arm and by the end-to-end parity test.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch
from PIL import Image

import clean_oracle
import rcan_oracle
import sam2_oracle
import yolo_oracle


class CpuPipeline:
    def __init__(self, seed: int = 0, yolo_variant: str = "m", rcan_groups: int = 10, rcan_blocks: int = 20):
        from mangatranslator_b200 import weights as W
        self.ycfg = W.yolo_cfg(yolo_variant)
        self.yolo = yolo_oracle.YoloV8Seg(**self.ycfg).eval()
        self.yolo.load_state_dict(W.yolo_state_dict(seed, self.ycfg))
        cfg, sd = W.sam2_model_and_state(seed)
        from transformers import Sam2Model
        self.sam = Sam2Model(cfg).eval()
        self.sam.load_state_dict(sd)
        self.proc = sam2_oracle.make_processor()
        self.rcan = rcan_oracle.RCAN(n_resgroups=rcan_groups, n_resblocks=rcan_blocks).eval()
        self.rcan.load_state_dict(W.rcan_state_dict(seed, n_resgroups=rcan_groups, n_resblocks=rcan_blocks))

    @torch.no_grad()
    def run_page(self, rgb: np.ndarray, boxes: Optional[np.ndarray], *, conf: float = 0.6, imgsz: int = 1600,
                 upscale_crop: Optional[int] = None) -> Dict[str, float]:
        """One page through all four stages; returns per-stage seconds (and the outputs).  `upscale_crop` runs the
        RCAN on a centre crop of that size and scales its time by the pixel ratio (the full frame takes minutes)."""
        h, w = rgb.shape[:2]
        bgr = np.ascontiguousarray(rgb[:, :, ::-1])
        t = {}
        t0 = time.perf_counter()
        det = yolo_oracle.predict(self.yolo, bgr, conf, imgsz)
        t["detect"] = time.perf_counter() - t0
        if boxes is None:
            boxes = det["xyxy"].numpy()
        t0 = time.perf_counter()
        # the reference groups overlapping primaries into synthetic conjoined bubbles (detection.py:1596-1619): SAM is
        # prompted with the simple boxes and each group's union box, the parent masks are split between the members
        import conjoined_oracle
        groups, simple = conjoined_oracle.overlapping_groups(boxes) if len(boxes) > 1 else ([], list(range(len(boxes))))
        prompts = [boxes[i] for i in simple] + [conjoined_oracle.union_box([boxes[i] for i in g]) for g in groups]
        seg = sam2_oracle.segment(self.sam, self.proc, Image.fromarray(rgb), np.asarray(prompts, np.float32)) if len(prompts) \
            else dict(masks=[])
        dets = [{"bbox": tuple(int(round(float(v))) for v in boxes[i]), "sam_mask": seg["masks"][n]} for n, i in enumerate(simple)]
        for gi, g in enumerate(groups):
            masks, bboxes = conjoined_oracle.split_group(seg["masks"][len(simple) + gi], [boxes[i] for i in g])
            for n in range(len(g)):
                dets.append({"bbox": bboxes[n], "sam_mask": masks[n],
                             "conjoined_neighbor_bboxes": [b for m, b in enumerate(bboxes) if m != n]})
        t["segment"] = time.perf_counter() - t0
        # pixels whose full-size logit is within 2e-3 of zero for ANY prompt: the only places where an fp32-grade segmenter
        # may legitimately decide a mask bit differently (a split child inherits its parent's knife-edge pixels)
        band = (seg["full_logits"].abs() < 2e-3).any(0).numpy() if len(prompts) else np.zeros((h, w), bool)
        seg = dict(seg, masks=np.stack([d["sam_mask"] for d in dets]) if dets else np.zeros((0, h, w), np.uint8))
        t0 = time.perf_counter()
        cleaned, bubbles = clean_oracle.clean_page(bgr, dets, processing_scale=(h * w / 1e6) ** 0.5)
        t["clean"] = time.perf_counter() - t0
        src = np.ascontiguousarray(cleaned[:, :, ::-1])
        scale = 1.0
        if upscale_crop is not None and upscale_crop < min(h, w):
            y0, x0 = (h - upscale_crop) // 2, (w - upscale_crop) // 2
            src = np.ascontiguousarray(src[y0:y0 + upscale_crop, x0:x0 + upscale_crop])
            scale = (h * w) / float(upscale_crop * upscale_crop)
        t0 = time.perf_counter()
        out_f, out_u8 = rcan_oracle.upscale_u8(self.rcan, src)
        t["upscale"] = (time.perf_counter() - t0) * scale
        t["total"] = t["detect"] + t["segment"] + t["clean"] + t["upscale"]
        return dict(times=t, masks=seg["masks"], band=band, cleaned=cleaned, upscaled=out_u8, upscaled_f=out_f,
                    bubbles=bubbles, det=det)
