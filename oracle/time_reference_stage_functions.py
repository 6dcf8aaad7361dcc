"""TEST INFRASTRUCTURE ONLY (build container: needs /root/reference) — how long do the reference's OWN stage functions take
on the CPU for the benchmarked page, next to the oracle port that bench.py's CPU arm times on the GPU box (where the
reference checkout does not exist)?

The unmodified `detect_speech_bubbles` (core/image/detection.py:1263), `clean_speech_bubbles` (core/image/cleaning.py:524) and
`upscale_image` (core/image/image_utils.py:503) are driven exactly as BASELINE.md section 4.2 planned: the third-party models
they would load are injected through `ModelManager.models` — the restated YOLOv8m-seg (oracle/yolo_oracle.py) behind an
ultralytics-shaped callable, the real transformers Sam2Model / Sam2Processor, the restated RCAN — with the bench's seeded
weights.  Writes profiles/r02_reference_stage_functions_cpu.json.

    python oracle/time_reference_stage_functions.py
"""
import json
import os
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _refimport  # noqa: E402
import pipeline_oracle  # noqa: E402
import yolo_oracle  # noqa: E402
from mangatranslator_b200 import synth  # noqa: E402

H, W, CROP = 1536, 1024, 256


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    _refimport.import_reference()
    import core.image.cleaning as ref_clean
    import core.image.detection as ref_det
    import core.image.image_utils as ref_iu
    from core.caching import get_cache
    from core.ml.model_manager import ModelType, get_model_manager
    pipe = pipeline_oracle.CpuPipeline(0)
    pg = synth.make_page(9000, H, W, n_bubbles=12)
    pil = Image.fromarray(pg.image_rgb)
    gt = torch.from_numpy(np.asarray(pg.boxes_xyxy, np.float32))

    class Boxes:
        def __init__(self):
            self.xyxy, self.conf, self.cls = gt.clone(), torch.full((len(gt),), 0.9), torch.zeros(len(gt))

        def __len__(self):
            return len(gt)

    class Detector:
        """ultralytics call shape; runs the full network (that is what is timed) and reports the page's ground-truth boxes,
        like the bench does for the stages downstream of the detector."""
        names = {0: "speech_bubble"}

        def __call__(self, im, conf=0.6, device=None, verbose=False, imgsz=1600, retina_masks=True):
            yolo_oracle.predict(pipe.yolo, im, conf, imgsz)
            return [SimpleNamespace(boxes=Boxes(), masks=None, orig_shape=im.shape[:2], names=self.names)]

    mm = get_model_manager()
    mm.device = torch.device("cpu")
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = Detector()
    mm.models[ModelType.SAM2] = (pipe.proc, pipe.sam)
    mm.models[ModelType.UPSCALE] = pipe.rcan
    scale = (H * W / 1e6) ** 0.5
    out = {"what": "the reference's own stage functions on this container's CPU (8 cores) vs the oracle port on the same page: "
                   "seconds per 1536x1024 page, RCAN on a 256x256 crop in both", "cores": os.cpu_count(), "runs": []}
    for rep in range(3):
        get_cache().clear_all() if hasattr(get_cache(), "clear_all") else None
        t0 = time.perf_counter()
        dets, _ = ref_det.detect_speech_bubbles(Path("page.png"), "x.pt", 0.6, verbose=False, device=torch.device("cpu"),
                                                seg_model="sam2", conjoined_detection=False, image_override=pil)
        t_det = time.perf_counter() - t0
        t0 = time.perf_counter()
        cleaned, info = ref_clean.clean_speech_bubbles(pil, "x.pt", 0.6, pre_computed_detections=dets, device=torch.device("cpu"),
                                                       processing_scale=scale)
        t_clean = time.perf_counter() - t0
        crop = Image.fromarray(np.ascontiguousarray(cleaned[:, :, ::-1])).crop(((W - CROP) // 2, (H - CROP) // 2,
                                                                                 (W + CROP) // 2, (H + CROP) // 2))
        get_cache().clear_all() if hasattr(get_cache(), "clear_all") else None
        t0 = time.perf_counter()
        up = ref_iu.upscale_image(crop, 2.0, model_type="model")
        t_up = time.perf_counter() - t0
        assert up.size == (2 * CROP, 2 * CROP)
        r = pipe.run_page(pg.image_rgb, pg.boxes_xyxy, upscale_crop=CROP)["times"]
        factor = H * W / float(CROP * CROP)
        out["runs"].append(dict(reference=dict(detect_and_segment=round(t_det, 3), clean=round(t_clean, 3), upscale_crop=round(t_up, 3)),
                                port=dict(detect_and_segment=round(r["detect"] + r["segment"], 3), clean=round(r["clean"], 3),
                                          upscale_crop=round(r["upscale"] / factor, 3)),
                                detections=len(dets), cleaned_bubbles=len(info)))
        print(out["runs"][-1], flush=True)
    ref = np.median([[v["reference"][k] for k in ("detect_and_segment", "clean", "upscale_crop")] for v in out["runs"]], 0)
    port = np.median([[v["port"][k] for k in ("detect_and_segment", "clean", "upscale_crop")] for v in out["runs"]], 0)
    out["median_seconds"] = dict(reference=dict(zip(("detect_and_segment", "clean", "upscale_crop"), map(float, ref))),
                                 port=dict(zip(("detect_and_segment", "clean", "upscale_crop"), map(float, port))))
    fr = H * W / float(CROP * CROP)
    out["page_seconds_with_rcan_scaled"] = dict(reference=float(ref[0] + ref[1] + ref[2] * fr), port=float(port[0] + port[1] + port[2] * fr))
    out["reading"] = ("the port that bench.py times on the GPU box costs what the reference's own functions cost on the same CPU, stage by "
                      "stage; the reference adds hashing / PIL conversions / a PNG round trip in upscale_image around the same arithmetic")
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_reference_stage_functions_cpu.json"), "w"), indent=1)
    print(json.dumps(out["median_seconds"]), out["page_seconds_with_rcan_scaled"])


if __name__ == "__main__":
    main()
