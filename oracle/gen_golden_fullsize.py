"""TEST INFRASTRUCTURE ONLY — golden vectors at the BENCHMARKED configuration (BASELINE.json configs[2]): one seeded
1536x1024 page through the CPU oracles with the bench's seeded weights (mangatranslator_b200.weights, seed 0):

  detect   YOLOv8m-seg @ imgsz 1600 (oracle/yolo_oracle.py): class logits of every anchor, box-distribution / mask-coefficient
           logits on a fixed anchor grid, prototypes on a fixed pixel grid, float64 checksums of the full tensors, and the NMS
           result (anchor ids, boxes, scores) at a confidence threshold placed in the widest score gap (margins recorded)
  segment  SAM 2.1-tiny through the real transformers classes (oracle/sam2_oracle.py) for the page's prompts (simple boxes +
           union boxes of the synthetic conjoined groups): final uint8 masks (bit-packed), the band of pixels whose full-size
           logit is within 2e-3 of zero (bit-packed), low-res logits on a grid, IoU scores
  clean    oracle/clean_oracle.py on those masks: sha256 of the cleaned page, per-bubble records
  upscale  RCAN 10x20 (oracle/rcan_oracle.py) on the cleaned page, whole frame, fp32: float output on a fixed grid and in
           four full 64x64 tiles (corners / centre), uint8 of the same tiles

The RCAN pass takes a few minutes on 8 cores, which is why it is generated once here and committed
(tests/golden/fullsize_golden.npz, ~1 MB) instead of being recomputed by the GPU test.

    python oracle/gen_golden_fullsize.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import clean_oracle  # noqa: E402
import conjoined_oracle  # noqa: E402
import pipeline_oracle  # noqa: E402
import rcan_oracle  # noqa: E402
import sam2_oracle  # noqa: E402
import yolo_oracle as Y  # noqa: E402
from mangatranslator_b200 import synth  # noqa: E402

H, W, SEED, IMGSZ = 1536, 1024, 9000, 1600
ANCHOR_STEP, PROTO_STEP, LOWRES_STEP, UP_STEP, TILE = 16, 8, 4, 16, 64


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out, meta = {}, dict(H=H, W=W, seed=SEED, imgsz=IMGSZ, anchor_step=ANCHOR_STEP, proto_step=PROTO_STEP,
                         lowres_step=LOWRES_STEP, up_step=UP_STEP, tile=TILE)
    pg = synth.make_page(SEED, H, W, n_bubbles=12)
    rgb = pg.image_rgb
    bgr = np.ascontiguousarray(rgb[:, :, ::-1])
    meta["page_sha256"] = sha(bgr)
    pipe = pipeline_oracle.CpuPipeline(0)

    # ---- detect ----
    t0 = time.time()
    x = Y.preprocess(bgr, IMGSZ)
    with torch.no_grad():
        raw, proto = pipe.yolo.heads_raw(x)
        pred, _ = pipe.yolo(x)
    meta["letterbox_hw"] = [int(x.shape[2]), int(x.shape[3])]
    for i, (box, cls, mc) in enumerate(raw):
        a = box.shape[2] * box.shape[3]
        out[f"yolo_cls_{i}"] = cls[0].reshape(-1, a).numpy().astype(np.float32)                    # [nc][A]
        out[f"yolo_box_{i}"] = box[0].reshape(64, a)[:, ::ANCHOR_STEP].numpy().astype(np.float32)    # [64][A/step]
        out[f"yolo_mc_{i}"] = mc[0].reshape(-1, a)[:, ::ANCHOR_STEP].numpy().astype(np.float32)
        meta[f"yolo_sum_{i}"] = [float(box.double().sum()), float(box.double().abs().sum()), float(mc.double().sum()),
                                 float(mc.double().abs().sum())]
    out["yolo_proto"] = proto[0, :, ::PROTO_STEP, ::PROTO_STEP].numpy().astype(np.float32)
    meta["yolo_proto_sum"] = [float(proto.double().sum()), float(proto.double().abs().sum())]
    # confidence threshold in the widest gap among the 20th..60th best scores; the oracle's own decision margins
    sc = torch.sort(pred[0, 4], descending=True).values
    gaps = sc[19:59] - sc[20:60]
    k = 20 + int(torch.argmax(gaps))
    conf = float((sc[k - 1] + sc[k]) / 2)
    det = Y.predict(pipe.yolo, bgr, conf, IMGSZ)
    p = pred[0].t()
    b = p[p[:, 4] > conf]
    xyxy = torch.cat((b[:, :2] - b[:, 2:4] / 2, b[:, :2] + b[:, 2:4] / 2), 1)
    iou = Y.box_iou_matrix(xyxy)
    dm = (iou - 0.7).abs()
    dm.fill_diagonal_(1.0)
    top = sc[:k]
    meta["yolo_conf"] = conf
    meta["yolo_margins"] = dict(cut_gap=float(gaps.max()), min_score_gap=float((top[:-1] - top[1:]).min()),
                                min_iou_margin=float(dm.min()), candidates=int(k))
    out["yolo_det_anchors"] = det["anchors"].numpy().astype(np.int64)
    out["yolo_det_xyxy"] = det["xyxy"].numpy().astype(np.float32)
    out["yolo_det_conf"] = det["conf"].numpy().astype(np.float32)
    out["yolo_det_mask_pixels"] = det["masks"].reshape(det["masks"].shape[0], -1).sum(1).numpy().astype(np.int64)
    print(f"detect {time.time() - t0:.1f}s: conf {conf:.6f}, {len(det['anchors'])} detections of {k} candidates, margins "
          f"{meta['yolo_margins']}", flush=True)

    # ---- segment (prompts exactly as the device path forms them: simple boxes, then the union box of each group) ----
    t0 = time.time()
    boxes = pg.boxes_xyxy
    groups, simple = conjoined_oracle.overlapping_groups(boxes)
    prompts = np.asarray([boxes[i] for i in simple] + [conjoined_oracle.union_box([boxes[i] for i in g]) for g in groups],
                         np.float32)
    seg = sam2_oracle.segment(pipe.sam, pipe.proc, Image.fromarray(rgb), prompts)
    meta["sam_simple"], meta["sam_groups"] = [int(i) for i in simple], [[int(i) for i in g] for g in groups]
    out["sam_prompts"] = prompts
    out["sam_masks_bits"] = np.packbits(seg["masks"] > 0, axis=-1)
    band = (seg["full_logits"].abs() < 2e-3).numpy()
    out["sam_band_bits"] = np.packbits(band, axis=-1)
    meta["sam_band_pixels"] = int(band.sum())
    out["sam_lowres"] = seg["pred_masks"][:, ::LOWRES_STEP, ::LOWRES_STEP].numpy().astype(np.float32)
    out["sam_iou"] = seg["iou"].reshape(len(prompts), -1).numpy().astype(np.float32)
    meta["sam_mask_pixels"] = [int((m > 0).sum()) for m in seg["masks"]]
    print(f"segment {time.time() - t0:.1f}s: {len(prompts)} prompts, band pixels {meta['sam_band_pixels']}", flush=True)

    # ---- detections as the pipeline assembles them (group masks split between the members), then clean ----
    t0 = time.time()
    dets = [{"bbox": tuple(int(round(float(v))) for v in boxes[i]), "sam_mask": seg["masks"][n]} for n, i in enumerate(simple)]
    for gi, g in enumerate(groups):
        masks, bboxes = conjoined_oracle.split_group(seg["masks"][len(simple) + gi], [boxes[i] for i in g])
        for n in range(len(g)):
            dets.append({"bbox": bboxes[n], "sam_mask": masks[n],
                         "conjoined_neighbor_bboxes": [bb for m, bb in enumerate(bboxes) if m != n]})
    out["det_masks_bits"] = np.packbits(np.stack([d["sam_mask"] for d in dets]) > 0, axis=-1)
    meta["det_bboxes"] = [[int(v) for v in d["bbox"]] for d in dets]
    meta["det_neighbors"] = [[[int(v) for v in bb] for bb in d.get("conjoined_neighbor_bboxes", [])] for d in dets]
    cleaned, bubbles = clean_oracle.clean_page(bgr, dets, processing_scale=(H * W / 1e6) ** 0.5)
    meta["cleaned_sha256"] = sha(cleaned)
    meta["bubbles"] = [dict(color=[int(v) for v in b["color"]], text_bbox=[int(v) for v in b["text_bbox"]],
                            mask_pixels=int((b["mask"] > 0).sum()), mask_sha256=sha(b["mask"])) for b in bubbles]
    print(f"clean {time.time() - t0:.1f}s: {len(bubbles)} bubbles, sha {meta['cleaned_sha256'][:16]}", flush=True)

    # ---- upscale: whole frame, fp32, the full-depth network ----
    t0 = time.time()
    src = np.ascontiguousarray(cleaned[:, :, ::-1])
    y, u8 = rcan_oracle.upscale_u8(pipe.rcan, src)
    y = y[0].numpy()                                                        # [3][2H][2W]
    out["up_grid"] = y[:, ::UP_STEP, ::UP_STEP].astype(np.float32)
    tiles = [(0, 0), (0, 2 * W - TILE), (2 * H - TILE, 0), (2 * H - TILE, 2 * W - TILE), (H - TILE // 2, W - TILE // 2),
             (1234, 777)]
    meta["up_tiles"] = [[int(a), int(b)] for a, b in tiles]
    out["up_tiles_f"] = np.stack([y[:, a:a + TILE, b:b + TILE] for a, b in tiles]).astype(np.float32)
    out["up_tiles_u8"] = np.stack([u8[a:a + TILE, b:b + TILE] for a, b in tiles])
    meta["up_range"] = [float(y.min()), float(y.max())]
    meta["up_clipped_frac"] = float(((y < 0) | (y > 1)).mean())
    print(f"upscale {time.time() - t0:.1f}s: range {meta['up_range']}, clipped {meta['up_clipped_frac']:.4f}", flush=True)

    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "fullsize_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
