#!/usr/bin/env python
"""Headline benchmark: manga pages/sec through detect -> segment -> clean -> upscale on N x B200, next to the
reference's CPU pipeline on the host cores (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle), rank 0 only

One step = one batch of 64 synthetic 1536x1024 pages through the full hot path (BASELINE.json configs[2]).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("MTB200_SYNTHETIC_WEIGHTS", "1")     # `data: synthetic`: seeded weights of the named architectures
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch

METRIC = "manga pages/sec (detect->segment->clean->upscale)"
WORKLOAD = ("Batch 64 synthetic 1536x1024 pages, full detect->segment->clean->2x upscale (BASELINE.json configs[2]): "
            "YOLOv8m-seg@1600 + NMS/dedup, SAM2.1-tiny (12 box prompts), bit-exact bubble clean, RCAN(10x20,64) 2x")
H, W, BUBBLES = 1536, 1024, 12


_REAL_STDOUT_FD = None


def _stdout_carries_only_the_json_line() -> None:
    """File descriptor 1 is pointed at stderr for the rest of the process and the real stdout kept aside: whatever Python
    code (weight-source notices, the page driver's progress) or native libraries (NCCL prints its version banner to fd 1)
    write goes to stderr, and `emit` alone writes to stdout."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict) -> None:
    """The ONE JSON line of this program, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT_FD, data)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def start(self):
        def loop():
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([c.strip() for c in o.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(self.rows))


def make_pages(n_distinct: int, seed0: int):
    from mangatranslator_b200 import synth
    pages = [synth.make_page(seed0 + i, H, W, n_bubbles=BUBBLES) for i in range(n_distinct)]
    return pages


_CPU_PIPE = {}
RCAN_CROP = 256            # fixed, host-independent: the CPU arm's RCAN sample (see cpu_baseline)


def _cpu_pipe():
    """The CPU pipeline (oracle modules with the bench's seeded weights) is built once per process."""
    if "pipe" not in _CPU_PIPE:
        import pipeline_oracle
        pipe = pipeline_oracle.CpuPipeline(0)
        _CPU_PIPE["pipe"] = pipe
        _CPU_PIPE["threads"] = _best_threads(pipe)
    torch.set_num_threads(_CPU_PIPE["threads"])
    return _CPU_PIPE["pipe"], _CPU_PIPE["threads"]


def _best_threads(pipe) -> int:
    """torch's CPU convolutions scale badly past a few dozen threads on big hosts: pick the fastest of a few settings on a
    short RCAN probe (a few seconds, once per process)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    x = torch.rand(1, 3, RCAN_CROP, RCAN_CROP)
    best, best_t = cands[-1], 1e9
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            pipe.rcan.head(x)
            t0 = time.perf_counter()
            y = pipe.rcan.head(x)
            for blk in list(pipe.rcan.body[0].body)[:4]:
                y = blk(y)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(sample_pages: int = 1, page_seed: int = 9000):
    """The reference's CPU pipeline (oracle port, same seeded weights) on a bounded sample of the workload:
    `sample_pages` pages through detect / segment / clean IN FULL at 1536x1024, and the RCAN on a fixed 256x256 centre crop
    of the cleaned page.  The RCAN's cost is proportional to the pixel count (every layer is a 3x3 conv at input
    resolution; measured here: 215-254 s for the whole frame on 8 cores = 24.6x the crop), so its page time is the crop
    time x (H*W / 256^2) — stated as `extrapolated` with the factor; everything else is measured."""
    from mangatranslator_b200 import synth
    cores = os.cpu_count() or 1
    pipe, threads = _cpu_pipe()
    tot, wall, stages = 0.0, 0.0, {}
    for i in range(sample_pages):
        pg = synth.make_page(page_seed + i, H, W, n_bubbles=BUBBLES)
        t0 = time.perf_counter()
        r = pipe.run_page(pg.image_rgb, pg.boxes_xyxy, upscale_crop=RCAN_CROP)
        wall += time.perf_counter() - t0
        tot += r["times"]["total"]
        for k, v in r["times"].items():
            stages[k] = stages.get(k, 0.0) + v / sample_pages
    factor = H * W / float(RCAN_CROP * RCAN_CROP)
    return dict(value=sample_pages / tot, unit="pages/s", cores=threads, kind="port",
                sample=f"{sample_pages} page(s) 1536x1024: YOLOv8m-seg@1600 + SAM2.1-tiny (12 boxes) + cv2 clean measured in full; "
                       f"RCAN(10x20,64) measured on a {RCAN_CROP}x{RCAN_CROP} crop and scaled by the pixel ratio {factor:.1f}; "
                       f"{threads} torch threads on a {cores}-core host",
                extrapolated=True, extrapolation=dict(stage="upscale", factor=round(factor, 2), measured_s=round(stages["upscale"] / factor, 3),
                                                      why="whole-frame 10x20 RCAN takes minutes per page on the host cores"),
                sample_wall_s=round(wall / sample_pages, 3),
                stage_seconds={k: round(v, 3) for k, v in stages.items()})


def gpu_baseline(dev, reps: int = 3):
    """Secondary comparator: the SAME torch modules the CPU arm runs (oracle/*), moved to the
    GPU with torch's defaults — cuDNN TF32 convolutions for the RCAN and YOLO (core/ml/model_manager.py:640-654 puts the
    fp32 module on the device and calls it), bf16 weights for SAM (model_manager.py:1001).  ms per 1536x1024 page and
    per stage; not the product path and not a parity reference (TF32 misses the 1e-3 bound)."""
    import pipeline_oracle
    import yolo_oracle
    from PIL import Image
    from mangatranslator_b200 import synth
    out = {}
    pipe = pipeline_oracle.CpuPipeline(0)
    pg = synth.make_page(9000, H, W, n_bubbles=BUBBLES)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(pg.image_rgb)).to(dev).permute(2, 0, 1).float().div(255.0).unsqueeze(0)
        rcan = pipe.rcan.to(dev)
        rcan.forward = _rcan_forward_on(rcan, dev)
        out["upscale_tf32_ms"] = round(timed(lambda: rcan(x)), 2)
        rcl = rcan.to(memory_format=torch.channels_last)
        xcl = x.contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        out["upscale_tf32_channels_last_autotuned_ms"] = round(timed(lambda: rcl(xcl)), 2)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out["upscale_bf16_autocast_ms"] = round(timed(lambda: rcl(xcl)), 2)
        torch.backends.cudnn.benchmark = False
        del rcan, rcl
        yolo = pipe.yolo.to(dev)
        xin = yolo_oracle.preprocess(np.ascontiguousarray(pg.image_rgb[:, :, ::-1]), 1600).to(dev)
        out["detect_network_tf32_ms"] = round(timed(lambda: yolo(xin)), 2)
        del yolo
        sam = pipe.sam.to(dev).to(torch.bfloat16)
        boxes = torch.as_tensor(pg.boxes_xyxy, dtype=torch.float32).unsqueeze(0)
        inputs = pipe.proc(Image.fromarray(pg.image_rgb), input_boxes=boxes, return_tensors="pt")
        inputs = {k: (v.to(dev).to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else
                      v.to(dev) if torch.is_tensor(v) else v) for k, v in inputs.items()}
        out["segment_network_bf16_ms"] = round(timed(lambda: sam(multimask_output=False, **inputs)), 2)
        del sam
    torch.cuda.empty_cache()
    tot = out["upscale_tf32_ms"] + out["detect_network_tf32_ms"] + out["segment_network_bf16_ms"]
    out["pages_per_s_networks_only"] = round(1000.0 / tot, 3)
    out["note"] = ("torch eager on the same B200, same seeded weights: RCAN/YOLO fp32 modules with cuDNN TF32 (torch default), "
                   "SAM 2.1 in bf16; networks only (no letterbox, NMS, mask post-processing, cleaning or copies)")
    return out


def _rcan_forward_on(rcan, dev):
    """The oracle builds its mean tensor on the CPU; bind a forward that keeps every operand on `dev`."""
    import torch.nn.functional as F

    def fwd(x):
        x = x * rcan.rgb_range
        h = rcan.head(x)
        y = rcan.body(h) + h
        return rcan.tail(y) / rcan.rgb_range
    return fwd


def run_reference(args, coord):
    """The reference arm: the CPU pipeline on the box's host cores, rank 0 only.  One step = one bounded sample (one page:
    detect / segment / clean in full, the RCAN on the fixed crop); `ms_per_step` is the WALL time of a step as it ran, and
    `value` the pages/s that follow from the per-page time with the RCAN scaled to the whole frame (`extrapolated`)."""
    if coord.rank != 0:
        return
    vals, walls = [], []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        b = cpu_baseline(sample_pages=1, page_seed=9000 + s)
        if s >= args.warmup:
            vals.append(b)
            walls.append(time.perf_counter() - t0)
    v = float(np.mean([b["value"] for b in vals]))
    base = vals[-1]
    base["value"] = v
    line = dict(metric=METRIC, value=v, unit="pages/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * float(np.mean(walls)), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", impl="reference", extrapolated=True,
                config=dict(workload=WORKLOAD, sample_per_step=base["sample"],
                            note="value = 1 / (measured detect + segment + clean seconds + crop RCAN seconds x pixel ratio); "
                                 "ms_per_step = wall time of one sample step, not of a 64-page batch"),
                cpu_baseline=base, e2e=dict(value=v, unit="pages/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def measure_dominant_kernel(pipe, page_dev, peaks):
    """CUDA-event duration of every RCAN body conv launch (the halo-tile tcgen05 kernel) over one page."""
    rcan = pipe.rcan
    rcan.upscale_u8(page_dev, swap_rb=True)
    torch.cuda.synchronize()
    steps = rcan.time_steps(page_dev)
    durs = [ms for k, ms in steps if k == "conv_body"]
    avg_ms = float(np.mean(durs))
    by_kind = {}
    for k, ms in steps:
        by_kind[k] = by_kind.get(k, 0.0) + ms
    flops = 2.0 * H * W * 64 * 64 * 9
    ach = flops / (avg_ms * 1e-3) / 1e12
    fp16c = getattr(rcan, "precision", "") == "fp16c"
    # tensor-pipe work actually issued, in units of the algorithmic FLOPs: bf16x3 = 2 M128 MMAs per (tap, k16) where half an
    # MMA would do = 4x; fp16c = 1 M128 fp16 MMA + half an e5m2 MMA slot (K = 32) = 3x
    issue = 3.0 if fp16c else 4.0
    kernel = "conv3x3_c64_fp16c_kernel<fp16 + e5m2 correction>" if fp16c else "conv3x3_c64_cm_kernel<bf16x3>"
    note = ("algorithmic FLOPs (2*MAC of the fp32-grade conv); the kernel issues one [W16_hi;W16_lo] x X16 kind::f16 MMA per (tap, 16 "
            "input channels) and one W8 x X8 kind::f8f6f4 MMA per (tap, 32 input channels): 54 MMA slots of 120 clk per 240-pixel "
            "tile" if fp16c else
            "algorithmic FLOPs (2*MAC of the fp32-grade conv); the channel-major bf16x3 kernel issues 4x that in bf16 MMAs "
            "([W_hi;W_lo] rows against the hi and the lo activation plane)")
    return dict(bound="tensor", achieved=ach, peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=ach / peaks["bf16_sustained"],
                traffic=None, kernel=kernel, launches_timed=len(durs), avg_ms=avg_ms,
                peak_source=peaks["source"] + " bf16_tflops_sustained",
                upscale_launch_ms_by_kind={k: round(v, 3) for k, v in by_kind.items()},
                conv1_avg_ms=float(np.mean(durs[0::2])), conv2_avg_ms=float(np.mean(durs[1::2])),
                issued=dict(tflops=issue * ach, frac=issue * ach / peaks["bf16_sustained"], factor=issue,
                            note="tensor-pipe slots the kernel occupies, expressed as bf16-rate FLOPs: `frac` above is bounded "
                                 f"at 1/{issue:g} by the formulation (fp32-grade result from 16/8-bit operands)"),
                note=note)


def _time_ms(fn, reps=5, warm=3):
    for _ in range(warm):                    # plans are built on the first call and graph-captured on the second
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_side_stages(page_dev, dev, peaks):
    """Stages of the path that are not part of the headline workload, timed on their own after the timed region (ms per
    call on one 1536x1024 page): the secondary RT-DETRv2 detector, a conjoined-bubble split, the exact-size LANCZOS
    resample of the upscaled page, the per-bubble crop upscaling with the lite model, and the safe text boxes of the
    page's bubbles."""
    from mangatranslator_b200 import conjoined as Cj
    from mangatranslator_b200 import weights as Wt
    from mangatranslator_b200.core.image.image_utils import process_page_bubbles_device
    from mangatranslator_b200.preproc import resize_lanczos_device
    from mangatranslator_b200.rcan import RcanB200
    from mangatranslator_b200.rtdetr import RtDetrB200
    out = {}
    try:
        cfg, sd = Wt.rtdetr_model_and_state(0)
        det = RtDetrB200(sd, cfg, dev)
        out["rtdetr_secondary_detect_ms"] = round(_time_ms(lambda: det(page_dev, conf=0.35, imgsz=640)), 3)
        del det
        mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
        mask[300:720, 150:820] = 255
        boxes = torch.tensor([[150.0, 300.0, 520.0, 700.0], [440.0, 320.0, 820.0, 720.0]])
        out["conjoined_split_2_children_ms"] = round(_time_ms(lambda: Cj.split_conjoined_device(mask, boxes, window=(150, 300, 820, 720))), 3)
        up = torch.randint(0, 256, (2 * H, 2 * W, 3), dtype=torch.uint8, device=dev)
        ms = _time_ms(lambda: resize_lanczos_device(up, int(1.5 * H), int(1.5 * W)))
        traffic = (2 * H * 2 * W * 3 + 2 * (2 * H) * int(1.5 * W) * 3 + int(1.5 * H) * int(1.5 * W) * 3) / 1e9
        out["lanczos_3072x2048_to_2304x1536_ms"] = round(ms, 3)
        out["lanczos_gbs"] = round(traffic / (ms * 1e-3), 1)
        out["lanczos_hbm_frac"] = round(traffic / (ms * 1e-3) / peaks["hbm_gbs"], 4)
        lite = RcanB200(Wt.rcan_state_dict(0, n_resgroups=4, n_resblocks=6, unshuffle=2), dev)
        rgb = page_dev[:, :, [2, 1, 0]].contiguous()
        crops = [(60 + 64 * i, 90 + 105 * i, 60 + 64 * i + 130 + 13 * i, 90 + 105 * i + 110 + 17 * i) for i in range(BUBBLES)]  # ragged
        process_page_bubbles_device(rgb, crops, lite, 200, "min")        # builds the 12 ragged plans; next call captures graphs
        out["bubble_crops_lite_12_ms"] = round(_time_ms(lambda: process_page_bubbles_device(rgb, crops, lite, 200, "min"), reps=3), 3)
        del lite
    except Exception as e:                       # side measurements must never take the headline line down
        out["error"] = repr(e)[:200]
    try:                                         # safe text boxes of the page's bubbles (image_utils.py:173-348), one launch
        from mangatranslator_b200 import safebox_host as Sb
        yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
        masks = [((((xx - (i % 3 + 0.5) * W / 3) / 140.0) ** 2 + ((yy - (i // 3 + 0.5) * H / 4) / 110.0) ** 2) <= 1.0)
                 .to(torch.uint8).mul_(255).contiguous() for i in range(BUBBLES)]
        recs = Sb.safe_boxes_device(masks, 6.0)
        if int((recs["status"] != 0).sum()):
            raise RuntimeError(f"safe box statuses {recs['status'].tolist()}")
        ms = _time_ms(lambda: Sb.safe_boxes_device(masks, 6.0))
        out["safe_text_boxes_12_ms"] = round(ms, 3)
        out["safe_text_boxes_mask_gbs"] = round(BUBBLES * H * W / 1e9 / (ms * 1e-3), 1)
        del masks
    except Exception as e:
        out["safebox_error"] = repr(e)[:200]
    torch.cuda.empty_cache()
    return out


def stage_rooflines(stage, peaks, sam_variant):
    """Per-stage achieved rate against the measured peaks, from SURVEY.md section 8(d)'s algorithmic work per page:
    YOLOv8m-seg@1600x1088 468 GFLOP, SAM 2.1 encoder 207.3 (tiny) / 1621.9 (large) GFLOP, decoder 3.56 GFLOP x 12 boxes,
    RCAN x2 48.6 TFLOP, cleaning 13 MB of ideal HBM traffic."""
    enc = 1621.9 if sam_variant == "large" else 207.3
    det_gflop = 468.0 + enc + 3.56 * BUBBLES
    out = {}
    ms = stage.get("detect_segment")
    if ms:
        t = det_gflop / ms                      # GFLOP/ms = TFLOP/s
        out["detect_segment"] = dict(bound="tensor", algorithmic_gflop=round(det_gflop, 1), ms=round(ms, 3),
                                     achieved_tflops=round(t, 1), frac=round(t / peaks["bf16_sustained"], 4),
                                     note="latency-bound: ~300 small launches per page (one page per pass)")
    ms = stage.get("upscale")
    if ms:
        t = 48600.0 / ms
        out["upscale"] = dict(bound="tensor", algorithmic_gflop=48600.0, ms=round(ms, 3), achieved_tflops=round(t, 1),
                              frac=round(t / peaks["bf16_sustained"], 4),
                              note="fp32-grade result from 16/8-bit operand planes: 3x (fp16 + e5m2 correction) or 4x (bf16x3) the "
                                   "algorithmic FLOPs occupy the tensor pipe, see roofline.issued")
    ms = stage.get("clean_grouped") or stage.get("clean")
    if ms:
        g = 13.0e-3 / (ms * 1e-3)               # GB/s
        out["clean"] = dict(bound="hbm", algorithmic_mb=13.0, ms=round(ms, 3), achieved_gbs=round(g, 1),
                            frac=round(g / peaks["hbm_gbs"], 5),
                            note="integer bit-plane kernel, one CTA per bubble: latency-bound, not bandwidth-bound")
    return out


def run_ours(args, coord):
    from mangatranslator_b200 import _lib
    from mangatranslator_b200.core.pipeline import HotPathPipeline
    dev = torch.device("cuda", coord.local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    n_distinct = min(args.batch, 32)                       # 32 distinct pages = 151 MB of inputs > 126 MB L2
    pages = make_pages(n_distinct, 1000 * (coord.rank + 1))
    host = [torch.from_numpy(np.ascontiguousarray(p.image_rgb[:, :, ::-1])).pin_memory() for p in pages]
    devp = [h.to(dev) for h in host]
    boxes = [p.boxes_xyxy for p in pages]
    sam_variant = args.sam
    os.environ["MTB200_SAM_VARIANT"] = sam_variant
    pipe = HotPathPipeline(seg_model="sam2", upscale=True, upscale_model="model", device=dev)
    G = max(1, min(args.group, args.batch))                # pages per cleaning launch (HotPathPipeline.run_pages)
    outs_host = [torch.empty((2 * H, 2 * W, 3), dtype=torch.uint8, pin_memory=True) for _ in range(G)]
    sink = torch.empty((2 * H, 2 * W, 3), dtype=torch.uint8, device=dev)

    def groups():
        for g0 in range(0, args.batch, G):
            yield [i % n_distinct for i in range(g0, min(g0 + G, args.batch))]

    def step_device():
        for idx in groups():
            pipe.run_pages_device([devp[i] for i in idx], [boxes[i] for i in idx],
                                  consume=lambda i, out: sink.copy_(out))        # result stays in HBM

    def step_e2e():
        for idx in groups():
            pipe.run_pages([host[i] for i in idx], outs_host[:len(idx)], [boxes[i] for i in idx])

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(coord.local_rank)
    coord.barrier()
    torch.cuda.synchronize()
    sampler.start()
    from mangatranslator_b200 import graphs
    l0 = _lib.launch_count() + graphs.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() + graphs.replayed_launches - l0
    coord.barrier()
    clocks = sampler.stop()
    ms = coord.all_reduce_max(e0.elapsed_time(e1))
    value = coord.world * args.batch * args.steps / (ms / 1e3)
    # end-to-end: pinned host page in, host upscaled page out, every step
    for _ in range(min(2, max(1, args.warmup))):
        step_e2e()
    coord.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    torch.cuda.synchronize()
    coord.barrier()
    ms_e2e = coord.all_reduce_max(e0.elapsed_time(e1))
    e2e_v = coord.world * args.batch * args.steps / (ms_e2e / 1e3)
    # per-page stage times: ONE page on its own, warm (the timed loop runs groups of G pages, whose detector is one batched
    # plan and whose cleaning is one launch — `clean_grouped` below; the single-page plans are built by the first call here)
    pipe.run_page_device(devp[0], injected_boxes=boxes[0])
    stage = {}
    pipe.run_page_device(devp[0], injected_boxes=boxes[0], timings=stage)
    # cleaning as the bench runs it: the bubbles of G pages in one launch
    from mangatranslator_b200.core.image.cleaning import clean_pages_device
    from mangatranslator_b200.core.image.detection import detect_pages_device
    gd = [detect_pages_device([devp[i]], injected_boxes=[boxes[i]], own_masks=True)[0] for i in range(G)]
    clean_pages_device([devp[i] for i in range(G)], gd, processing_scale=math.sqrt(H * W / 1e6))
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    clean_pages_device([devp[i] for i in range(G)], gd, processing_scale=math.sqrt(H * W / 1e6))
    c1.record()
    torch.cuda.synchronize()
    stage["clean_grouped"] = c0.elapsed_time(c1) / G
    del gd
    if coord.rank != 0:
        return
    extras = measure_side_stages(devp[0], dev, peaks) if coord.world == 1 else {}     # N = 1 only, like cpu_baseline
    roof = measure_dominant_kernel(pipe, devp[0], peaks)
    # dram__bytes_read + dram__bytes_write per launch come from the committed `ncu --set full` capture of the same kernel
    # (profiles/): ncu cannot run inside the timed bench, so the figure is labelled with its source
    fp16c = "fp16c" in roof["kernel"]
    prof = os.path.join(ROOT, "profiles", "r02_fp16c_ncu.json" if fp16c else "r01_halo_cm_ncu.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
            roof["traffic_source"] = "from_profile:" + os.path.relpath(prof, ROOT)
        except Exception:
            pass
    gbase = None
    if coord.world == 1 and not args.no_gpu_baseline:
        try:
            gbase = gpu_baseline(dev)
        except Exception as e:                     # a comparator must never take the headline line down
            gbase = dict(error=repr(e)[:200])
    base = cpu_baseline() if coord.world == 1 and not args.no_cpu_baseline else None
    line = dict(metric=METRIC, value=value, unit="pages/s", n_gpus=coord.world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("fp16+e5m2 (fp32-grade: fp16 operand planes + an e5m2 correction product on tcgen05, fp32 accumulate; bf16 hi/lo "
                       "planes = bf16x3 outside the RCAN body); integer u8/bit ops for cleaning"
                       if "fp16c" in roof["kernel"] else
                       "bf16x3 (fp32-grade: hi/lo bf16 operand planes on tcgen05, fp32 accumulate); integer u8/bit ops for cleaning"),
                data="synthetic",
                config=dict(workload=(WORKLOAD if args.batch == 64 else WORKLOAD.replace("Batch 64", f"Batch {args.batch}")
                                      ).replace("SAM2.1-tiny", f"SAM2.1-{sam_variant}"),
                            pages_per_step_per_gpu=args.batch, page="1536x1024x3 u8", bubbles_per_page=BUBBLES,
                            weights="seeded synthetic (no checkpoints offline); detector runs in full, its boxes are "
                                    "replaced by the page's ground-truth boxes for the downstream stages",
                            l2="inputs larger than L2 (32 distinct pages = 151 MB; activations are GBs per page)",
                            pages_per_clean_launch=G, sam_variant=sam_variant,
                            parallelism=f"pages sharded i mod {coord.world}, no data-path collective"),
                e2e=dict(value=e2e_v, unit="pages/s", h2d_bytes_per_step=coord.world * args.batch * H * W * 3,
                         d2h_bytes_per_step=coord.world * args.batch * 4 * H * W * 3, ms_per_step=ms_e2e / args.steps),
                gpu_launches=int(launches) * coord.world, clocks=clocks, roofline=roof,
                stage_ms_per_page={k: round(v, 3) for k, v in stage.items()},
                stage_roofline=stage_rooflines(stage, peaks, sam_variant), side_stage_ms=extras)
    if base is not None:
        line["cpu_baseline"] = base
    if gbase is not None:
        line["gpu_baseline"] = gbase
    emit(line)


def run_corpus(args, coord):
    """BASELINE.json configs[4]: a corpus of page FILES through `batch_translate_images` (the reference's batch entry,
    core/pipeline.py:2481-2731), sharded page i -> rank i mod R: decode -> detect -> segment -> clean -> 2x upscale ->
    encode -> write, strong scaling (the corpus is fixed, ranks split it).  Reports files/s over the wall time of the
    slowest rank and where the time went (seconds summed over a rank's threads, max over ranks)."""
    import shutil
    import tempfile
    from PIL import Image
    from mangatranslator_b200.core import pipeline as P
    from mangatranslator_b200.core.config import MangaTranslatorConfig
    dev = torch.device("cuda", coord.local_rank)
    torch.cuda.set_device(dev)
    root = coord.broadcast(tempfile.mkdtemp(prefix="mtb200_corpus_") if coord.rank == 0 else None)
    inp, out = os.path.join(root, "in"), os.path.join(root, "out")
    n, fmt = args.corpus, args.corpus_format
    if coord.rank == 0:
        os.makedirs(inp, exist_ok=True)
        distinct = min(n, 16)
        pages = make_pages(distinct, 7000)
        for i in range(distinct):
            Image.fromarray(pages[i].image_rgb).save(os.path.join(inp, f"page_{i:05d}.png"), compress_level=1)
        for i in range(distinct, n):                      # hard links: a large corpus without minutes of PNG writing
            os.link(os.path.join(inp, f"page_{i % distinct:05d}.png"), os.path.join(inp, f"page_{i:05d}.png"))
    coord.barrier()
    cores = os.cpu_count() or 1
    workers = args.save_workers if args.save_workers > 0 else max(1, cores // coord.world - 1)
    os.environ["MTB200_SAVE_WORKERS"] = str(workers)
    cfg = MangaTranslatorConfig(cleaning_only=True)
    cfg.detection.seg_model = "sam2"
    cfg.detection.conjoined_detection = False             # the secondary RT-DETR detector has no checkpoint offline
    cfg.output.upscale_final_image, cfg.output.image_upscale_factor, cfg.output.image_upscale_model = True, 2.0, "model"
    cfg.output.output_format = fmt
    os.environ["MTB200_PNG_WRITER"] = args.png_writer     # the config classes keep the reference's fields: the knob is an env var
    # warm-up: models, plans and CUDA graphs (a separate tiny batch that is not timed)
    warm = os.path.join(root, "warm")
    if coord.rank == 0:
        os.makedirs(warm, exist_ok=True)
        for i in range(2 * coord.world):
            os.link(os.path.join(inp, f"page_{i % min(n, 16):05d}.png"), os.path.join(warm, f"w_{i:03d}.png"))
    coord.barrier()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):          # the page driver logs to stdout; this program prints ONE JSON line
        P.batch_translate_images(warm, cfg, os.path.join(root, "warm_out"))
    P.STAGE_CLOCK.reset()
    sampler = ClockSampler(coord.local_rank)
    coord.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        res = P.batch_translate_images(inp, cfg, out)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    stage = P.STAGE_CLOCK.snapshot()
    wall_max = coord.all_reduce_max(wall)
    stage_max = {k: coord.all_reduce_max(stage.get(k, 0.0)) for k in ("decode", "decode_wait", "render", "render_host_prep",
                                                                     "render_device", "render_device_png", "render_to_pil", "prep_ahead",
                                                                     "save_wait", "encode", "bubbles")}
    ok = res["success_count"] if coord.rank == 0 else 0
    if coord.rank == 0:
        written = sum(len(f) for _, _, f in os.walk(out))
        sizes = [os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(out) for f in fs][:64]
        per_rank = n / coord.world
        line = dict(metric="manga page files/sec through batch_translate_images (decode->detect->segment->clean->upscale->encode)",
                    value=n / wall_max, unit="files/s", n_gpus=coord.world, steps=1, warmup=1, ms_per_step=1000.0 * wall_max,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="fp16+e5m2 / bf16x3 (see the default line)",
                    data="synthetic", mode="corpus",
                    config=dict(workload=f"{n}-page corpus of 1536x1024 PNG files, full pipeline through batch_translate_images, "
                                         f"{fmt} output, sharded i mod {coord.world} (BASELINE.json configs[4])",
                                files=n, output_format=fmt, png_writer=args.png_writer, save_workers_per_rank=workers,
                                host_cores=cores, l2="file-backed pages: every page is decoded and uploaded once"),
                    files_written=written, success_count=ok, error_count=res.get("error_count"),
                    avg_output_mb=round(float(np.mean(sizes)) / 1e6, 2) if sizes else None,
                    seconds_per_rank_max=stage_max,
                    per_page_ms={k: round(1000.0 * v / per_rank, 2) for k, v in stage_max.items()},
                    limiter=max(("render", "encode", "decode"),
                                key=lambda k: stage_max[k] / (workers if k == "encode" else 1)),
                    clocks=clocks,
                    e2e=dict(value=n / wall_max, unit="files/s", h2d_bytes_per_step=n * H * W * 3, d2h_bytes_per_step=n * 4 * H * W * 3),
                    gpu_launches=None)
        emit(line)
        shutil.rmtree(root, ignore_errors=True)
    coord.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--group", type=int, default=8, help="pages whose bubbles share one cleaning launch")
    ap.add_argument("--sam", default="tiny", choices=["tiny", "large"],
                    help="SAM 2.1 variant (BASELINE.json names tiny; large = the checkpoint the reference loads)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--only-gpu-baseline", action="store_true", help="print the torch-eager-on-GPU comparator and exit")
    ap.add_argument("--corpus", type=int, default=0, help="N > 0: corpus mode, N page files through batch_translate_images")
    ap.add_argument("--corpus-format", default="png", choices=["png", "jpeg"])
    ap.add_argument("--png-writer", default="auto", choices=["auto", "pil", "device"],
                    help="PNG encoder of the batch path: PIL on host threads, or the device deflate encoder")
    ap.add_argument("--save-workers", type=int, default=0, help="writer threads per rank (0 = host cores / ranks - 1)")
    args = ap.parse_args()
    _stdout_carries_only_the_json_line()
    if args.only_gpu_baseline:
        emit(dict(gpu_baseline=gpu_baseline(torch.device("cuda", 0))))
        return
    from mangatranslator_b200.core.batch_coordinator import PageShardCoordinator
    if args.impl == "reference":
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    coord = PageShardCoordinator(backend="gloo" if args.impl == "reference" else None)
    try:
        if args.impl == "reference":
            run_reference(args, coord)
        elif args.corpus > 0:
            run_corpus(args, coord)
        else:
            run_ours(args, coord)
    finally:
        coord.close()


if __name__ == "__main__":
    main()
