"""Pipeline orchestration for the vision hot path — drop-in for the stage-calling skeleton of the reference's
core/pipeline.py (`translate_and_render` :638-2024, `_clean_speech_bubbles_for_page` :76-130,
`batch_translate_images` :2481-2731) in the modes that need no LLM / text renderer:

    cleaning_only = True   detect -> segment -> clean [-> upscale_final_image]     (pipeline.py:995-996,1993-2000)
    upscaling_only = True  upscale only                                             (pipeline.py:723-762)

Anything that needs OCR/translation/rendering is outside this build and raises.

Two entry levels:
  * `translate_and_render` / `batch_translate_images`: the reference's signatures, PIL in / PIL + files out.
  * `HotPathPipeline`: the device-resident batch engine the bench and the batch path use — pages go H2D once (pinned),
    every stage runs in CUDA kernels, one D2H of the finished page.
"""
from __future__ import annotations

import math
import os
import re
import time
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image

from mangatranslator_b200.core.batch_coordinator import PageShardCoordinator
from mangatranslator_b200.core.caching import get_cache
from mangatranslator_b200.core.config import MangaTranslatorConfig
from mangatranslator_b200.core.image.cleaning import clean_pages_device, clean_speech_bubbles
from mangatranslator_b200.core.image.detection import detect_pages_device, detect_speech_bubbles
from mangatranslator_b200.core.image.image_utils import (convert_image_to_target_mode, cv2_to_pil,
                                                          save_image_with_compression, upscale_image)
from mangatranslator_b200.core.ml.model_manager import get_model_manager
from mangatranslator_b200.utils.exceptions import CancellationError, ImageProcessingError, ValidationError
from mangatranslator_b200.utils.logging import log_message



class _StageClock:
    """Where a batch spends its time (seconds, summed over threads): `decode` (file -> PIL), `render` (everything between
    the decoded page and the finished PIL image: H2D, the device stages, D2H, host glue), `encode` (PIL -> file),
    `save_wait` (the page loop blocked on the writer pool's back-pressure).  Read by bench.py --corpus."""

    def __init__(self):
        import threading
        self._lock, self.seconds, self.counts = threading.Lock(), {}, {}

    def add(self, name: str, dt: float) -> None:
        with self._lock:
            self.seconds[name] = self.seconds.get(name, 0.0) + dt
            self.counts[name] = self.counts.get(name, 0) + 1

    def reset(self) -> None:
        with self._lock:
            self.seconds, self.counts = {}, {}

    def snapshot(self) -> dict:
        with self._lock:
            return {k: round(v, 4) for k, v in self.seconds.items()}


STAGE_CLOCK = _StageClock()


def _processing_scale(width: int, height: int, auto: bool = True) -> float:
    return math.sqrt((width * height) / 1_000_000) if auto else 1.0     # pipeline.py:765-772


def _clean_speech_bubbles_for_page(pil_image, config: MangaTranslatorConfig, bubble_data, processing_scale, verbose):
    """pipeline.py:76-130: clean with the pre-computed detections; on failure keep the original page."""
    try:
        cleaned, processed = clean_speech_bubbles(
            pil_image, config.yolo_model_path, config.detection.confidence, pre_computed_detections=bubble_data,
            device=config.device, thresholding_value=config.cleaning.thresholding_value,
            use_otsu_threshold=config.cleaning.use_otsu_threshold, roi_shrink_px=config.cleaning.roi_shrink_px,
            verbose=verbose, processing_scale=processing_scale,
            conjoined_confidence=config.detection.conjoined_confidence,
            inpaint_colored_bubbles=config.cleaning.inpaint_colored_bubbles,
            bubble_detector_model=config.detection.bubble_detector_model)
        return cv2_to_pil(cleaned), processed
    except Exception as e:
        log_message(f"Error during bubble cleaning: {e}. Proceeding with uncleaned image.", always_print=True)
        return pil_image, []


def _target_mode(config: MangaTranslatorConfig, image_path, output_path) -> str:
    """pipeline.py:702-712: RGB for JPEG output (or `auto` with a .jpg/.jpeg target), RGBA for everything else."""
    ext = (Path(output_path) if output_path else Path(image_path)).suffix.lower()
    fmt = config.output.output_format
    return "RGB" if fmt == "jpeg" or (fmt == "auto" and ext in (".jpg", ".jpeg")) else "RGBA"


def _load_page(image_path) -> Image.Image:
    """Open and fully decode a page file (pipeline.py:689-696)."""
    t0 = time.perf_counter()
    try:
        pil = Image.open(image_path)
        pil.load()
        return pil
    except Exception as e:
        raise ImageProcessingError(f"Error loading image {image_path}: {e}")
    finally:
        STAGE_CLOCK.add("decode", time.perf_counter() - t0)


def _resolve_pre_upscale_factor(pre_cfg, verbose: bool = False) -> float:
    """core/pipeline.py:602-614: the initial-upscale factor is clamped to [1, 8] and switched off at or below 1.01."""
    if pre_cfg is None or not pre_cfg.enabled:
        return 1.0
    factor = max(1.0, min(float(pre_cfg.factor or 1.0), 8.0))
    if factor <= 1.01:
        return 1.0
    log_message(f"Initial upscaling enabled: {factor:.2f}x", verbose=verbose)
    return factor


def _apply_pre_upscale_if_needed(image: Image.Image, config: MangaTranslatorConfig, verbose: bool = False):
    """core/pipeline.py:617-635: the initial upscale uses the OUTPUT upscale model setting.  Returns (image, factor)."""
    factor = _resolve_pre_upscale_factor(getattr(config, "preprocessing", None), verbose)
    if factor == 1.0:
        return image, 1.0
    model_type = getattr(config.output, "image_upscale_model", "model_lite") if hasattr(config, "output") else "model_lite"
    return upscale_image(image, factor, model_type=model_type, verbose=verbose), factor


class EncodedPng:
    """A finished page that left the device already PNG-encoded (mangatranslator_b200/png_device.py): `_save_page` writes
    the bytes as they are.  `size` / `mode` describe the picture like a PIL image would."""

    def __init__(self, data, size, mode: str):
        self._data, self.size, self.mode = data, size, mode

    @property
    def data(self) -> bytes:
        """The file's bytes.  The encoder hands over the deflate stream and its checksums' ingredients; the container
        (zlib trailer, chunk CRCs) is put around it here, i.e. on the writer thread that saves the page."""
        if not isinstance(self._data, (bytes, bytearray)):
            from mangatranslator_b200.png_device import finalize_png
            self._data = finalize_png(self._data)
        return self._data

    @property
    def width(self):
        return self.size[0]

    @property
    def height(self):
        return self.size[1]


_FAST: Dict[tuple, "HotPathPipeline"] = {}
_PNG = {}


class _PinnedPages:
    """Pinned BGR staging buffers for decoded pages, a few per page size: the thread that decodes page i + 1 also converts it
    and fills one while page i is on the device, so neither the RGB conversion nor the copy into pinned memory sits on the
    page loop's critical path (cudaHostAlloc per page would cost more than it saves, hence the pool)."""

    def __init__(self, per_size: int = 3):
        import threading
        self._lock, self._free, self._made, self._per = threading.Lock(), {}, {}, per_size

    def acquire(self, h: int, w: int):
        with self._lock:
            free = self._free.setdefault((h, w), [])
            if free:
                return free.pop()
            if self._made.get((h, w), 0) >= self._per:
                return None                              # all in flight: the caller falls back to the unpinned path
            if len(self._made) >= 4 and (h, w) not in self._made:
                return None
            self._made[(h, w)] = self._made.get((h, w), 0) + 1
        return torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True)

    def release(self, buf) -> None:
        with self._lock:
            self._free.setdefault((int(buf.shape[0]), int(buf.shape[1])), []).append(buf)


_PINNED = _PinnedPages()


def _load_page_prepared(image_path, config: MangaTranslatorConfig) -> Image.Image:
    """`_load_page` for the decode-ahead thread: additionally, for pages the device engine can take (opaque RGB / RGBA / L,
    `cleaning_only`), the BGR copy the engine uploads is made here, into a pinned buffer attached to the image."""
    pil = _load_page(image_path)
    try:
        if (config.cleaning_only and os.environ.get("MTB200_FAST_PATH", "1") != "0" and torch.cuda.is_available()
                and pil.mode in ("RGB", "RGBA", "L") and not (pil.mode == "RGBA" and pil.getextrema()[3][0] < 255)):
            t0 = time.perf_counter()
            rgb = np.asarray(pil if pil.mode == "RGB" else pil.convert("RGB"))
            buf = _PINNED.acquire(rgb.shape[0], rgb.shape[1])
            if buf is not None:
                np.copyto(buf.numpy(), rgb[:, :, ::-1])
                pil._mtb_bgr_pinned = buf
            STAGE_CLOCK.add("prep_ahead", time.perf_counter() - t0)
    except Exception:
        pass                                             # preparation is an optimisation: the page itself is fine
    return pil


def _png_encoder(device):
    from mangatranslator_b200.png_device import PngEncoderB200
    if device not in _PNG:
        _PNG[device] = PngEncoderB200(device)
    return _PNG[device]


def _device_png_wanted(output_path) -> bool:
    """Who encodes the PNG files of the batch path: MTB200_PNG_WRITER = "device" (csrc/png_kernels.cu), "pil" (host
    threads, like the reference's writer) or "auto".  A 3072x2048 page costs ~0.63 s of one host core with PIL (measured),
    so PIL keeps up only with ~6 or more writer cores per GPU; measured on 2 GPUs: with 11 writer threads per rank PIL
    13.5 files/s vs device 11.6 (the device encoder adds ~20 ms to a page's critical path), with 3 threads per rank — the
    share of a 32-core 8-GPU box — PIL 8.9 vs device 11.6.  "auto" = device when the rank has fewer than 8 host cores."""
    if output_path is None or Path(output_path).suffix.lower() != ".png":
        return False
    mode = os.environ.get("MTB200_PNG_WRITER", "auto")
    if mode == "auto":
        world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
        return (os.cpu_count() or 1) // world < 8
    return mode == "device"


def _fast_path_pipeline(config: MangaTranslatorConfig, pil: Image.Image) -> Optional["HotPathPipeline"]:
    """The device-resident engine for this configuration, or None when the page must go through the stage functions.

    `cleaning_only` pages whose stages are all on the device path (SAM 2.1 masks, threshold cleaning, optional final
    upscale) need no PIL / numpy round trips between the stages: the page is uploaded once, every stage runs in CUDA
    kernels and the finished page comes back once (MTB200_FAST_PATH=0 forces the stage functions).  The results are the
    stage functions' results (tests/test_pipeline_gpu.py::test_batch_fast_path_equals_stage_functions).  Not eligible:
    translucent pages (the stage functions clean the RGBA page itself), the proto-mask segmenter, coloured-bubble
    inpainting, outside-text processing, OSB-text verification when its detector is available."""
    if os.environ.get("MTB200_FAST_PATH", "1") == "0" or not config.cleaning_only:
        return None
    if (config.detection.seg_model != "sam2" or config.cleaning.inpaint_colored_bubbles
            or getattr(config.outside_text, "enabled", False) or not config.preprocessing.auto_scale
            or not torch.cuda.is_available()):
        return None
    if pil.mode not in ("RGB", "RGBA", "L"):
        return None
    if pil.mode == "RGBA" and pil.getextrema()[3][0] < 255:
        return None
    if getattr(config.detection, "use_osb_text_verification", False):
        # OSB-text verification (box expansion, text-safe conjoined cuts) lives in the stage function.  Without its model
        # the reference skips it (detection.py:198-201), and so does the device path; with a model available the page goes
        # through `detect_speech_bubbles`
        from mangatranslator_b200.core.ml.model_manager import ModelType
        mm0 = get_model_manager()
        if (mm0.is_loaded(ModelType.YOLO_OSBTEXT) or mm0.model_paths[ModelType.YOLO_OSBTEXT].is_file()
                or os.environ.get("MTB200_SYNTHETIC_OSBTEXT", "0") == "1"):
            return None
    key = (float(config.detection.confidence), int(config.cleaning.thresholding_value), float(config.cleaning.roi_shrink_px),
           bool(config.cleaning.use_otsu_threshold), bool(config.detection.conjoined_detection),
           float(config.detection.conjoined_confidence), bool(config.output.upscale_final_image),
           str(config.output.image_upscale_model), str(config.yolo_model_path or ""))
    if key not in _FAST:
        # objects a caller injected into ModelManager.models (duck-typed detectors / segmenters, like the reference's own
        # integrations do) only promise the reference's call shapes: they go through the stage functions
        from mangatranslator_b200.rcan import RcanB200
        from mangatranslator_b200.sam2_api import Sam2ModelB200
        from mangatranslator_b200.yolo import YoloB200
        mm = get_model_manager()
        try:
            det = mm.load_yolo_speech_bubble(key[8] or None)
            seg = mm.load_sam2()
            up = (mm.load_upscale() if key[7] == "model" else mm.load_upscale_lite()) if key[6] else None
        except Exception:
            return None                            # the stage functions report / degrade like the reference does
        if not (isinstance(det, YoloB200) and isinstance(seg, tuple) and isinstance(seg[1], Sam2ModelB200)
                and (up is None or isinstance(up, RcanB200))):
            return None
        _FAST[key] = HotPathPipeline(confidence=key[0], seg_model="sam2", upscale=key[6],
                                     upscale_model="model" if key[7] == "model" else "model_lite",
                                     thresholding_value=key[1], roi_shrink_px=key[2], use_otsu_threshold=key[3],
                                     conjoined_detection=key[4], conjoined_confidence=key[5], yolo_model_path=key[8])
    return _FAST[key]


def _render_fast(pipe: "HotPathPipeline", pil: Image.Image, config: MangaTranslatorConfig, png_mode: Optional[str] = None):
    """One page through the device engine: RGB page -> cleaned (and, for a 2x final upscale, upscaled) RGB page.  With
    `png_mode` ("RGB" / "RGBA") the finished page is PNG-encoded on the device and only the compressed bytes come back."""
    from mangatranslator_b200._lib import device_section
    t_prep = time.perf_counter()
    ahead = getattr(pil, "_mtb_bgr_pinned", None)        # filled by the decode-ahead thread (_load_page_prepared)
    if ahead is not None:
        h, w, rgb = int(ahead.shape[0]), int(ahead.shape[1]), None
    else:
        rgb = np.asarray(pil if pil.mode == "RGB" else pil.convert("RGB"))
        h, w = rgb.shape[:2]
    with device_section:
        # pinned staging buffers are kept per page size (cudaHostAlloc per page costs more than the copy it speeds up)
        stage = pipe.__dict__.setdefault("_staging", {})
        key = (h, w)
        if key not in stage:
            if len(stage) >= 4:
                stage.pop(next(iter(stage)))
            stage[key] = dict(inp=torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True))
        sb = stage[key]
        if ahead is not None:
            host = ahead
        else:
            np.copyto(sb["inp"].numpy(), rgb[:, :, ::-1])                    # BGR, like the cleaning stage wants
            host = sb["inp"]
        STAGE_CLOCK.add("render_host_prep", time.perf_counter() - t_prep)
        try:
            return _render_fast_device(pipe, pil, config, png_mode, host, sb, h, w)
        finally:
            if ahead is not None:                        # every path below has synchronised: the upload is done
                pil._mtb_bgr_pinned = None
                _PINNED.release(ahead)


def _render_fast_device(pipe: "HotPathPipeline", pil: Image.Image, config: MangaTranslatorConfig, png_mode, host, sb, h: int, w: int):
    """The device half of `_render_fast` (called inside the device section with the page in a pinned BGR buffer)."""
    two_x = pipe.rcan is not None and abs(float(config.output.image_upscale_factor) - float(pipe.rcan.scale)) < 1e-9
    if two_x and png_mode:
        t0 = time.perf_counter()
        out, dets, _ = pipe.run_page_device(host.to(pipe.device, non_blocking=True))    # RGB, 2H x 2W, stays on the device
        torch.cuda.current_stream().synchronize()
        STAGE_CLOCK.add("render_device", time.perf_counter() - t0)
        STAGE_CLOCK.add("bubbles", float(len(dets)))
        t0 = time.perf_counter()
        data = _png_encoder(pipe.device).encode(out, out_channels=4 if png_mode == "RGBA" else 3, finalize=False)
        STAGE_CLOCK.add("render_device_png", time.perf_counter() - t0)
        return EncodedPng(data, (int(out.shape[1]), int(out.shape[0])), png_mode)
    if two_x:
        if "out" not in sb:
            sb["out"] = torch.empty((pipe.rcan.scale * h, pipe.rcan.scale * w, 3), dtype=torch.uint8, pin_memory=True)
        t0 = time.perf_counter()
        out, dets, _ = pipe.run_page(host, out_host=sb["out"])      # RGB, 2H x 2W
        STAGE_CLOCK.add("render_device", time.perf_counter() - t0)
        STAGE_CLOCK.add("bubbles", float(len(dets)))
        t0 = time.perf_counter()
        img = Image.fromarray(out.numpy().copy())
        STAGE_CLOCK.add("render_to_pil", time.perf_counter() - t0)
        return img
    page = host.to(pipe.device)
    dets = detect_pages_device([page], confidence=pipe.confidence, imgsz=pipe.imgsz, seg_model="sam2", **pipe.conjoined)[0]
    batch = clean_pages_device([page], [dets], thresholding_value=pipe.thr, roi_shrink_px=pipe.shrink,
                               use_otsu_threshold=pipe.otsu, processing_scale=_processing_scale(pil.width, pil.height,
                                                                                              config.preprocessing.auto_scale))
    if png_mode and not config.output.upscale_final_image:
        out = batch.pages_out[0][:, :, [2, 1, 0]].contiguous()
        data = _png_encoder(pipe.device).encode(out, out_channels=4 if png_mode == "RGBA" else 3, finalize=False)
        return EncodedPng(data, (int(out.shape[1]), int(out.shape[0])), png_mode)
    cleaned = Image.fromarray(np.ascontiguousarray(batch.pages_out[0].cpu().numpy()[:, :, ::-1]))
    if config.output.upscale_final_image:      # any other factor: the wrapper's pass loop + exact-size resample
        cleaned = upscale_image(cleaned, config.output.image_upscale_factor, model_type=config.output.image_upscale_model,
                                verbose=config.verbose)
    return cleaned


def _render_page(image_path, config: MangaTranslatorConfig, output_path=None, cancellation_manager=None, preloaded=None,
                 device_png: bool = False):
    """Everything of translate_and_render up to (not including) the save: returns (image, target_mode).  `preloaded`:
    a pending decode of this page (concurrent.futures.Future of _load_page) started while the previous page was on the
    device."""
    verbose = config.verbose
    if not (config.cleaning_only or config.upscaling_only):
        raise ValidationError("this build implements the vision hot path only: set cleaning_only or upscaling_only "
                              "(translation/rendering are outside scope, SURVEY.md §8)")
    if cancellation_manager is not None and cancellation_manager.is_cancelled():
        raise CancellationError("Process cancelled by user.")
    image_path = Path(image_path)
    t0 = time.perf_counter()
    pil = preloaded.result() if preloaded is not None else _load_page(image_path)
    if preloaded is not None:                       # the decode has its own clock; what was left of it is waited for here
        STAGE_CLOCK.add("decode_wait", time.perf_counter() - t0)
    t_render = time.perf_counter()
    try:
        return _render_decoded(pil, image_path, config, output_path, verbose, device_png)
    finally:
        STAGE_CLOCK.add("render", time.perf_counter() - t_render)


def _render_decoded(pil: Image.Image, image_path: Path, config: MangaTranslatorConfig, output_path, verbose: bool,
                    device_png: bool = False):
    target_mode = _target_mode(config, image_path, output_path)
    get_cache().clear_all()                 # every mode: nothing cached for the previous page outlives it
    pre = _resolve_pre_upscale_factor(getattr(config, "preprocessing", None))
    if not config.upscaling_only and pre == 1.0 and (fast := _fast_path_pipeline(config, pil)) is not None:
        # the device engine works on the opaque RGB page: converting to the target mode first (an alpha plane of 255 for
        # PNG output) and back would only cost two host passes over the page
        return (_render_fast(fast, pil, config, target_mode if device_png and _device_png_wanted(output_path) else None),
                target_mode)
    pil = convert_image_to_target_mode(pil, target_mode, verbose)
    pil, _ = _apply_pre_upscale_if_needed(pil, config, verbose)
    if config.upscaling_only:
        # core/pipeline.py:723-737: in this mode the final upscale still depends on output.upscale_final_image (the
        # "initial" upscale above is what runs when only preprocessing.enabled is set)
        log_message("Upscaling only mode - skipping detection and translation", always_print=True)
        out = pil
        if config.output.upscale_final_image:
            out = upscale_image(out, config.output.image_upscale_factor, model_type=config.output.image_upscale_model,
                                verbose=verbose)
    elif (fast := _fast_path_pipeline(config, pil)) is not None:
        out = _render_fast(fast, pil, config, target_mode if device_png and _device_png_wanted(output_path) else None)
    else:
        scale = _processing_scale(pil.width, pil.height, config.preprocessing.auto_scale)
        get_cache().set_current_image(pil, verbose)
        try:
            bubbles, _ = detect_speech_bubbles(
                image_path, config.yolo_model_path, config.detection.confidence, verbose=verbose,
                device=config.device, seg_model=config.detection.seg_model,
                conjoined_detection=config.detection.conjoined_detection,
                conjoined_confidence=config.detection.conjoined_confidence, image_override=pil,
                osb_enabled=config.outside_text.enabled,
                osb_text_verification=config.detection.use_osb_text_verification,
                bubble_detector_model=config.detection.bubble_detector_model)
        except Exception as e:      # pipeline.py:797-800
            log_message(f"Error during detection: {e}", always_print=True)
            bubbles = []
        out, _ = _clean_speech_bubbles_for_page(pil, config, bubbles, scale, verbose)
        if config.output.upscale_final_image:
            out = upscale_image(out, config.output.image_upscale_factor, model_type=config.output.image_upscale_model,
                                verbose=verbose)
    return out, target_mode


def _save_page(image: Image.Image, target_mode: str, output_path, config: MangaTranslatorConfig) -> None:
    """pipeline.py:1996-2018: convert to the target mode, save with the configured compression; a failed save is logged
    and re-raised."""
    t0 = time.perf_counter()
    if isinstance(image, EncodedPng):               # encoded on the device: the bytes are the file
        try:
            path = Path(output_path)
            path.parent.mkdir(parents=True, exist_ok=True)
            path.write_bytes(image.data)
            return
        except Exception as e:
            log_message(f"Failed to save image: Failed to save image to {output_path}", always_print=True)
            raise ImageProcessingError(f"Failed to save image to {output_path}") from e
        finally:
            STAGE_CLOCK.add("encode", time.perf_counter() - t0)
    if image.mode != target_mode:
        image = image.convert(target_mode)
    try:
        save_image_with_compression(image, output_path, jpeg_quality=config.output.jpeg_quality,
                                    png_compression=config.output.png_compression, verbose=config.verbose)
    except ImageProcessingError as e:
        log_message(f"Failed to save image: {e}", always_print=True)
        raise
    finally:
        STAGE_CLOCK.add("encode", time.perf_counter() - t0)


def translate_and_render(image_path, config: MangaTranslatorConfig, output_path=None, cancellation_manager=None,
                         previous_context_images=None, previous_context_texts=None,
                         previous_context_texts_provider=None, ocr_texts_out=None) -> Image.Image:
    start = time.time()
    out, target_mode = _render_page(image_path, config, output_path, cancellation_manager)
    if output_path:
        _save_page(out, target_mode, output_path, config)
    log_message(f"Processing completed in {time.time() - start:.2f}s", always_print=True)
    return out


def _process_page(path, config, out_path, cancellation_manager, saver, preloaded=None):
    """One page of the batch: with a writer pool the encode + write of this page overlaps the next page's device work
    (returns the pending save), otherwise exactly translate_and_render."""
    if saver is None:
        translate_and_render(path, config, out_path, cancellation_manager=cancellation_manager)
        return None
    start = time.time()
    out, target_mode = _render_page(path, config, out_path, cancellation_manager, preloaded, device_png=True)
    log_message(f"Processing completed in {time.time() - start:.2f}s (save queued)", verbose=config.verbose)
    t0 = time.perf_counter()
    fut = saver.submit(_save_page, out, target_mode, out_path, config)
    STAGE_CLOCK.add("save_wait", time.perf_counter() - t0)
    return fut


class _BoundedPool:
    """ThreadPoolExecutor whose submit() blocks while `depth` tasks are pending, so finished pages cannot pile up in host
    memory faster than they are written (an upscaled page is ~25 MB)."""

    def __init__(self, workers: int, depth: int):
        import concurrent.futures
        import threading
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=workers, thread_name_prefix="mtb200-save")
        self._slots = threading.BoundedSemaphore(depth)

    def submit(self, fn, *args):
        self._slots.acquire()
        fut = self._pool.submit(fn, *args)
        fut.add_done_callback(lambda _f: self._slots.release())
        return fut

    def submit_unbounded(self, fn, *args):
        """For the page decode that runs one page ahead: at most one is pending, it must never wait behind the saves'
        back-pressure."""
        return self._pool.submit(fn, *args)

    def close(self):
        self._pool.shutdown(wait=True)


_DIGITS = re.compile(r"(\d+)")
BATCH_EXTS = (".jpg", ".jpeg", ".png", ".webp")                    # pipeline.py:2533


def natural_path_key(path: Path):
    """Reading-order sort of page files (pipeline.py:133-142): per path component, digit runs compare as numbers and sort
    before text, text compares case-insensitively, the raw token breaks ties."""
    return tuple(tuple((0, int(t), t) if t.isdigit() else (1, t.lower(), t) for t in _DIGITS.split(part) if t)
                 for part in path.parts)


def resolve_output_path(img_path: Path, input_dir: Path, output_dir: Path, config: MangaTranslatorConfig,
                        preserve_structure: bool):
    """(output file, display name, error key) of one page (pipeline.py:2027-2064): `<stem>_translated` + the extension
    output_format asks for ("jpeg" -> .jpg, "png" -> .png, "auto" / anything else -> the source's), under the mirrored
    sub-directory when the structure is preserved."""
    if preserve_structure:
        rel = img_path.relative_to(input_dir)
        folder = output_dir / rel.parent
        folder.mkdir(parents=True, exist_ok=True)
        shown = key = str(rel)
        stem = rel.stem
    else:
        folder, shown, key, stem = output_dir, img_path.name, img_path.name, img_path.stem
    src_ext = img_path.suffix.lower()
    fmt = config.output.output_format
    if fmt not in ("jpeg", "png", "auto"):
        log_message(f"Warning: Invalid output_format '{fmt}' in config. Using original extension '{src_ext}'.",
                    always_print=True)
    ext = {"jpeg": ".jpg", "png": ".png"}.get(fmt, src_ext)
    return folder / f"{stem}_translated{ext}", shown, key


def resolve_source_path(img_path, source_path_map=None) -> str:
    """utils/path_list.py:12-33: absolute path of a page, mapped back to where the UI / CLI copied it from."""
    path = Path(img_path)
    try:
        resolved = str(path.resolve())
    except OSError:
        resolved = str(path)
    if source_path_map:
        for k in (resolved, str(path)):
            if k in source_path_map:
                return source_path_map[k]
    return resolved


def write_failed_paths(output_dir, paths) -> Optional[Path]:
    """utils/path_list.py:36-75: unique absolute paths, one per line, to <output_dir>/failed_paths.txt."""
    unique: List[str] = []
    for raw in paths:
        text = "" if raw is None else str(raw).strip()
        if not text:
            continue
        try:
            text = str(Path(text).resolve())
        except OSError:
            pass
        if text not in unique:
            unique.append(text)
    if not unique:
        return None
    try:
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        (out / "failed_paths.txt").write_text("\n".join(unique) + "\n", encoding="utf-8")
        return out / "failed_paths.txt"
    except OSError as e:
        log_message(f"Warning: failed to write failed_paths.txt: {e}", always_print=True)
        return None


def _list_batch_images(input_dir: Path, preserve_structure: bool) -> List[Path]:
    if preserve_structure:
        files = [Path(root) / f for root, _, names in os.walk(input_dir) for f in names]
        files = [f for f in files if f.suffix.lower() in BATCH_EXTS]
        return sorted(files, key=lambda f: natural_path_key(f.relative_to(input_dir)))
    files = [f for f in input_dir.iterdir() if f.is_file() and f.suffix.lower() in BATCH_EXTS]
    return sorted(files, key=lambda f: natural_path_key(Path(f.name)))


def batch_translate_images(input_dir, config: MangaTranslatorConfig, output_dir=None, progress_callback=None,
                           preserve_structure: bool = False, cancellation_manager=None, source_path_map=None) -> dict:
    """Reference contract (pipeline.py:2481-2731): the images of `input_dir` (its sub-directories too when
    `preserve_structure`) in natural order, each written as `<stem>_translated.<ext>`; returns {success_count,
    error_count, errors{name: message}, failed_image_paths[source paths][, failed_paths_file][, retry_*]}; failed pages
    are retried once when `config.retry_failed_once`; a cancellation propagates as CancellationError.
    The encode + write of a finished page runs on a small worker pool (`MTB200_SAVE_WORKERS`, default 2, 0 = inline) so it
    overlaps the next page's device work, and the next page's file is decoded one page ahead on the same pool; a page is
    counted once its file is on disk.
    Under torchrun the sorted page list is sharded over the ranks (`PageShardCoordinator`: page i goes to rank i mod R,
    no data-path collective); every rank returns its own counts and rank 0 the merged result with the failure file."""
    empty = {"success_count": 0, "error_count": 0, "errors": {}, "failed_image_paths": []}
    input_dir = Path(input_dir)
    if not input_dir.is_dir():
        log_message(f"Input path '{input_dir}' is not a directory", always_print=True)
        return empty
    coord = PageShardCoordinator()
    # the default directory carries a timestamp: every rank must use rank 0's
    out_dir = Path(coord.broadcast(str(Path(output_dir) if output_dir else Path("./output") / time.strftime("%Y%m%d_%H%M%S"))))
    out_dir.mkdir(parents=True, exist_ok=True)
    files = _list_batch_images(input_dir, preserve_structure)
    if not files:
        log_message(f"No image files found in '{input_dir}'", always_print=True)
        return empty
    mine = coord.shard(files)
    total = len(mine)
    t0 = time.time()
    if progress_callback:
        progress_callback(0.0, f"Starting batch processing of {total} images...")
    log_message(f"Starting batch processing: {len(files)} images" +
                (f" ({total} on rank {coord.rank} of {coord.world})" if coord.world > 1 else ""), always_print=True)
    res = {"success_count": 0, "error_count": 0, "errors": {}, "failed_image_paths": []}
    failed_jobs = []
    workers = int(os.environ.get("MTB200_SAVE_WORKERS", "2"))
    saver = _BoundedPool(workers, depth=2 * workers) if workers > 0 else None
    pending = []                                   # (future, key, path, shown) of queued saves

    def failed(key, path, shown, e):
        log_message(f"Error processing {shown}: {e}", always_print=True)
        src = resolve_source_path(path, source_path_map)
        res["error_count"] += 1
        res["errors"][key] = str(e)
        res["failed_image_paths"].append(src)
        failed_jobs.append((key, path, src))

    def settle(block: bool):
        """Book finished saves (all of them when `block`): a page counts as done once its file is written."""
        keep = []
        for fut, key, path, shown in pending:
            if not block and not fut.done():
                keep.append((fut, key, path, shown))
                continue
            try:
                fut.result()
                res["success_count"] += 1
            except Exception as e:
                failed(key, path, shown, e)
        pending[:] = keep

    decode = {}                                    # index -> pending decode of that page (one page ahead)

    def prefetch(j):
        if saver is not None and j < total and j not in decode:
            decode[j] = saver.submit_unbounded(_load_page_prepared, mine[j], config)

    try:
        for i, path in enumerate(mine):
            prefetch(i)
            prefetch(i + 1)                        # decoded while page i is on the device
            out_path, shown, key = resolve_output_path(path, input_dir, out_dir, config, preserve_structure)
            if cancellation_manager is not None and cancellation_manager.is_cancelled():
                raise CancellationError("Batch process cancelled by user.")
            if progress_callback:
                progress_callback(i / total, f"Processing image {i + 1}/{total}: {shown}")
            log_message(f"Processing {i + 1}/{total}: {shown}", always_print=True)
            try:
                fut = _process_page(path, config, out_path, cancellation_manager, saver, decode.pop(i, None))
                if fut is None:
                    res["success_count"] += 1
                else:
                    pending.append((fut, key, path, shown))
                note = f"Completed {i + 1}/{total} images"
            except CancellationError:
                raise
            except Exception as e:
                failed(key, path, shown, e)
                note = f"Completed {i + 1}/{total} images (with errors)"
            settle(block=False)
            if progress_callback:
                progress_callback((i + 1) / total, note)
        settle(block=True)
    except BaseException:
        # a rank that stops early (cancellation, a fatal error) still takes part in the gather below, or the other ranks
        # would wait in it forever; its partial counts carry an `aborted` mark
        res["aborted"] = 1
        if saver is not None:
            saver.close()
        try:
            coord.gather(res)
        finally:
            raise
    finally:
        if saver is not None:
            saver.close()                          # queued files are still written when the batch is cancelled
    cancelled = cancellation_manager is not None and cancellation_manager.is_cancelled()
    if getattr(config, "retry_failed_once", False) and failed_jobs and not cancelled:
        _retry_failed(failed_jobs, res, config, input_dir, out_dir, preserve_structure, progress_callback, cancellation_manager)
    if progress_callback:
        progress_callback(1.0, "Processing complete")
    parts = coord.gather(res)
    if coord.rank != 0 or parts is None:
        return res
    merged = {"success_count": 0, "error_count": 0, "errors": {}, "failed_image_paths": []}
    for part in parts:
        for k, v in part.items():
            if isinstance(v, int):
                merged[k] = merged.get(k, 0) + v
            elif isinstance(v, dict):
                merged.setdefault(k, {}).update(v)
            else:
                merged.setdefault(k, []).extend(v)
    dt = time.time() - t0
    log_message(f"Batch complete: {merged['success_count']}/{len(files)} images in {dt:.2f}s "
                f"({dt / len(files):.2f}s/image)", always_print=True)
    if merged["error_count"] > 0:
        log_message(f"Failed: {merged['error_count']} images", always_print=True)
        for name, msg in merged["errors"].items():
            log_message(f"  - {name}: {msg}", always_print=True)
    failed_file = write_failed_paths(out_dir, merged["failed_image_paths"])
    if failed_file:
        merged["failed_paths_file"] = str(failed_file)
    return merged


def _retry_failed(failed_jobs, res, config, input_dir, out_dir, preserve_structure, progress_callback, cancellation_manager):
    """One more attempt per failed page (pipeline.py:2081-2212); updates the counts, the error table and the failed
    path list in place and records retry_attempted_count / retry_success_count / retry_failed_count."""
    attempted = recovered = still = 0
    n = len(failed_jobs)
    log_message(f"Retrying {n} failed image(s) once...", always_print=True)
    for i, (key, path, src) in enumerate(failed_jobs):
        if cancellation_manager is not None and cancellation_manager.is_cancelled():
            break
        out_path, shown, _ = resolve_output_path(path, input_dir, out_dir, config, preserve_structure)
        if progress_callback:
            progress_callback(0.95 + 0.05 * (i / max(n, 1)), f"Retrying failed image {i + 1}/{n}: {shown}")
        attempted += 1
        try:
            translate_and_render(path, config, out_path, cancellation_manager=cancellation_manager)
        except CancellationError:
            attempted -= 1
            break
        except Exception as e:
            res["errors"][key] = str(e)
            still += 1
            log_message(f"Retry failed for {shown}: {e}", always_print=True)
            continue
        res["success_count"] += 1
        res["error_count"] = max(0, res["error_count"] - 1)
        res["errors"].pop(key, None)
        res["failed_image_paths"] = [p for p in res["failed_image_paths"] if p != src]
        recovered += 1
        log_message(f"Retry succeeded: {shown}", always_print=True)
    res["retry_attempted_count"], res["retry_success_count"], res["retry_failed_count"] = attempted, recovered, still


# ---- device-resident batch engine --------------------------------------------------------------------------------------
class HotPathPipeline:
    """detect -> segment -> clean -> upscale for a stream of pages, everything on the device.

    `run_page(page_bgr_u8_host_pinned)`: H2D copy, YOLO detect (+NMS/dedup on device), SAM 2.1 masks, bubble cleaning,
    RCAN 2x upscale, D2H of the uint8 result.  Per page there is one small D2H (the detection table) besides the final
    image."""

    def __init__(self, *, confidence: float = 0.6, imgsz: int = 1600, seg_model: str = "sam2", upscale: bool = True,
                 upscale_model: str = "model", thresholding_value: int = 200, roi_shrink_px: int = 5,
                 device: Optional[torch.device] = None, conjoined_detection: bool = False,
                 conjoined_confidence: float = 0.35, use_otsu_threshold: bool = False, yolo_model_path=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.confidence, self.imgsz, self.seg_model = confidence, imgsz, seg_model
        self.upscale, self.upscale_model = upscale, upscale_model
        self.thr, self.shrink, self.otsu = thresholding_value, roi_shrink_px, bool(use_otsu_threshold)
        # secondary RT-DETR detector merged like the reference's default (detection.py:1392-1548); off here by default
        # because without its checkpoint the loader refuses and every page would log the swallowed failure
        self.conjoined = dict(conjoined_detection=conjoined_detection, conjoined_confidence=conjoined_confidence)
        mm = get_model_manager()
        self.yolo = mm.load_yolo_speech_bubble(yolo_model_path or None)
        self.sam = mm.load_sam2() if seg_model == "sam2" else None
        self.rcan = (mm.load_upscale() if upscale_model == "model" else mm.load_upscale_lite()) if upscale else None
        self.stage_ms: Dict[str, float] = {}

    def run_page_device(self, page_bgr: torch.Tensor, injected_boxes: Optional[np.ndarray] = None,
                        timings: Optional[dict] = None):
        """page_bgr: device uint8 HxWx3.  Returns (cleaned [and upscaled] page uint8 on device in RGB order when
        upscaled else BGR, detections, clean results)."""
        ev = lambda: torch.cuda.Event(enable_timing=True)
        t = [ev() for _ in range(4)] if timings is not None else None
        if t: t[0].record()
        h, w = int(page_bgr.shape[0]), int(page_bgr.shape[1])
        dets = detect_pages_device([page_bgr], confidence=self.confidence, imgsz=self.imgsz, seg_model=self.seg_model,
                                   injected_boxes=None if injected_boxes is None else [injected_boxes], **self.conjoined)[0]
        if t: t[1].record()
        scale = _processing_scale(w, h)
        batch = clean_pages_device([page_bgr], [dets], thresholding_value=self.thr, roi_shrink_px=self.shrink,
                                   use_otsu_threshold=self.otsu, processing_scale=scale)
        cleaned = batch.pages_out[0]
        if t: t[2].record()
        out = cleaned
        if self.rcan is not None:
            out = self.rcan.upscale_u8(cleaned, swap_rb=True)        # BGR page -> RGB model input/output
        if t:
            t[3].record()
            torch.cuda.synchronize()
            for name, a, b in (("detect_segment", 0, 1), ("clean", 1, 2), ("upscale", 2, 3)):
                timings[name] = timings.get(name, 0.0) + t[a].elapsed_time(t[b])
        return out, dets, batch

    def run_pages_device(self, pages_bgr: Sequence[torch.Tensor],
                         injected_boxes: Optional[Sequence[Optional[np.ndarray]]] = None,
                         consume: Optional[Callable[[int, torch.Tensor], None]] = None,
                         timings: Optional[dict] = None):
        """A group of device pages: detect+segment page by page, ONE cleaning launch for every bubble of the group (the
        cleaning kernel runs one CTA per bubble, so a single page's ~12 bubbles leave most SMs idle), then the upscale
        page by page.  `consume(i, out)` is called as each page's result becomes available (the upscaler's output is a
        static buffer that the next page overwrites).  Returns (outputs or None when consumed, detections, clean batch)."""
        n = len(pages_bgr)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        t = [ev() for _ in range(4)] if timings is not None else None
        if t: t[0].record()
        if injected_boxes is not None and any(b is None for b in injected_boxes):
            dets = []
            for i, page in enumerate(pages_bgr):
                inj = None if injected_boxes[i] is None else [injected_boxes[i]]
                dets.append(detect_pages_device([page], confidence=self.confidence, imgsz=self.imgsz, seg_model=self.seg_model,
                                                injected_boxes=inj, own_masks=n > 1, **self.conjoined)[0])
        else:
            # the whole group in one call: every page's detector is enqueued before the host reads the box tables once
            dets = detect_pages_device(list(pages_bgr), confidence=self.confidence, imgsz=self.imgsz, seg_model=self.seg_model,
                                       injected_boxes=None if injected_boxes is None else list(injected_boxes),
                                       own_masks=n > 1, **self.conjoined)
        if t: t[1].record()
        scales = {_processing_scale(int(p.shape[1]), int(p.shape[0])) for p in pages_bgr}
        if len(scales) == 1:
            batch = clean_pages_device(list(pages_bgr), dets, thresholding_value=self.thr, roi_shrink_px=self.shrink,
                                       use_otsu_threshold=self.otsu, processing_scale=scales.pop())
            cleaned = batch.pages_out
        else:                                   # mixed page sizes: the scale-derived parameters differ per page
            batch, cleaned = None, []
            for page, d in zip(pages_bgr, dets):
                b1 = clean_pages_device([page], [d], thresholding_value=self.thr, roi_shrink_px=self.shrink,
                                        use_otsu_threshold=self.otsu, processing_scale=_processing_scale(int(page.shape[1]), int(page.shape[0])))
                cleaned.append(b1.pages_out[0])
        if t: t[2].record()
        outs = []
        for i in range(n):
            out = cleaned[i]
            if self.rcan is not None:
                out = self.rcan.upscale_u8(cleaned[i], swap_rb=True)
            if consume is not None:
                consume(i, out)
            else:
                outs.append(out.clone() if self.rcan is not None else out)
        if t:                                        # per-page stage times of the group, as the batch path runs them
            t[3].record()
            torch.cuda.synchronize()
            for name, a, b in (("detect_segment", 0, 1), ("clean_grouped", 1, 2), ("upscale", 2, 3)):
                timings[name] = timings.get(name, 0.0) + t[a].elapsed_time(t[b]) / max(n, 1)
        return (None if consume is not None else outs), dets, batch

    def run_pages(self, pages_bgr_host: Sequence[torch.Tensor], outs_host: Optional[Sequence[torch.Tensor]] = None,
                  injected_boxes: Optional[Sequence[Optional[np.ndarray]]] = None):
        """Host (pinned) uint8 pages in, host uint8 pages out, as a group (see run_pages_device)."""
        pages = [p.to(self.device, non_blocking=True) for p in pages_bgr_host]
        results: List[Optional[torch.Tensor]] = [None] * len(pages)

        def consume(i: int, out: torch.Tensor) -> None:
            dst = outs_host[i] if outs_host is not None else torch.empty(out.shape, dtype=torch.uint8, pin_memory=True)
            dst.copy_(out, non_blocking=True)      # stream-ordered: done before the next page overwrites `out`
            results[i] = dst

        _, dets, batch = self.run_pages_device(pages, injected_boxes, consume)
        torch.cuda.current_stream().synchronize()
        return results, dets, batch

    def run_page(self, page_bgr_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, **kw):
        """Host (pinned) uint8 HxWx3 in, host uint8 out (the call a user of the stage API makes)."""
        page = page_bgr_host.to(self.device, non_blocking=True)
        out, dets, batch = self.run_page_device(page, **kw)
        if out_host is None:
            out_host = torch.empty(out.shape, dtype=torch.uint8, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host, dets, batch
