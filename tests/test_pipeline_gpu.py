"""GPU: the whole hot path (detect -> segment -> clean -> upscale) through the reference-shaped stage functions and the
device-resident HotPathPipeline, against the CPU oracle pipeline with the same seeded weights."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_models(monkeypatch_module=None):
    """Small synthetic models registered in the ModelManager the way a user of the reference would inject objects."""
    import os
    from mangatranslator_b200 import weights as W
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    from mangatranslator_b200.rcan import RcanB200
    from mangatranslator_b200.sam2 import Sam2B200
    from mangatranslator_b200.sam2_api import Sam2ModelB200, Sam2ProcessorB200
    from mangatranslator_b200.yolo import YoloB200
    mm = get_model_manager()
    mm.unload_all()
    dev = torch.device("cuda:0")
    ycfg = W.yolo_cfg("n")
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = YoloB200(W.yolo_state_dict(0, ycfg), ycfg, dev)
    cfg, sd = W.sam2_model_and_state(0)
    net = Sam2B200(sd, cfg, dev)
    mm.models[ModelType.SAM2] = (Sam2ProcessorB200(net), Sam2ModelB200(net))
    mm.models[ModelType.UPSCALE] = RcanB200(W.rcan_state_dict(0, n_resgroups=2, n_resblocks=2), dev)
    yield mm, ycfg, cfg, sd
    mm.unload_all()


def test_hot_path_pipeline_matches_cpu_pipeline(small_models):
    import pipeline_oracle
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.pipeline import HotPathPipeline
    mm, ycfg, cfg, sd = small_models
    h, w = 448, 384
    pg = synth.make_page(21, h, w, n_bubbles=4)
    cpu = pipeline_oracle.CpuPipeline(0, yolo_variant="n", rcan_groups=2, rcan_blocks=2)
    ref = cpu.run_page(pg.image_rgb, pg.boxes_xyxy, imgsz=640)
    pipe = HotPathPipeline(seg_model="sam2", upscale=True, imgsz=640)
    host = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).pin_memory()
    out, dets, batch = pipe.run_page(host, injected_boxes=pg.boxes_xyxy)
    # masks: bit-exact outside the oracle's knife-edge band, flipped bits counted; identical cleaning decisions (same
    # bubbles processed, same fill colours); the RCAN within 1e-3 abs on the oracle's cleaned page (helpers)
    from helpers import check_page_against_cpu_pipeline
    ok = [r for r in batch.results[0] if r is not None and r.status == 0]
    assert len(ok) == len(ref["bubbles"])
    for r, b in zip(ok, ref["bubbles"]):
        assert tuple(r.fill_bgr) == tuple(b["color"])
    assert tuple(out.shape) == (2 * h, 2 * w, 3)
    check_page_against_cpu_pipeline(pipe, ref, out, dets, batch, "448x384 page, 4 bubbles")


def test_grouped_pages_equal_page_by_page(small_models):
    """run_pages (one cleaning launch for a group, masks copied out of the segmenter's static buffer) must return
    byte-identical pages to run_page called page by page."""
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.pipeline import HotPathPipeline
    h, w = 448, 384
    pages = [synth.make_page(30 + i, h, w, n_bubbles=3 + i) for i in range(3)]
    pipe = HotPathPipeline(seg_model="sam2", upscale=True, imgsz=640)
    hosts = [torch.from_numpy(np.ascontiguousarray(p.image_rgb[:, :, ::-1])).pin_memory() for p in pages]
    single = [pipe.run_page(hh, injected_boxes=p.boxes_xyxy)[0].clone() for hh, p in zip(hosts, pages)]
    outs, dets, batch = pipe.run_pages(hosts, None, [p.boxes_xyxy for p in pages])
    assert len(outs) == 3 and len(dets) == 3
    for a, b in zip(single, outs):
        assert torch.equal(a, b)


def test_reference_shaped_stage_functions(small_models, tmp_path):
    """detect_speech_bubbles / clean_speech_bubbles / upscale_image / translate_and_render with the reference's
    signatures (cleaning_only + upscale_final_image), fake detector injected through ModelManager.models exactly like
    SURVEY.md §8c did with the reference."""
    from types import SimpleNamespace
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.config import MangaTranslatorConfig
    from mangatranslator_b200.core.image.detection import detect_speech_bubbles
    from mangatranslator_b200.core.ml.model_manager import ModelType
    from mangatranslator_b200.core.pipeline import translate_and_render
    mm, *_ = small_models
    h, w = 400, 352
    pg = synth.make_page(33, h, w, n_bubbles=3)
    boxes = torch.from_numpy(pg.boxes_xyxy)
    dup = torch.cat([boxes, boxes[:1] + 1.0])                        # one near-duplicate the dedup must remove

    class FakeBoxes:
        def __init__(self):
            self.xyxy, self.conf, self.cls = dup.clone(), torch.tensor([0.9, 0.8, 0.7, 0.65]), torch.zeros(4)

        def __len__(self):
            return 4

    class FakeYolo:
        names = {0: "speech_bubble"}

        def __call__(self, im, conf, device, verbose, imgsz, retina_masks):
            assert imgsz == 1600 and retina_masks
            return [SimpleNamespace(boxes=FakeBoxes(), masks=None, orig_shape=im.shape[:2], names=self.names)]

    real = mm.models[ModelType.YOLO_SPEECH_BUBBLE]
    mm.models[ModelType.YOLO_SPEECH_BUBBLE] = FakeYolo()
    try:
        pil = Image.fromarray(pg.image_rgb)
        dets, free = detect_speech_bubbles(tmp_path / "x.png", "x.pt", 0.6, seg_model="sam2", conjoined_detection=True,
                                           image_override=pil)
        assert len(dets) == 3 and free == []
        for d, b in zip(dets, pg.boxes_xyxy):
            assert d["bbox"] == tuple(int(round(float(v))) for v in b)
            assert d["sam_mask"].shape == (h, w) and d["sam_mask"].dtype == np.uint8
            assert set(np.unique(d["sam_mask"])) <= {0, 255}
        p = tmp_path / "page.png"
        pil.save(p)
        cfg = MangaTranslatorConfig(cleaning_only=True)
        cfg.detection.seg_model = "sam2"
        cfg.output.upscale_final_image = True
        cfg.output.image_upscale_model = "model"
        out = translate_and_render(p, cfg, tmp_path / "out.png")
        assert out.size == (2 * w, 2 * h) and (tmp_path / "out.png").exists()
    finally:
        mm.models[ModelType.YOLO_SPEECH_BUBBLE] = real


def test_stage_functions_are_safe_under_page_threads(small_models):
    """The reference's batch mode calls the same model objects from several page threads (core/pipeline.py:2470); the
    B200 models keep static buffers per input size, so the stage functions serialise on one device section: concurrent
    callers must get exactly what sequential callers get."""
    import threading
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.caching import get_cache
    from mangatranslator_b200.core.image.cleaning import clean_speech_bubbles
    from mangatranslator_b200.core.image.image_utils import upscale_image
    pages = [synth.make_page(40 + i, 320 + 32 * (i % 2), 288, n_bubbles=3) for i in range(4)]

    def work(pg):
        pil = Image.fromarray(pg.image_rgb)
        cleaned, info = clean_speech_bubbles(pil, "x.pt", pre_computed_detections=synth.detections_from_page(pg),
                                             processing_scale=1.0)
        get_cache().clear()
        up = upscale_image(Image.fromarray(np.ascontiguousarray(cleaned[:, :, ::-1])), 2.0, model_type="model")
        return cleaned, np.asarray(up), len(info)

    expected = [work(p) for p in pages]
    got = [None] * len(pages)
    errors = []

    def runner(i):
        try:
            for _ in range(3):
                got[i] = work(pages[i])
        except Exception as e:                                   # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=runner, args=(i,)) for i in range(len(pages))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (c0, u0, n0), (c1, u1, n1) in zip(expected, got):
        assert n0 == n1 and np.array_equal(c0, c1) and np.array_equal(u0, u1)


def test_hot_path_with_conjoined_group_matches_cpu_pipeline(small_models):
    """A page whose injected boxes overlap (two of them form a synthetic conjoined group): the device pipeline and the CPU
    pipeline (transformers SAM on the union box, oracle split, cv2 clean with neighbour boxes, oracle RCAN) agree like the
    plain pages do — SAM knife-edge pixels aside."""
    import pipeline_oracle
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.pipeline import HotPathPipeline
    h, w = 448, 384
    pg = synth.make_page(22, h, w, n_bubbles=4)
    boxes = np.asarray(pg.boxes_xyxy, np.float32).copy()
    boxes[1] = boxes[0] + np.array([0.45, 0.1, 0.45, 0.1], np.float32) * (boxes[0, 2] - boxes[0, 0])   # overlaps box 0
    boxes[1, [0, 2]] = np.clip(boxes[1, [0, 2]], 0, w)
    cpu = pipeline_oracle.CpuPipeline(0, yolo_variant="n", rcan_groups=2, rcan_blocks=2)
    ref = cpu.run_page(pg.image_rgb, boxes, imgsz=640)
    pipe = HotPathPipeline(seg_model="sam2", upscale=True, imgsz=640)
    host = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).pin_memory()
    out, dets, batch = pipe.run_page(host, injected_boxes=boxes)
    assert sum(1 for d in dets if d.get("conjoined_neighbor_bboxes")) == 2
    assert [d["bbox"] for d in dets] == [b["bbox"] for b in ref["bubbles"]] or len(dets) == ref["masks"].shape[0]
    from helpers import check_page_against_cpu_pipeline
    check_page_against_cpu_pipeline(pipe, ref, out, dets, batch, "448x384 page with a conjoined group")


@pytest.mark.parametrize("final,factor", [(True, 2.0), (False, 2.0), (True, 1.5)], ids=["upscale_2x", "clean_only", "upscale_1p5x"])
def test_batch_fast_path_equals_stage_functions(small_models, tmp_path, monkeypatch, final, factor):
    """`translate_and_render` sends eligible cleaning_only pages through the device-resident engine (one upload, one
    download); the result must be byte-identical to the reference-shaped stage functions (detect_speech_bubbles ->
    clean_speech_bubbles -> upscale_image) it replaces, for opaque RGB and RGBA sources."""
    from mangatranslator_b200 import synth
    from mangatranslator_b200.core.config import MangaTranslatorConfig
    from mangatranslator_b200.core import pipeline as P
    pg = synth.make_page(51, 416, 352, n_bubbles=4)
    cfg = MangaTranslatorConfig(cleaning_only=True)
    cfg.detection.seg_model, cfg.detection.conjoined_detection, cfg.detection.confidence = "sam2", False, 0.25
    cfg.output.upscale_final_image, cfg.output.image_upscale_factor, cfg.output.image_upscale_model = final, factor, "model"
    for mode in ("RGB", "RGBA"):
        src = tmp_path / f"page_{mode}.png"
        Image.fromarray(pg.image_rgb).convert(mode).save(src)
        monkeypatch.setenv("MTB200_FAST_PATH", "0")
        slow = P.translate_and_render(src, cfg, None)
        monkeypatch.setenv("MTB200_FAST_PATH", "1")
        P._FAST.clear()
        fast = P.translate_and_render(src, cfg, None)
        assert P._FAST, "the fast path was not taken"
        assert fast.size == slow.size
        assert np.array_equal(np.asarray(fast.convert("RGB")), np.asarray(slow.convert("RGB"))), (mode, final, factor)
