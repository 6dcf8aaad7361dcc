"""CPU: host-side logic of the boundary — parameter scaling vs the reference, box post-processing vs the reference,
page sharding over torch.distributed (gloo, world size 2), drop-in aliasing."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

H_ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))


def test_scaling_matches_reference_when_present():
    import _refimport
    from mangatranslator_b200.core import scaling as ours
    if not _refimport.available():
        pytest.skip("reference tree not present")
    _refimport.import_reference()
    import core.scaling as ref
    for s in [None, 0.0, 0.3, 0.8868, 1.0, 1.2541, 2.0, 3.7, 9.5, 40.0]:
        for k in [(7, 7), (5, 5), (3, 9)]:
            assert ours.scale_kernel(k, s) == ref.scale_kernel(k, s)
        assert ours.scale_area(50, s, minimum=50, maximum=5000) == ref.scale_area(50, s, minimum=50, maximum=5000)
        assert ours.scale_scalar(5, s, minimum=0.0, maximum=64.0) == ref.scale_scalar(5, s, minimum=0.0, maximum=64.0)
        assert ours.scale_length(12, s) == ref.scale_length(12, s)


def test_scaling_known_values():
    from mangatranslator_b200 import clean_host as H
    p = H.build_params(200, False, 5, (1536 * 1024 / 1e6) ** 0.5)
    assert (p.kd, p.ke, p.min_area) == (9, 7, 79.0)
    p = H.build_params(200, False, 5, (1024 * 768 / 1e6) ** 0.5)
    assert (p.kd, p.ke, p.min_area) == (7, 5, 50.0)


def test_box_postprocessing_matches_reference():
    import _refimport
    from mangatranslator_b200.core.image import detection as ours
    if not _refimport.available():
        pytest.skip("reference tree not present")
    core = _refimport.import_reference()
    import core.image.detection as ref
    rng = np.random.default_rng(1)
    for _ in range(30):
        n = int(rng.integers(1, 30))
        xy = rng.uniform(0, 800, size=(n, 2))
        wh = rng.uniform(10, 300, size=(n, 2))
        b = torch.from_numpy(np.concatenate([xy, xy + wh], 1).astype(np.float32))
        if n > 3:
            b[1] = b[0] + 2.0
            b[2, :2] = b[0, :2] + 5
            b[2, 2:] = b[0, 2:] - 5
        c = torch.from_numpy(rng.uniform(0.3, 1.0, size=n).astype(np.float32))
        rb, rk = ref._deduplicate_primary_boxes(b, c, 0.7)
        ob, ok = ours._deduplicate_primary_boxes(b, c, 0.7)
        assert rk == ok and torch.equal(rb, ob)
        rb2, ri = ref._remove_contained_boxes(rb, [("primary", i) for i in rk])
        ob2, oi = ours._remove_contained_boxes(ob, [("primary", i) for i in ok])
        assert ri == oi and torch.equal(rb2, ob2)
        m1 = ref._build_rect_mask_from_box(b[0], 600, 700)
        m2 = ours._build_rect_mask_from_box(b[0], 600, 700)
        assert np.array_equal(m1, m2)


def test_batch_coordinator_helpers():
    from mangatranslator_b200.core import batch_coordinator as bc
    c = bc.BatchRequestCoordinator(2)
    assert c.map_ordered([lambda i=i: i * i for i in range(6)]) == [0, 1, 4, 9, 16, 25]
    with c.slot():
        assert c.in_slot()
        with c.slot():
            pass
    assert bc.bboxes_overlap((0, 0, 10, 10), (5, 5, 20, 20)) and not bc.bboxes_overlap((0, 0, 10, 10), (10, 0, 20, 10))
    m = np.zeros((100, 200), np.uint8)
    m[40:60, 50:90] = 1
    assert bc.expanded_mask_bbox(m, (200, 100)) == (0, 0, 170, 100)
    waves = bc.partition_non_overlapping_waves([(0, 0, 10, 10), (20, 0, 30, 10), (5, 5, 25, 25), None], lambda b: b)
    assert [len(w) for w in waves] == [2, 1, 1]


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
from mangatranslator_b200.core.batch_coordinator import PageShardCoordinator
c = PageShardCoordinator(backend="gloo")
pages = list(range(11))
mine = c.shard(pages)
assert all(c.owner_of(i) == c.rank for i in mine)
c.barrier()
got = c.gather(dict(rank=c.rank, pages=mine))
mx = c.all_reduce_max(float(c.rank + 1))
assert mx == float(c.world)
if c.rank == 0:
    allp = sorted(p for g in got for p in g["pages"])
    assert allp == pages, allp
    print("SHARD_OK", [g["pages"] for g in got])
c.close()
"""


def test_page_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHARD_OK [[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9]]" in out.stdout


def test_drop_in_aliases_share_singletons():
    import mangatranslator_b200.drop_in as d
    saved = {k: sys.modules.get(k) for k, _ in d._MODULES}
    try:
        for k, _ in d._MODULES:
            sys.modules.pop(k, None)
        d.install()
        import core.ml.model_manager as a
        import mangatranslator_b200.core.ml.model_manager as b
        assert a is b and a.get_model_manager() is b.get_model_manager()
        from core.image.cleaning import clean_speech_bubbles  # noqa: F401
        from utils.exceptions import CleaningError  # noqa: F401
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_cache_keys_follow_the_image_content():
    """Keys are content fingerprints like the reference's (core/caching.py:28-50): an image mutated in place misses, an equal
    copy hits, and every kind of result keeps only a couple of entries (no page pairs pile up in host memory)."""
    from PIL import Image
    from mangatranslator_b200.core import caching
    c = caching.UnifiedCache()
    a = Image.new("RGB", (4, 4), (1, 2, 3))
    c.set_upscaled_image(c.get_bubble_processing_cache_key(a, 200, "min", "model_lite"), "result-a")
    assert c.get_upscaled_image(c.get_bubble_processing_cache_key(a, 200, "min", "model_lite")) == "result-a"
    assert c.get_upscaled_image(c.get_bubble_processing_cache_key(a.copy(), 200, "min", "model_lite")) == "result-a"
    assert c.get_upscaled_image(c.get_bubble_processing_cache_key(a, 201, "min", "model_lite")) is None
    a.putpixel((1, 1), (9, 9, 9))                      # in-place edit (the reference pastes / inpaints, then re-upscales)
    assert c.get_upscaled_image(c.get_bubble_processing_cache_key(a, 200, "min", "model_lite")) is None
    for i in range(10):
        c.set_upscaled_image(c.get_upscale_cache_key(Image.new("RGB", (2, 2), (i, 0, 0)), 2.0, "model"), i)
    assert len(c._store["upscale"]) == caching.MAX_ENTRIES and len(c) == caching.MAX_ENTRIES + 1
    # a new page drops what was cached for the previous one; the same page (equal content) keeps it
    page = Image.new("RGB", (8, 8), (5, 5, 5))
    c.set_current_image(page)
    c.set_yolo_detection(c.get_yolo_cache_key(page, "m.pt", 0.6), "dets")
    c.set_current_image(page.copy())
    assert c.get_yolo_detection(c.get_yolo_cache_key(page, "m.pt", 0.6)) == "dets"
    c.set_current_image(Image.new("RGB", (8, 8), (6, 5, 5)))
    assert len(c) == 0


def test_yolo_oracle_nms_is_pinned_to_torchvision():
    """ultralytics' non_max_suppression ends in `torchvision.ops.nms` (present in this image even though ultralytics is
    not): the oracle's restated greedy NMS must return exactly the library's indices, on random boxes with heavy overlap,
    class offsets and exact score ties."""
    import torch
    import torchvision
    import yolo_oracle
    g = torch.Generator().manual_seed(0)
    for trial in range(40):
        n = int(torch.randint(1, 400, (1,), generator=g))
        centers = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 120 + 4
        boxes = torch.cat([centers - wh / 2, centers + wh / 2], 1)
        if trial % 3 == 0:
            boxes = boxes + (torch.randint(0, 3, (n, 1), generator=g).float() * 7680)      # class offsets like ultralytics
        scores = torch.rand(n, generator=g)
        if trial % 4 == 1:
            scores = (scores * 8).round() / 8                                              # many exact ties
        order = scores.argsort(descending=True, stable=True)                               # what non_max_suppression feeds
        b, s = boxes[order], scores[order]
        for thr in (0.45, 0.7):
            assert torch.equal(yolo_oracle.nms(b, s, thr), torchvision.ops.nms(b, s, thr))


def test_core_package_exports_the_reference_names_lazily():
    """`from core import ...` (core/__init__.py:8-41): hot-path names resolve to this build's objects, names of
    subsystems outside the build raise AttributeError, and importing the package alone stays cheap."""
    import importlib
    import subprocess
    import sys
    import mangatranslator_b200.core as core
    from mangatranslator_b200.core.caching import UnifiedCache, get_cache
    from mangatranslator_b200.core.image.cleaning import clean_speech_bubbles
    from mangatranslator_b200.core.pipeline import batch_translate_images
    assert core.get_cache is get_cache and core.UnifiedCache is UnifiedCache
    assert core.clean_speech_bubbles is clean_speech_bubbles and core.batch_translate_images is batch_translate_images
    assert core.__version_info__ == (1, 22, 2) and core.__version__.startswith("1.22.2")
    for name in ("render_text_skia", "call_translation_api_batch", "FluxKontextInpainter", "OutsideTextDetector"):
        try:
            getattr(core, name)
            raised = False
        except AttributeError as e:
            raised = "outside the B200 hot path" in str(e)
        assert raised, name
    code = "import sys; import mangatranslator_b200.core; print('torch' in sys.modules)"
    out = subprocess.check_output([sys.executable, "-c", code], cwd=H_ROOT).decode().strip()
    assert out == "False"
    assert importlib.import_module("mangatranslator_b200.core._version").__version__ == core.__version__


def test_model_manager_lifecycle_follows_the_reference_conventions():
    """unload_ocr_models releases what the reference calls OCR-related (detectors, SAM, manga-ocr: model_manager.py
    :1397-1432) and keeps the upscalers; a slot left as None (the reference's state after an unload, :1384-1385) counts as
    not loaded."""
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    mm = get_model_manager()
    saved = dict(mm.models)
    try:
        mm.models.clear()
        for t in (ModelType.YOLO_SPEECH_BUBBLE, ModelType.YOLO_SPEECH_BUBBLE_2, ModelType.RTDETR_CONJOINED_BUBBLE,
                  ModelType.SAM2, ModelType.MANGA_OCR, ModelType.UPSCALE, ModelType.UPSCALE_LITE):
            mm.models[t] = object()
        mm.unload_ocr_models()
        assert sorted(t.name for t in mm.models if mm.is_loaded(t)) == ["UPSCALE", "UPSCALE_LITE"]
        mm.models[ModelType.UPSCALE] = None
        assert not mm.is_loaded(ModelType.UPSCALE) and mm.is_loaded(ModelType.UPSCALE_LITE)
        stats = mm.get_memory_stats()
        assert stats["loaded_models"] == ["upscale_lite"]
        assert ("allocated_gb" in stats and "reserved_gb" in stats) or stats.get("memory") == "N/A"   # core/device.py:116-172
        mm.print_memory_stats()
        mm.unload_upscale_models()
        assert not any(mm.is_loaded(t) for t in ModelType)
    finally:
        mm.models.clear()
        mm.models.update(saved)
