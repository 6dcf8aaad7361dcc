"""CPU: host side of the page driver (core/pipeline.py) against the reference's own helpers (live, when the checkout is
present): natural page order, output naming, source-path mapping, failed-path file, target mode, save rules, config
defaults; and the batch flow itself (counts, error keys, retry pass, cancellation, sharding over two gloo ranks) with
the render and save halves of a page replaced by stand-ins (the real render needs a GPU)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

import _refimport
from helpers import ROOT
from mangatranslator_b200.core import pipeline as P
from mangatranslator_b200.core.config import (CleaningConfig, DetectionConfig, MangaTranslatorConfig, OutputConfig,
                                              PreprocessingConfig)

needs_ref = pytest.mark.skipif(not _refimport.available(), reason="reference checkout not present (GPU box)")

NAMES = ["10.png", "2.png", "page_2.jpg", "page_10.jpg", "Page_3.jpg", "a/1.png", "a/10.png", "B/2.png", "b/2.png",
         "001.png", "1.png", "x9y10.webp", "x9y2.webp", "ch2/p1.png", "ch10/p1.png", "ch2/p01.png", "é.png", "Z.PNG"]


def _ref():
    _refimport.import_reference()
    import core.pipeline as RP
    import utils.path_list as RL
    return RP, RL


def test_natural_order_known_cases():
    order = sorted(["10.png", "2.png", "1.png", "page_10.jpg", "page_2.jpg", "Page_3.jpg"], key=lambda n: P.natural_path_key(Path(n)))
    # the raw token breaks case-insensitive ties before the next token is looked at: "Page_" < "page_"
    assert order == ["1.png", "2.png", "10.png", "Page_3.jpg", "page_2.jpg", "page_10.jpg"]


@needs_ref
def test_natural_order_and_output_naming_match_live_reference(tmp_path):
    RP, RL = _ref()
    paths = [Path(n) for n in NAMES]
    assert sorted(paths, key=P.natural_path_key) == sorted(paths, key=RP._natural_path_sort_key)
    rng = np.random.default_rng(0)
    rand = [Path("/".join("".join(rng.choice(list("aAbB019_ ."), size=rng.integers(1, 7))) for _ in range(rng.integers(1, 4))))
            for _ in range(300)]
    rand = [p for p in rand if all(part not in ("", ".", "..") for part in p.parts)]
    assert sorted(rand, key=P.natural_path_key) == sorted(rand, key=RP._natural_path_sort_key)
    inp, out = tmp_path / "in", tmp_path / "out"
    for fmt in ("png", "jpeg", "auto", "tiff"):
        cfg = MangaTranslatorConfig()
        cfg.output.output_format = fmt
        for n in NAMES:
            for keep in (False, True):
                ours = P.resolve_output_path(inp / n, inp, out, cfg, keep)
                theirs = RP._resolve_output_path(inp / n, inp, out, cfg, keep)
                assert ours == theirs, (fmt, n, keep)
    # source path mapping and the failed-path file
    f = tmp_path / "in" / "x.png"
    f.parent.mkdir(parents=True, exist_ok=True)
    f.write_bytes(b"")
    for m in (None, {}, {str(f.resolve()): "/orig/x.png"}, {str(f): "/orig/y.png"}, {"other": "z"}):
        assert P.resolve_source_path(f, m) == RL.resolve_source_path(f, m)
    lists = [[], [None, "", "  "], [str(f), str(f), " " + str(f) + " ", "rel/a.png", None, "/abs/b.png"]]
    for i, lst in enumerate(lists):
        a, b = P.write_failed_paths(tmp_path / f"o{i}", lst), RL.write_failed_paths(tmp_path / f"r{i}", lst)
        assert (a is None) == (b is None)
        if a is not None:
            assert a.name == b.name and a.read_text() == b.read_text()


@needs_ref
def test_config_defaults_match_live_reference():
    _refimport.import_reference()
    import dataclasses
    import core.config as RC
    for ours, theirs in ((DetectionConfig, RC.DetectionConfig), (CleaningConfig, RC.CleaningConfig),
                         (OutputConfig, RC.OutputConfig), (PreprocessingConfig, RC.PreprocessingConfig)):
        mine = {f.name: f.default for f in dataclasses.fields(ours)}
        ref = {f.name: f.default for f in dataclasses.fields(theirs)}
        assert mine == ref, (ours.__name__, mine, ref)
    ref_top = {f.name: f.default for f in dataclasses.fields(RC.MangaTranslatorConfig) if f.default is not dataclasses.MISSING}
    mine_top = {f.name: f.default for f in dataclasses.fields(MangaTranslatorConfig) if f.default is not dataclasses.MISSING}
    for name, val in mine_top.items():
        if name in ref_top:
            assert ref_top[name] == val, name
    assert {"retry_failed_once", "processing_scale", "parallel_requests", "request_coordinator", "cleaning_only",
            "upscaling_only"} <= set(mine_top)


def test_target_mode_rule():
    cfg = MangaTranslatorConfig()
    table = [("png", "a.jpg", None, "RGBA"), ("jpeg", "a.png", None, "RGB"), ("auto", "a.jpg", None, "RGB"),
             ("auto", "a.png", "o.jpeg", "RGB"), ("auto", "a.jpg", "o.png", "RGBA"), ("auto", "a.webp", None, "RGBA")]
    for fmt, src, out, mode in table:
        cfg.output.output_format = fmt
        assert P._target_mode(cfg, Path(src), out) == mode, (fmt, src, out)


def test_save_image_rules(tmp_path):
    rgba = Image.new("RGBA", (8, 6), (10, 200, 30, 0))
    rgba.putpixel((1, 1), (10, 200, 30, 255))
    p = P.save_image_with_compression(rgba, tmp_path / "a.jpg", jpeg_quality=500)
    back = Image.open(p)
    assert back.mode == "RGB" and back.getpixel((6, 4))[0] > 240          # transparent pixels land on white
    p = P.save_image_with_compression(rgba, tmp_path / "sub" / "b.png", png_compression=99)
    assert np.array_equal(np.asarray(Image.open(p)), np.asarray(rgba))     # lossless
    p = P.save_image_with_compression(rgba.convert("RGB"), tmp_path / "c.webp")
    assert np.array_equal(np.asarray(Image.open(p).convert("RGB")), np.asarray(rgba.convert("RGB")))
    p = P.save_image_with_compression(rgba, tmp_path / "d.tiff")
    assert p.suffix == ".png" and p.exists()
    pal = Image.new("P", (4, 4))
    assert Image.open(P.save_image_with_compression(pal, tmp_path / "e.jpeg")).mode == "RGB"


class _Cancel:
    def __init__(self, after):
        self.n, self.after = 0, after

    def is_cancelled(self):
        self.n += 1
        return self.n > self.after


def _make_dir(tmp_path):
    inp = tmp_path / "in"
    (inp / "ch2").mkdir(parents=True)
    for n in ["10.png", "2.png", "1.jpg", "bad.png", "notes.txt", "ch2/3.png", "x.bmp"]:
        (inp / n).write_bytes(b"x")
    return inp


@pytest.mark.parametrize("workers", ["0", "2"], ids=["inline_save", "writer_pool"])
def test_batch_flow_counts_errors_retry_and_cancellation(tmp_path, monkeypatch, workers):
    """The render half (device work) and the save half are replaced by stand-ins; `workers` selects saving inline
    (translate_and_render as a whole) or on the writer pool."""
    monkeypatch.setenv("MTB200_SAVE_WORKERS", workers)
    inp = _make_dir(tmp_path)
    seen, fail_once = [], {"bad.png": 1, "stage": "render"}

    def fake_render(path, config, output_path=None, cancellation_manager=None, preloaded=None, device_png=False):
        seen.append((Path(path).name, Path(output_path).name))
        if Path(path).name == "bad.png" and fail_once["stage"] == "render" and fail_once["bad.png"] != 0:
            fail_once["bad.png"] -= 1
            raise RuntimeError("boom")
        return Path(path).name, "RGB"

    def fake_save(image, target_mode, output_path, config):
        if image == "bad.png" and fail_once["stage"] == "save" and fail_once["bad.png"] != 0:
            fail_once["bad.png"] -= 1
            raise RuntimeError("disk full")
        Path(output_path).write_bytes(b"ok")

    monkeypatch.setattr(P, "_render_page", fake_render)
    monkeypatch.setattr(P, "_save_page", fake_save)
    cfg = MangaTranslatorConfig(cleaning_only=True)
    prog = []
    res = P.batch_translate_images(inp, cfg, tmp_path / "out", progress_callback=lambda f, m: prog.append(f),
                                   source_path_map={str((inp / "bad.png").resolve()): "/orig/bad.png"})
    # root only, natural order, .bmp / .txt ignored, <stem>_translated.png naming (output_format default "png")
    assert seen == [("1.jpg", "1_translated.png"), ("2.png", "2_translated.png"), ("10.png", "10_translated.png"),
                    ("bad.png", "bad_translated.png")]
    assert res["success_count"] == 3 and res["error_count"] == 1 and res["errors"] == {"bad.png": "boom"}
    assert res["failed_image_paths"] == ["/orig/bad.png"]
    assert Path(res["failed_paths_file"]).read_text() == "/orig/bad.png\n"
    assert prog[0] == 0.0 and prog[-1] == 1.0 and all(0 <= f <= 1 for f in prog)
    assert sorted(p.name for p in (tmp_path / "out").glob("*_translated.png")) == ["10_translated.png", "1_translated.png",
                                                                                  "2_translated.png"]
    # a failure while SAVING is booked the same way, whichever thread wrote the file
    fail_once.update({"bad.png": 1, "stage": "save"})
    res = P.batch_translate_images(inp, cfg, tmp_path / "out_s")
    assert res["success_count"] == 3 and res["error_count"] == 1 and res["errors"] == {"bad.png": "disk full"}
    # with retry_failed_once the page recovers and the failure bookkeeping is undone
    seen.clear()
    fail_once.update({"bad.png": 1, "stage": "render"})
    cfg.retry_failed_once = True
    res = P.batch_translate_images(inp, cfg, tmp_path / "out2", preserve_structure=True)
    assert [s[0] for s in seen] == ["1.jpg", "2.png", "10.png", "bad.png", "3.png", "bad.png"]
    assert (tmp_path / "out2" / "ch2" / "3_translated.png").exists()
    assert res["success_count"] == 5 and res["error_count"] == 0 and res["errors"] == {} and res["failed_image_paths"] == []
    assert (res["retry_attempted_count"], res["retry_success_count"], res["retry_failed_count"]) == (1, 1, 0)
    assert "failed_paths_file" not in res
    # a permanently failing page stays failed after the retry
    fail_once["bad.png"] = -1
    res = P.batch_translate_images(inp, cfg, tmp_path / "out3")
    assert res["error_count"] == 1 and res["retry_failed_count"] == 1 and res["errors"] == {"bad.png": "boom"}
    # cancellation propagates (pipeline.py:2601-2602); pages already rendered are still written
    from mangatranslator_b200.utils.exceptions import CancellationError
    fail_once["bad.png"] = 0
    with pytest.raises(CancellationError):
        P.batch_translate_images(inp, cfg, tmp_path / "out4", cancellation_manager=_Cancel(after=2))
    assert sorted(p.name for p in (tmp_path / "out4").glob("*")) == ["1_translated.png", "2_translated.png"]
    # not a directory / nothing to do
    empty = {"success_count": 0, "error_count": 0, "errors": {}, "failed_image_paths": []}
    assert P.batch_translate_images(inp / "nope", cfg, tmp_path / "o5") == empty
    (tmp_path / "void").mkdir()
    assert P.batch_translate_images(tmp_path / "void", cfg, tmp_path / "o6") == empty


def test_writer_pool_overlaps_saving_with_the_next_pages_device_work(tmp_path, monkeypatch):
    import time
    inp = tmp_path / "in"
    inp.mkdir()
    for i in range(8):
        (inp / f"{i}.png").write_bytes(b"x")
    monkeypatch.setattr(P, "_render_page", lambda path, config, output_path=None, cancellation_manager=None, preloaded=None, device_png=False:
                        (time.sleep(0.05), ("img", "RGB"))[1])
    monkeypatch.setattr(P, "_save_page", lambda image, mode, out, config: (time.sleep(0.15), Path(out).write_bytes(b"ok"))[1])
    cfg = MangaTranslatorConfig(cleaning_only=True)
    took = {}
    for workers in ("0", "2"):
        monkeypatch.setenv("MTB200_SAVE_WORKERS", workers)
        t0 = time.perf_counter()
        res = P.batch_translate_images(inp, cfg, tmp_path / f"out{workers}")
        took[workers] = time.perf_counter() - t0
        assert res["success_count"] == 8 and len(list((tmp_path / f"out{workers}").glob("*.png"))) == 8
    assert took["0"] > 1.5                       # 8 x (0.05 + 0.15) s back to back
    assert took["2"] < 0.75 * took["0"], took    # saves ride under the following renders, two at a time


@pytest.mark.parametrize("workers", ["0", "2"], ids=["inline", "decoded_ahead"])
def test_unreadable_page_is_booked_with_the_reference_message(tmp_path, monkeypatch, workers):
    """The real load path (inline or one page ahead on the worker pool): a corrupt file fails with "Error loading image"
    (pipeline.py:689-696) and the following page is still decoded and handed on."""
    monkeypatch.setenv("MTB200_SAVE_WORKERS", workers)
    inp = tmp_path / "in"
    inp.mkdir()
    (inp / "1.png").write_bytes(b"not a png")
    Image.new("RGB", (8, 8), (1, 2, 3)).save(inp / "2.png")
    got = []
    real_load = P._load_page

    def stage(pil, config, bubbles, scale, verbose):           # stands for detect + clean (needs a GPU)
        got.append(pil.size)
        return pil, []

    monkeypatch.setattr(P, "detect_speech_bubbles", lambda *a, **k: ([], []))
    monkeypatch.setattr(P, "_clean_speech_bubbles_for_page", stage)
    monkeypatch.setattr(P, "get_cache", lambda: type("C", (), {"set_current_image": staticmethod(lambda *a, **k: None), "clear_all": staticmethod(lambda: None)})())
    res = P.batch_translate_images(inp, MangaTranslatorConfig(cleaning_only=True), tmp_path / "out")
    assert P._load_page is real_load
    assert res["success_count"] == 1 and res["error_count"] == 1
    assert list(res["errors"]) == ["1.png"] and res["errors"]["1.png"].startswith("Error loading image")
    assert got == [(8, 8)] and (tmp_path / "out" / "2_translated.png").exists()
    assert Image.open(tmp_path / "out" / "2_translated.png").mode == "RGBA"       # PNG target mode (pipeline.py:702-712)


_WORKER = r"""
import os, sys
from pathlib import Path
sys.path.insert(0, {root!r})
from mangatranslator_b200.core import pipeline as P
from mangatranslator_b200.core.config import MangaTranslatorConfig
rank = int(os.environ["RANK"])
def render(path, config, output_path=None, cancellation_manager=None, preloaded=None, device_png=False):
    if Path(path).name == "7.png":
        raise RuntimeError("bad page")
    return "img", "RGB"
P._render_page = render
P._save_page = lambda image, mode, out, config: Path(out).write_text(str(rank))
res = P.batch_translate_images({inp!r}, MangaTranslatorConfig(cleaning_only=True), {out!r})
print("RANK", rank, res["success_count"], res["error_count"], sorted(res["errors"]), flush=True)
"""


def test_batch_flow_shards_pages_over_two_gloo_ranks(tmp_path):
    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir()
    for i in range(1, 12):
        (inp / f"{i}.png").write_bytes(b"x")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, inp=str(inp), out=str(out)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    # natural order 1..11 -> rank 0 takes pages 1,3,5,7,9,11, rank 1 takes 2,4,6,8,10; rank 0 returns the merged result
    assert "RANK 0 10 1 ['7.png']" in r.stdout and "RANK 1 5 0 []" in r.stdout, r.stdout
    owners = {int(p.name.split("_")[0]): p.read_text() for p in out.glob("*_translated.png")}
    assert owners == {i: str((i - 1) % 2) for i in range(1, 12) if i != 7}
    assert (out / "failed_paths.txt").read_text().strip().endswith("7.png")


@needs_ref
@pytest.mark.parametrize("pre_enabled,pre_factor,final,model", [(True, 2.0, False, "model"), (True, 2.0, True, "model_lite"),
                                                                (False, 2.0, True, "model"), (True, 1.005, False, "model"),
                                                                (True, 12.0, False, "model_lite"), (False, 2.0, False, "model")])
def test_upscaling_only_calls_the_upscaler_like_the_live_reference(tmp_path, monkeypatch, pre_enabled, pre_factor, final, model):
    """`upscaling_only` (core/pipeline.py:718-737): the initial upscale uses the OUTPUT model setting and a factor clamped
    to [1, 8] (off at <= 1.01); the final upscale happens only with output.upscale_final_image.  Both drivers run with
    their `upscale_image` replaced by a recorder that resizes with PIL, and must make the same calls and return the same
    size."""
    RP, _ = _ref()
    import core.config as RC
    src = tmp_path / "p.png"
    Image.fromarray(np.full((20, 30, 3), 200, np.uint8)).save(src)

    def recorder(log):
        def up(image, factor, model_type="model", verbose=False):
            log.append((round(float(factor), 4), model_type, image.size))
            return image.resize((int(image.width * factor), int(image.height * factor)))
        return up

    ours_log, ref_log = [], []
    monkeypatch.setattr(P, "upscale_image", recorder(ours_log))
    monkeypatch.setattr(RP, "upscale_image", recorder(ref_log))
    monkeypatch.setattr(RP, "save_image_with_compression", lambda *a, **k: True)

    def cfg_for(mod):
        c = mod.MangaTranslatorConfig(yolo_model_path="", upscaling_only=True)
        c.preprocessing.enabled, c.preprocessing.factor = pre_enabled, pre_factor
        c.output.upscale_final_image, c.output.image_upscale_factor, c.output.image_upscale_model = final, 1.5, model
        return c

    import mangatranslator_b200.core.config as OC
    ours = P.translate_and_render(src, cfg_for(OC), None)
    theirs = RP.translate_and_render(src, cfg_for(RC), None)
    assert ours_log == ref_log, (ours_log, ref_log)
    assert ours.size == theirs.size


def test_fast_path_yields_to_the_stage_function_when_osb_text_verification_can_run(monkeypatch, tmp_path):
    """`use_osb_text_verification` (default on, core/config.py:21) lives in `detect_speech_bubbles` (box expansion,
    text-safe conjoined cuts: detection.py:1555-1571).  The device-resident fast path keeps the page only while that step
    would be skipped anyway (no OSB-text detector available, like the reference after a failed load, :198-201)."""
    import torch
    from PIL import Image
    from mangatranslator_b200.core import pipeline as P
    from mangatranslator_b200.core.config import MangaTranslatorConfig
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    cfg = MangaTranslatorConfig()
    cfg.cleaning_only = True
    cfg.detection.seg_model = "sam2"
    pil = Image.new("RGB", (64, 96), (200, 200, 200))
    sentinel = object()
    key_seen = []

    class _Fake(dict):                          # stands in for the cache of built engines: any key "is there"
        def __contains__(self, k):
            key_seen.append(k)
            return True

        def __getitem__(self, k):
            return sentinel
    mm = get_model_manager()                    # built before CUDA is "made available" (it asks for the best device)
    monkeypatch.setattr(P, "_FAST", _Fake())
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.delenv("MTB200_SYNTHETIC_OSBTEXT", raising=False)
    monkeypatch.delenv("MTB200_FAST_PATH", raising=False)
    mm.models.pop(ModelType.YOLO_OSBTEXT, None)
    monkeypatch.setitem(mm.model_paths, ModelType.YOLO_OSBTEXT, tmp_path / "absent.pt")
    assert cfg.detection.use_osb_text_verification is True
    assert P._fast_path_pipeline(cfg, pil) is sentinel                      # no detector anywhere: the step is a no-op
    monkeypatch.setenv("MTB200_SYNTHETIC_OSBTEXT", "1")
    assert P._fast_path_pipeline(cfg, pil) is None                          # it could load: stage functions
    monkeypatch.delenv("MTB200_SYNTHETIC_OSBTEXT")
    (tmp_path / "present.pt").write_bytes(b"x")
    monkeypatch.setitem(mm.model_paths, ModelType.YOLO_OSBTEXT, tmp_path / "present.pt")
    assert P._fast_path_pipeline(cfg, pil) is None                          # a checkpoint is there
    monkeypatch.setitem(mm.model_paths, ModelType.YOLO_OSBTEXT, tmp_path / "absent.pt")
    mm.models[ModelType.YOLO_OSBTEXT] = object()
    try:
        assert P._fast_path_pipeline(cfg, pil) is None                      # an injected detector
    finally:
        mm.models.pop(ModelType.YOLO_OSBTEXT, None)
    cfg.detection.use_osb_text_verification = False
    assert P._fast_path_pipeline(cfg, pil) is sentinel
