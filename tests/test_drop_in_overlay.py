"""CPU (build container only — needs the reference checkout): `drop_in.install_overlay()` puts the B200 hot path under
the UNMODIFIED reference packages.  Each case runs in a subprocess so the rebinding never leaks into tests that compare
against the unmodified reference functions."""
import os
import subprocess
import sys
import textwrap

import pytest

import _refimport
from helpers import ROOT

pytestmark = pytest.mark.skipif(not _refimport.available(), reason="reference checkout not present (GPU box)")

PRELUDE = f"""
import sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, "oracle")!r})
import _refimport
_refimport.import_reference()      # reference on sys.path + stubs for the third-party packages absent from this image
"""


def _run(body: str) -> str:
    code = PRELUDE + textwrap.dedent(body)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return p.stdout


def test_overlay_rebinds_stage_functions_exceptions_and_loaders_and_restores_them():
    out = _run("""
        import core.pipeline as RP                  # imported BEFORE the overlay: its bound names must be rebound too
        import core.text.text_renderer as RT
        import core.services.translation as RS
        import core.image.detection as RD
        import utils.exceptions as RE
        before = (RP.detect_speech_bubbles, RD.detect_speech_bubbles, RT.calculate_centroid_expansion_box)
        import mangatranslator_b200.drop_in as D
        stats = D.install_overlay()
        assert stats["functions"] == 10 and stats["loaders"] == 7 and stats["importers"] >= 10 and stats["exceptions"] >= 5, stats
        assert D.install_overlay() == {"already_installed": True}
        import mangatranslator_b200.core.image.detection as OD
        import mangatranslator_b200.core.image.cleaning as OC
        import mangatranslator_b200.core.image.image_utils as OU
        assert RP.detect_speech_bubbles is OD.detect_speech_bubbles is RD.detect_speech_bubbles
        assert RP.clean_speech_bubbles is OC.clean_speech_bubbles and RP.retry_cleaning_with_otsu is OC.retry_cleaning_with_otsu
        assert RP.upscale_image is OU.upscale_image and RS.process_bubble_image_cached is OU.process_bubble_image_cached
        assert RT.calculate_centroid_expansion_box is OU.calculate_centroid_expansion_box
        import core
        assert core.detect_speech_bubbles is OD.detect_speech_bubbles          # the package-level re-export as well
        assert RD.detect_panels is OD.detect_panels                            # the panel detector is this build's too
        # the B200 modules now raise the reference's exception classes
        assert OU.ImageProcessingError is RE.ImageProcessingError and OD.ModelError is RE.ModelError
        # hot-path loaders of the reference's manager delegate (no GPU here: the B200 manager refuses, with THEIR class)
        import core.ml.model_manager as RM
        for m in ("load_upscale", "load_upscale_lite", "load_sam2"):
            try:
                getattr(RM.get_model_manager(), m)()
                raise SystemExit(m + " did not raise")
            except RE.ModelError as e:
                assert "no CPU fallback" in str(e)
        # a delegated load is entered in the reference manager's own table; its unloads reach the B200 manager
        import mangatranslator_b200.core.ml.model_manager as OM
        ours, theirs, dummy = OM.get_model_manager(), RM.get_model_manager(), object()

        def fake_load(verbose=False):
            ours.models[OM.ModelType.UPSCALE] = dummy
            return dummy
        ours.load_upscale = fake_load
        assert theirs.load_upscale() is dummy and theirs.is_loaded(RM.ModelType.UPSCALE)
        assert theirs.models[RM.ModelType.UPSCALE] is dummy and ours.is_loaded(OM.ModelType.UPSCALE)
        theirs.unload_upscale_models()
        assert not theirs.is_loaded(RM.ModelType.UPSCALE) and not ours.is_loaded(OM.ModelType.UPSCALE)
        ours.models[OM.ModelType.SAM2] = dummy
        theirs.unload_all()
        assert not ours.is_loaded(OM.ModelType.SAM2) and not ours.models
        del ours.load_upscale
        D.uninstall_overlay()
        assert (RP.detect_speech_bubbles, RD.detect_speech_bubbles, RT.calculate_centroid_expansion_box) == before
        assert OU.ImageProcessingError is not RE.ImageProcessingError
        assert RM.ModelManager.load_upscale.__doc__ is None or "B200 overlay" not in RM.ModelManager.load_upscale.__doc__
        print("OK")
    """)
    assert out.strip().endswith("OK")


def test_reference_pipeline_runs_through_the_overlay_with_our_signatures():
    """The reference's own translate_and_render (cleaning_only + final upscale) under the overlay: every call it makes
    into the hot path must bind to the REAL B200 function signatures (checked with inspect.Signature.bind); the stage
    bodies are replaced by canned results because they need a GPU."""
    out = _run("""
        import inspect, os, tempfile
        import numpy as np
        from PIL import Image
        import mangatranslator_b200.core.image.detection as OD
        import mangatranslator_b200.core.image.cleaning as OC
        import mangatranslator_b200.core.image.image_utils as OU
        from mangatranslator_b200 import synth
        calls = {}

        def checked(mod, name, result):
            sig = inspect.signature(getattr(mod, name))
            def w(*a, **k):
                sig.bind(*a, **k)
                calls[name] = sorted(k)
                return result(*a, **k)
            w.__name__ = name
            setattr(mod, name, w)

        pg = synth.make_page(3, 384, 256, n_bubbles=2)
        dets = synth.detections_from_page(pg)
        checked(OD, "detect_speech_bubbles", lambda *a, **k: (dets, []))
        checked(OC, "clean_speech_bubbles",
                lambda image, *a, **k: (np.ascontiguousarray(np.asarray(image.convert("RGB"))[:, :, ::-1]), []))
        checked(OU, "upscale_image", lambda image, factor, **k: image.resize((int(image.width * factor), int(image.height * factor))))
        import mangatranslator_b200.drop_in as D
        D.install_overlay()
        import core.pipeline as RP
        from core.config import MangaTranslatorConfig
        d = tempfile.mkdtemp()
        src = os.path.join(d, "page.png")
        Image.fromarray(pg.image_rgb).save(src)
        cfg = MangaTranslatorConfig(yolo_model_path="x.pt", cleaning_only=True)
        cfg.output.upscale_final_image, cfg.output.image_upscale_factor, cfg.output.output_format = True, 2.0, "jpeg"
        out = RP.translate_and_render(src, cfg, output_path=os.path.join(d, "out.jpg"))
        assert out.size == (512, 768) and os.path.exists(os.path.join(d, "out.jpg"))
        assert set(calls) == {"detect_speech_bubbles", "clean_speech_bubbles", "upscale_image"}, calls
        assert {"seg_model", "conjoined_detection", "image_override", "bubble_detector_model"} <= set(calls["detect_speech_bubbles"])
        assert {"pre_computed_detections", "processing_scale", "request_coordinator", "inpaint_colored_bubbles"} <= set(calls["clean_speech_bubbles"])
        print("OK")
    """)
    assert out.strip().endswith("OK")


def test_overlaid_functions_have_the_reference_signatures():
    """Every function / loader the overlay replaces takes the reference's parameters: same names, same order, same
    defaults (reading signatures only; nothing is rebound here)."""
    import importlib
    import inspect
    _refimport.import_reference()
    from mangatranslator_b200 import drop_in as D
    checked = 0
    for ref_name, (our_name, names) in D._STAGE_FUNCTIONS.items():
        ref_mod, our_mod = importlib.import_module(ref_name), importlib.import_module(our_name)
        for n in names:
            ref_p = list(inspect.signature(getattr(ref_mod, n)).parameters.values())
            our_p = list(inspect.signature(getattr(our_mod, n)).parameters.values())
            assert [p.name for p in our_p[:len(ref_p)]] == [p.name for p in ref_p], n
            for r, o in zip(ref_p, our_p):
                if r.default is not inspect.Parameter.empty:
                    assert o.default == r.default, (n, r.name, r.default, o.default)
            assert all(p.default is not inspect.Parameter.empty for p in our_p[len(ref_p):]), n
            checked += 1
    ref_mm = importlib.import_module("core.ml.model_manager").ModelManager
    our_mm = importlib.import_module("mangatranslator_b200.core.ml.model_manager").ModelManager
    for m in D._LOADERS:
        ref_p = list(inspect.signature(getattr(ref_mm, m)).parameters.values())
        our_p = list(inspect.signature(getattr(our_mm, m)).parameters.values())
        assert [(p.name, p.default) for p in ref_p] == [(p.name, p.default) for p in our_p], m
        checked += 1
    assert checked == 17
