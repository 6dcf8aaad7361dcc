"""Two ways to put the B200 hot path under the reference's entry points (`app.py`, `main.py`, `ui/*`).

`install()`          aliases this build's modules under the reference's package names (`core`, `utils`): same module
                     objects under both names, so singletons (ModelManager, cache) are shared.  Enough for callers that
                     only touch the vision hot path (cleaning_only / upscaling_only); everything else of the reference's
                     `core` package is then absent.

`install_overlay()`  keeps the UNMODIFIED reference packages importable as `core` / `utils` (LLM calls, OCR, text
                     rendering, UI, validation ... stay the reference's) and replaces only the hot path inside them:
                       * the stage functions (detect_speech_bubbles, clean_speech_bubbles, retry_cleaning_with_otsu,
                         upscale_image, upscale_image_to_dimension, process_bubble_image_cached, resize_to_min/max_side,
                         calculate_centroid_expansion_box) in their defining modules AND in every already imported module
                         that bound them by name (`from .image.detection import detect_speech_bubbles`,
                         core/pipeline.py:33-43, core/text/text_renderer.py:8, core/services/translation.py:13, ...);
                       * the hot-path loaders of the reference's ModelManager (load_yolo_speech_bubble, load_sam2,
                         load_upscale, load_upscale_lite, load_rtdetr_conjoined_bubble: core/ml/model_manager.py:617-1010)
                         delegate to this build's manager, so reference code that loads a model itself and passes it on
                         (core/pipeline.py:235-237,889-891, core/outside_text_processor.py:94-107) gets a B200 object, which is
                         also entered in the reference manager's `models` table; `unload_model` / `unload_all` (and
                         through them unload_ocr_models / unload_upscale_models) are forwarded;
                       * this build's exception classes are rebound to the reference's (utils/exceptions.py), so the
                         reference's `except ImageProcessingError` / `except ModelError` clauses catch what the B200 stage
                         functions raise.
                     `uninstall_overlay()` restores every binding.
"""
from __future__ import annotations

import importlib
import sys
from typing import Any, Dict, List, Optional, Tuple

_MODULES = [
    ("utils", "mangatranslator_b200.utils"),
    ("utils.logging", "mangatranslator_b200.utils.logging"),
    ("utils.exceptions", "mangatranslator_b200.utils.exceptions"),
    ("core", "mangatranslator_b200.core"),
    ("core._version", "mangatranslator_b200.core._version"),
    ("core.scaling", "mangatranslator_b200.core.scaling"),
    ("core.device", "mangatranslator_b200.core.device"),
    ("core.caching", "mangatranslator_b200.core.caching"),
    ("core.config", "mangatranslator_b200.core.config"),
    ("core.batch_coordinator", "mangatranslator_b200.core.batch_coordinator"),
    ("core.ml", "mangatranslator_b200.core.ml"),
    ("core.ml.model_manager", "mangatranslator_b200.core.ml.model_manager"),
    ("core.image", "mangatranslator_b200.core.image"),
    ("core.image.detection", "mangatranslator_b200.core.image.detection"),
    ("core.image.cleaning", "mangatranslator_b200.core.image.cleaning"),
    ("core.image.image_utils", "mangatranslator_b200.core.image.image_utils"),
    ("core.pipeline", "mangatranslator_b200.core.pipeline"),
]


def install(force: bool = False) -> None:
    for alias, real in _MODULES:
        if alias in sys.modules and not force:
            continue
        sys.modules[alias] = importlib.import_module(real)


# ---- overlay on the unmodified reference ---------------------------------------------------------------------------------
# reference module -> (our module, names replaced there)
_STAGE_FUNCTIONS = {
    "core.image.detection": ("mangatranslator_b200.core.image.detection", ["detect_speech_bubbles", "detect_panels"]),
    "core.image.cleaning": ("mangatranslator_b200.core.image.cleaning", ["clean_speech_bubbles", "retry_cleaning_with_otsu"]),
    "core.image.image_utils": ("mangatranslator_b200.core.image.image_utils",
                               ["upscale_image", "upscale_image_to_dimension", "process_bubble_image_cached",
                                "resize_to_max_side", "resize_to_min_side", "calculate_centroid_expansion_box"]),
}
_LOADERS = ["load_yolo_speech_bubble", "load_sam2", "load_upscale", "load_upscale_lite", "load_rtdetr_conjoined_bubble",
            "load_yolo_panel", "load_yolo_osbtext"]
_undo: List[Tuple[Any, str, Any]] = []          # (namespace dict or class, name, original value)


def _set(ns, name: str, value) -> None:
    if isinstance(ns, dict):
        _undo.append((ns, name, ns[name]))
        ns[name] = value
    else:
        _undo.append((ns, name, ns.__dict__[name]))
        setattr(ns, name, value)


def _rebind(prefixes: Tuple[str, ...], mapping: Dict[int, Any], skip: Tuple[str, ...] = ()) -> int:
    """In every loaded module whose name starts with one of `prefixes`, rebind globals that ARE a key object."""
    n = 0
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name in prefixes or mod_name.startswith(tuple(p + "." for p in prefixes))):
            continue
        if mod_name in skip:
            continue
        ns = getattr(mod, "__dict__", None)
        if not isinstance(ns, dict):
            continue
        for key, val in list(ns.items()):
            new = mapping.get(id(val))
            if new is not None and new is not val:
                _set(ns, key, new)
                n += 1
    return n


def install_overlay(reference_root: Optional[str] = None, extra_prefixes: Tuple[str, ...] = ("main", "app", "ui")) -> dict:
    """See the module docstring.  `reference_root`: directory holding the reference's `core/` and `utils/` (put on
    sys.path when given).  Returns counts of what was rebound.  Idempotent; undo with `uninstall_overlay()`."""
    if _undo:
        return {"already_installed": True}
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    ref_core = importlib.import_module("core")
    if getattr(ref_core, "__name__", "") != "core" or "mangatranslator_b200" in (getattr(ref_core, "__file__", "") or ""):
        raise RuntimeError("install_overlay needs the reference's own `core` package (install() aliased ours under that name)")
    stats = {"functions": 0, "importers": 0, "exceptions": 0, "loaders": 0}
    # 1. exceptions first: the B200 modules must raise the reference's classes
    ref_exc = importlib.import_module("utils.exceptions")
    our_exc = importlib.import_module("mangatranslator_b200.utils.exceptions")
    our_mods = [importlib.import_module(real) for _, real in _MODULES] + \
               [importlib.import_module("mangatranslator_b200.safebox_host")]
    exc_map = {}
    for name, cls in list(vars(our_exc).items()):
        ref_cls = getattr(ref_exc, name, None)
        if isinstance(cls, type) and issubclass(cls, BaseException) and isinstance(ref_cls, type):
            exc_map[id(cls)] = ref_cls
    stats["exceptions"] = _rebind(("mangatranslator_b200",), exc_map)
    # 2. stage functions, in the defining modules and in every importer that bound them by name
    fn_map = {}
    for ref_name, (our_name, names) in _STAGE_FUNCTIONS.items():
        ref_mod, our_mod = importlib.import_module(ref_name), importlib.import_module(our_name)
        for n in names:
            old = getattr(ref_mod, n, None)
            if old is None:
                continue
            fn_map[id(old)] = getattr(our_mod, n)
            _set(ref_mod.__dict__, n, getattr(our_mod, n))
            stats["functions"] += 1
    stats["importers"] = _rebind(("core",) + tuple(extra_prefixes), fn_map)
    # 3. hot-path loaders of the reference's ModelManager delegate to this build's manager
    ref_mm = importlib.import_module("core.ml.model_manager")
    our_mm = importlib.import_module("mangatranslator_b200.core.ml.model_manager")

    def mirror(ref_self, obj):
        """Make the delegated object visible in the reference manager's own table (is_loaded, memory stats)."""
        for t, v in list(our_mm.get_model_manager().models.items()):
            if v is obj and hasattr(ref_mm.ModelType, t.name):
                ref_self.models[ref_mm.ModelType[t.name]] = obj

    def delegate(method: str):
        def loader(self, *args, **kwargs):
            obj = getattr(our_mm.get_model_manager(), method)(*args, **kwargs)
            mirror(self, obj)
            return obj
        loader.__name__ = method
        loader.__doc__ = f"B200 overlay: delegates to mangatranslator_b200 ModelManager.{method}"
        return loader

    for m in _LOADERS:
        if m in ref_mm.ModelManager.__dict__:
            _set(ref_mm.ModelManager, m, delegate(m))
            stats["loaders"] += 1

    # unloads are forwarded so the B200 objects really leave the device when the reference drops its handle
    def forward(method: str, call):
        orig = ref_mm.ModelManager.__dict__.get(method)
        if orig is None:
            return

        def wrapper(self, *args, **kwargs):
            call(our_mm.get_model_manager(), *args, **kwargs)
            return orig(self, *args, **kwargs)
        wrapper.__name__ = method
        _set(ref_mm.ModelManager, method, wrapper)

    def unload_one(ours, model_type, force_gc=True, verbose=False):
        name = getattr(model_type, "name", None)
        if name in our_mm.ModelType.__members__:
            ours.unload_model(our_mm.ModelType[name], force_gc=force_gc, verbose=verbose)

    forward("unload_model", unload_one)
    forward("unload_all", lambda ours, verbose=False: ours.unload_all(verbose=verbose))
    del our_mods
    return stats


def uninstall_overlay() -> None:
    while _undo:
        ns, name, value = _undo.pop()
        if isinstance(ns, dict):
            ns[name] = value
        else:
            setattr(ns, name, value)
