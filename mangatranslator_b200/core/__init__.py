"""Reference-facing stage interfaces (mirror of the reference's `core` package for the vision hot path only).

The names the reference re-exports from `core` (core/__init__.py:8-41) resolve lazily, so importing the package does not
pull torch / the CUDA library in; names whose subsystem is outside this build (LLM calls, text rendering, OCR, Flux,
reading-order sorting: SURVEY.md §2 "OUT") raise AttributeError with a pointer instead of importing."""
from ._version import __version__, __version_info__

_EXPORTS = {
    "UnifiedCache": "caching", "get_cache": "caching",
    "clean_speech_bubbles": "image.cleaning",
    "detect_speech_bubbles": "image.detection",
    "cv2_to_pil": "image.image_utils", "pil_to_cv2": "image.image_utils",
    "save_image_with_compression": "image.image_utils",
    "ModelManager": "ml.model_manager", "get_model_manager": "ml.model_manager",
    "batch_translate_images": "pipeline", "translate_and_render": "pipeline",
}
_OUT_OF_SCOPE = {"render_text_skia", "call_translation_api_batch", "sort_bubbles_by_reading_order",
                 "OutsideTextDetector", "FluxKontextInpainter", "FluxKleinInpainter"}
__all__ = ["__version__", "__version_info__", *sorted(_EXPORTS)]


def __getattr__(name):
    if name in _EXPORTS:
        import importlib
        value = getattr(importlib.import_module(f"{__name__}.{_EXPORTS[name]}"), name)
        globals()[name] = value
        return value
    if name in _OUT_OF_SCOPE:
        raise AttributeError(f"core.{name} is outside the B200 hot path of this build (keep the reference's module for "
                             "it; SURVEY.md §8f)")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
