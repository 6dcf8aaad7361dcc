"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (/root/reference) with stub modules for the
third-party packages that are absent from this image and not on the hot path.

Only golden-vector generation (oracle/gen_golden.py) and the oracle self-checks in tests/ use this, and only in
the build container: /root/reference does not exist on the GPU box.  Nothing in mangatranslator_b200/ imports it.
"""
import os
import sys
from unittest.mock import MagicMock

REF = os.environ.get("MT_REFERENCE_DIR", "/root/reference")
_STUBS = ["ultralytics", "spandrel", "oxipng", "skia", "uharfbuzz", "fontTools", "fontTools.ttLib", "pythainlp",
          "pythainlp.tokenize", "pythainlp.util", "manga_ocr", "diffusers", "sdnq", "gradio", "nunchaku"]


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "core"))


def import_reference():
    """Returns the reference's `core` package (cleaning/detection/batch_coordinator/scaling importable)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF}")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import core  # noqa: F401  (the reference's package)
    import core.image.cleaning  # noqa: F401
    import core.image.detection  # noqa: F401
    import core.batch_coordinator  # noqa: F401
    import core.scaling  # noqa: F401
    return core
