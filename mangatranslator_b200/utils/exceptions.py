"""Exception taxonomy of the hot path (same names and base classes as the reference's utils/exceptions.py:1-50,
so callers' `except` clauses keep working)."""


class ValidationError(ValueError):
    """Invalid configuration or argument."""


class ModelError(RuntimeError):
    """Model loading / inference failure (including a missing CUDA library: there is no CPU fallback)."""


class ImageProcessingError(Exception):
    """Image decode / conversion failure."""


class DetectionError(RuntimeError):
    """Speech-bubble detection failure."""


class CleaningError(Exception):
    """Bubble cleaning failure."""


class CancellationError(Exception):
    """Cooperative cancellation of a batch."""


class FontError(RuntimeError):
    pass


class RenderingError(RuntimeError):
    pass


class TranslationError(RuntimeError):
    pass
