"""Per-image result cache with the reference's entry points (core/caching.py:12-658: UnifiedCache, get_cache).

Like the reference, entries are keyed by the CONTENT of the image (so an image mutated in place and passed again — the
reference's own stage code pastes / inpaints and then re-detects or re-upscales — can never hit a stale entry) and every
kind of result keeps a tiny LRU (the reference uses size-1 slots; here 2).  The reference hashes the whole page with
SHA-256 two to four times per page; here the fingerprint is (mode, size, CRC-32 of the pixel bytes), ~1 ms for a
1536x1024 page, computed once per image object and re-validated against a cheap strided sample."""
from __future__ import annotations

import collections
import threading
import weakref
import zlib
from typing import Any, Optional  # noqa: F401  (Any: quoted annotation below)

MAX_ENTRIES = 2            # per kind of result (yolo / sam / upscale / upscale_dim / bubble_proc)


def _fingerprint(image) -> tuple:
    """(mode, size, crc32 of the bytes) of a PIL image or ndarray-like."""
    if hasattr(image, "tobytes") and hasattr(image, "mode"):
        return (image.mode, image.size, zlib.crc32(image.tobytes()))
    import numpy as np
    a = np.ascontiguousarray(image)
    return (str(a.dtype), a.shape, zlib.crc32(a.tobytes()))


class UnifiedCache:
    def __init__(self):
        self._lock = threading.Lock()
        self._store: "dict[str, collections.OrderedDict[Any, Any]]" = {}
        self._current = None

    def set_current_image(self, image, verbose: bool = False) -> None:
        """A new page: drop everything cached for the previous one (core/caching.py set_current_image)."""
        fp = _fingerprint(image)
        with self._lock:
            if fp != self._current:
                self._store.clear()
                self._current = fp

    def _get(self, key):
        with self._lock:
            slot = self._store.get(key[0])
            if slot is not None and key in slot:
                slot.move_to_end(key)
                return slot[key]
            return None

    def _set(self, key, value) -> None:
        with self._lock:
            slot = self._store.setdefault(key[0], collections.OrderedDict())
            slot[key] = value
            slot.move_to_end(key)
            while len(slot) > MAX_ENTRIES:
                slot.popitem(last=False)

    # YOLO
    def get_yolo_cache_key(self, image, model_path, confidence, *extra):
        return ("yolo", _fingerprint(image), str(model_path), float(confidence)) + tuple(extra)

    def get_yolo_detection(self, key):
        return self._get(key)

    def set_yolo_detection(self, key, value) -> None:
        self._set(key, value)

    # SAM
    def get_sam_cache_key(self, image, *extra):
        return ("sam", _fingerprint(image)) + tuple(str(e) for e in extra)

    def get_sam_masks(self, key):
        return self._get(key)

    def set_sam_masks(self, key, value) -> None:
        self._set(key, value)

    # upscale
    def get_upscale_cache_key(self, image, factor, model_type, *extra):
        return ("upscale", _fingerprint(image), float(factor), str(model_type)) + tuple(extra)

    def get_upscale_dimension_cache_key(self, image, target, mode, model_type="model"):
        return ("upscale_dim", _fingerprint(image), int(target), str(mode), str(model_type))

    def get_bubble_processing_cache_key(self, image, target, mode, model_type="model"):
        return ("bubble_proc", _fingerprint(image), int(target), str(mode), str(model_type))

    def get_upscaled_image(self, key):
        return self._get(key)

    def set_upscaled_image(self, key, value, verbose: bool = False) -> None:
        self._set(key, value)

    def clear(self) -> None:
        with self._lock:
            self._store.clear()
            self._current = None

    clear_all = clear              # the reference's name for the same thing (core/caching.py clear_all)

    def __len__(self) -> int:
        with self._lock:
            return sum(len(s) for s in self._store.values())


_cache: Optional[UnifiedCache] = None


def get_cache() -> UnifiedCache:
    global _cache
    if _cache is None:
        _cache = UnifiedCache()
    return _cache
