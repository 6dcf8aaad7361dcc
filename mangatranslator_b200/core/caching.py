"""Per-image result cache with the reference's entry points (core/caching.py:12-658: UnifiedCache, get_cache).
The reference hashes the whole page with SHA-256 two to four times per page to key size-1 LRU slots; on the batch
path that is pure overhead, so this build keeps the API (set_current_image / get_* / set_*) but keys by object identity
of the image.  A key holds a strong reference to its image, so an identity can never be recycled for another image
while its entry is alive; the store is a small LRU."""
from __future__ import annotations

import collections
import threading
from typing import Any, Optional  # noqa: F401  (Any: quoted annotation below)

MAX_ENTRIES = 64


class _ImageKey:
    """(image identity, parameters): equal only for the very same image object."""
    __slots__ = ("image", "rest", "_hash")

    def __init__(self, image, *rest):
        self.image, self.rest = image, rest
        self._hash = hash((id(image),) + tuple(rest))

    def __hash__(self):
        return self._hash

    def __eq__(self, other):
        return isinstance(other, _ImageKey) and other.image is self.image and other.rest == self.rest


class UnifiedCache:
    def __init__(self):
        self._lock = threading.Lock()
        self._store: "collections.OrderedDict[Any, Any]" = collections.OrderedDict()
        self._current = None

    def set_current_image(self, image, verbose: bool = False) -> None:
        with self._lock:
            if image is not self._current:
                self._store.clear()
                self._current = image

    def _get(self, key):
        with self._lock:
            if key in self._store:
                self._store.move_to_end(key)
                return self._store[key]
            return None

    def _set(self, key, value) -> None:
        with self._lock:
            self._store[key] = value
            self._store.move_to_end(key)
            while len(self._store) > MAX_ENTRIES:
                self._store.popitem(last=False)

    # YOLO
    def get_yolo_cache_key(self, image, model_path, confidence, *extra):
        return _ImageKey(image, "yolo", str(model_path), float(confidence), *extra)

    def get_yolo_detection(self, key):
        return self._get(key)

    def set_yolo_detection(self, key, value) -> None:
        self._set(key, value)

    # SAM
    def get_sam_cache_key(self, image, *extra):
        return _ImageKey(image, "sam", *[str(e) for e in extra])

    def get_sam_masks(self, key):
        return self._get(key)

    def set_sam_masks(self, key, value) -> None:
        self._set(key, value)

    # upscale
    def get_upscale_cache_key(self, image, factor, model_type, *extra):
        return _ImageKey(image, "upscale", float(factor), str(model_type), *extra)

    def get_upscale_dimension_cache_key(self, image, target, mode, model_type="model"):
        return _ImageKey(image, "upscale_dim", int(target), str(mode), str(model_type))

    def get_bubble_processing_cache_key(self, image, target, mode, model_type="model"):
        return _ImageKey(image, "bubble_proc", int(target), str(mode), str(model_type))

    def get_upscaled_image(self, key):
        return self._get(key)

    def set_upscaled_image(self, key, value, verbose: bool = False) -> None:
        self._set(key, value)

    def clear(self) -> None:
        with self._lock:
            self._store.clear()
            self._current = None


_cache: Optional[UnifiedCache] = None


def get_cache() -> UnifiedCache:
    global _cache
    if _cache is None:
        _cache = UnifiedCache()
    return _cache
