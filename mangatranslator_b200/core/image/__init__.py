"""Stage functions of the vision hot path under the names the reference exports from `core.image`
(core/image/__init__.py:12-42).  OCR and Flux helpers of that package are outside this build."""
from .cleaning import clean_speech_bubbles, retry_cleaning_with_otsu
from .detection import detect_panels, detect_speech_bubbles
from .image_utils import (calculate_centroid_expansion_box, convert_image_to_target_mode, cv2_to_pil, pil_to_cv2,
                          process_bubble_image_cached, resize_to_max_side, resize_to_min_side,
                          save_image_with_compression, upscale_image, upscale_image_to_dimension)

__all__ = ["clean_speech_bubbles", "retry_cleaning_with_otsu", "detect_speech_bubbles", "detect_panels",
           "calculate_centroid_expansion_box", "convert_image_to_target_mode",
           "cv2_to_pil", "pil_to_cv2", "process_bubble_image_cached", "resize_to_max_side", "resize_to_min_side",
           "save_image_with_compression", "upscale_image", "upscale_image_to_dimension"]

_OUT_OF_SCOPE = {"FluxKontextInpainter", "OutsideTextDetector"}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise AttributeError(f"core.image.{name} is outside the B200 hot path of this build (keep the reference's module "
                             "for it; SURVEY.md §8f)")
    raise AttributeError(name)
