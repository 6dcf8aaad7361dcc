"""Resolution-dependent parameter scaling — same functions/semantics as the reference's core/scaling.py:18-109
(host-side scalars only; they select structuring elements and thresholds, so they must match exactly)."""
from typing import Optional, Tuple

from mangatranslator_b200.clean_host import scale_area as _scale_area
from mangatranslator_b200.clean_host import scale_kernel_dim as _scale_kernel_dim
from mangatranslator_b200.clean_host import scale_scalar as _scale_scalar


def scale_scalar(value: float, scale: Optional[float], *, minimum: Optional[float] = None,
                 maximum: Optional[float] = None) -> float:
    return _scale_scalar(value, scale, minimum, maximum)


def scale_length(value: float, scale: Optional[float], *, minimum: Optional[float] = 1.0,
                 maximum: Optional[float] = None) -> int:
    return max(1, int(round(_scale_scalar(value, scale, minimum, maximum))))


def scale_area(value: float, scale: Optional[float], *, minimum: Optional[float] = 1.0,
               maximum: Optional[float] = None) -> int:
    return _scale_area(value, scale, minimum, maximum)


def scale_kernel(kernel: Tuple[int, int], scale: Optional[float], *, minimum: int = 1,
                 maximum: int = 63) -> Tuple[int, int]:
    return (_scale_kernel_dim(kernel[0], scale, minimum, maximum), _scale_kernel_dim(kernel[1], scale, minimum, maximum))


def scale_font_size(value: float, scale: Optional[float], *, minimum: int = 4, maximum: int = 256) -> int:
    return scale_length(value, scale, minimum=minimum, maximum=maximum)
