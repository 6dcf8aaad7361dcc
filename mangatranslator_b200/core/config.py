"""Configuration dataclasses for the hot path — field names and defaults of the reference's core/config.py
(DetectionConfig :11-21, CleaningConfig :25-31, OutputConfig :177-185, PreprocessingConfig :271-276,
MangaTranslatorConfig :189-267) restricted to what detect -> segment -> clean -> upscale reads; the translation /
rendering sections belong to subsystems outside this build.  tests/test_host_logic.py compares every default with the
live reference."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class DetectionConfig:
    confidence: float = 0.6
    conjoined_confidence: float = 0.35
    panel_confidence: float = 0.25
    seg_model: str = "yolo"
    bubble_detector_model: str = "yolo_2"
    conjoined_detection: bool = True
    use_panel_sorting: bool = True
    use_osb_text_verification: bool = True


@dataclass
class CleaningConfig:
    thresholding_value: int = 200
    use_otsu_threshold: bool = False
    roi_shrink_px: int = 5
    inpaint_colored_bubbles: bool = False


@dataclass
class OutsideTextConfig:
    enabled: bool = False


@dataclass
class OutputConfig:
    jpeg_quality: int = 95
    png_compression: int = 2
    output_format: str = "png"
    upscale_final_image: bool = False
    image_upscale_factor: float = 2.0
    image_upscale_model: str = "model_lite"


@dataclass
class PreprocessingConfig:
    enabled: bool = False
    factor: float = 2.0
    auto_scale: bool = True


@dataclass
class MangaTranslatorConfig:
    yolo_model_path: str = ""
    verbose: bool = False
    device: Optional[torch.device] = None
    cleaning_only: bool = False
    upscaling_only: bool = False
    test_mode: bool = False
    processing_scale: float = 1.0
    parallel_requests: int = 1
    batch_parallel_within_pages: bool = False
    overlap_llm_with_inpaint: bool = False
    retry_failed_once: bool = False
    request_coordinator: Optional[object] = None
    detection: DetectionConfig = field(default_factory=DetectionConfig)
    cleaning: CleaningConfig = field(default_factory=CleaningConfig)
    outside_text: OutsideTextConfig = field(default_factory=OutsideTextConfig)
    output: OutputConfig = field(default_factory=OutputConfig)
    preprocessing: PreprocessingConfig = field(default_factory=PreprocessingConfig)
