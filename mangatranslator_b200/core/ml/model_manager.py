"""Model loader — drop-in for the reference's core/ml/model_manager.py (ModelManager singleton :57-69, load_* :617-1046,
unload_* :1369-1493, get_model_manager :1520) for the hot-path models.

`load_yolo_speech_bubble`, `load_sam2`, `load_upscale`, `load_upscale_lite` return B200-native objects that satisfy
exactly what the stage functions touch on the third-party objects the reference returns:
  YOLO     -> mangatranslator_b200.yolo.YoloB200     (callable -> [Results], .names)
  SAM 2.1  -> (Sam2ProcessorB200, Sam2ModelB200)     (processor(...), model(...).pred_masks, post_process_masks)
  upscaler -> mangatranslator_b200.rcan.RcanB200     (callable float32 (1,3,h,w) -> (1,3,2h,2w))
Objects injected into ``ModelManager.models[ModelType.X]`` by a caller are returned as they are (the reference's tests /
integrations do that).  Models outside the hot path (Flux, SAM3, OCR) raise ModelError.
"""
from __future__ import annotations

import gc
import os
import threading
from enum import Enum
from pathlib import Path
from typing import Any, Dict, Optional

import torch

from mangatranslator_b200 import weights as W
from mangatranslator_b200.core.device import get_best_device, get_best_dtype
from mangatranslator_b200.utils.exceptions import ModelError
from mangatranslator_b200.utils.logging import log_message


class ModelType(Enum):
    """Same member names as the reference (core/ml/model_manager.py:31-54)."""
    UPSCALE = "upscale"
    UPSCALE_LITE = "upscale_lite"
    YOLO_SPEECH_BUBBLE = "yolo_speech_bubble"
    YOLO_SPEECH_BUBBLE_2 = "yolo_speech_bubble_2"
    RTDETR_CONJOINED_BUBBLE = "rtdetr_conjoined_bubble"
    YOLO_OSBTEXT = "yolo_osbtext"
    YOLO_PANEL = "yolo_panel"
    SAM2 = "sam2"
    SAM3 = "sam3"
    MANGA_OCR = "manga_ocr"
    PADDLE_OCR_VL = "paddle_ocr_vl"
    FLUX_TRANSFORMER = "flux_transformer"
    FLUX_TEXT_ENCODER = "flux_text_encoder"
    FLUX_PIPELINE = "flux_pipeline"
    FLUX_KONTEXT_SDNQ_PIPELINE = "flux_kontext_sdnq_pipeline"
    FLUX_KLEIN_9B_PIPELINE = "flux_klein_9b_pipeline"
    FLUX_KLEIN_4B_PIPELINE = "flux_klein_4b_pipeline"
    SDCPP_SERVER = "sdcpp_server"
    FLUX_KLEIN_SDCPP_VAE = "flux_klein_sdcpp_vae"
    FLUX_KONTEXT_SDCPP_CLIP_L = "flux_kontext_sdcpp_clip_l"
    FLUX_KONTEXT_SDCPP_VAE = "flux_kontext_sdcpp_vae"


class ModelManager:
    _instance = None
    _lock = threading.RLock()

    def __new__(cls, *args, **kwargs):
        with cls._lock:
            if cls._instance is None:
                cls._instance = super().__new__(cls)
                cls._instance._initialized = False
            return cls._instance

    def __init__(self, models_dir: str = "./models"):
        with self._lock:
            if self._initialized:
                return
            self.device = get_best_device()
            self.dtype = get_best_dtype(self.device)
            self.models: Dict[ModelType, Any] = {}
            self.models_dir = Path(os.environ.get("MT_MODELS_DIR") or models_dir)
            self.model_paths: Dict[ModelType, Path] = {
                ModelType.UPSCALE: self.models_dir / "upscale" / "2x-AnimeSharpV4_RCAN.safetensors",
                ModelType.UPSCALE_LITE: self.models_dir / "upscale" / "2x-AnimeSharpV4_Fast_RCAN_PU.safetensors",
                ModelType.YOLO_SPEECH_BUBBLE: self.models_dir / "yolo" / "yolov8m_seg-speech-bubble.pt",
                ModelType.YOLO_SPEECH_BUBBLE_2: self.models_dir / "yolo" / "manga109-segmentation-bubble.pt",
                ModelType.RTDETR_CONJOINED_BUBBLE: self.models_dir / "rtdetr" / "comic-text-and-bubble-detector",
                ModelType.YOLO_OSBTEXT: self.models_dir / "yolo" / "animetext_yolov12x.pt",
                ModelType.YOLO_PANEL: self.models_dir / "yolo" / "manga109_v2023.12.07_l_yolov11.pt",
            }
            self.model_hf_repos: Dict[ModelType, str] = {
                ModelType.SAM2: "facebook/sam2.1-hiera-large",
                ModelType.YOLO_SPEECH_BUBBLE: "kitsumed/yolov8m_seg-speech-bubble",
                ModelType.YOLO_SPEECH_BUBBLE_2: "huyvux3005/manga109-segmentation-bubble",
                ModelType.YOLO_OSBTEXT: "deepghs/AnimeText_yolo",
                ModelType.YOLO_PANEL: "deepghs/manga109_yolo",
            }
            self.hf_token: str = ""
            self.flux_inference_lock = threading.Lock()
            self.precision = os.environ.get("MTB200_PRECISION", "bf16x3")
            self.synthetic_seed = int(os.environ.get("MTB200_WEIGHT_SEED", "0"))
            self._initialized = True

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _require_cuda(self) -> torch.device:
        if not torch.cuda.is_available():
            raise ModelError("CUDA device required: the B200 build has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def set_hf_token(self, token: str) -> None:
        self.hf_token = token or ""

    def is_loaded(self, model_type: ModelType) -> bool:
        return self.models.get(model_type) is not None        # the reference leaves `models[t] = None` after an unload

    def _load_file_state_dict(self, path: Path) -> Optional[dict]:
        if not path.exists():
            return None
        if path.suffix == ".safetensors":
            from safetensors.torch import load_file
            return load_file(str(path))
        obj = torch.load(str(path), map_location="cpu", weights_only=True)
        if not isinstance(obj, dict):
            raise ModelError(f"{path}: not a state dict")
        return obj

    # ---- loaders -----------------------------------------------------------------------------------------------
    def _resolve_yolo_type(self, model_path) -> ModelType:
        """core/ml/model_manager.py:702-709: only the resolved default v2 path maps to the V2 slot."""
        try:
            if model_path and Path(model_path).resolve() == self.model_paths[ModelType.YOLO_SPEECH_BUBBLE_2].resolve():
                return ModelType.YOLO_SPEECH_BUBBLE_2
        except Exception:
            pass
        return ModelType.YOLO_SPEECH_BUBBLE

    def _synthetic_or_raise(self, what: str, looked: list):
        """Weight policy: a missing checkpoint is an error (the reference would download it; there is no network here)
        unless seeded synthetic weights were asked for explicitly with MTB200_SYNTHETIC_WEIGHTS=1 (tests, bench, smoke)."""
        where = ", ".join(str(p) for p in looked) or "no path given"
        if not W.synthetic_allowed():
            raise ModelError(f"{what}: checkpoint not found ({where}). Put the reference's model file there (or point "
                             "MT_MODELS_DIR at the models directory); set MTB200_SYNTHETIC_WEIGHTS=1 only to run with seeded "
                             "synthetic weights (benchmarks / tests: outputs are meaningless on real pages).")
        log_message(f"{what}: NO CHECKPOINT ({where}) - using seeded SYNTHETIC weights (MTB200_SYNTHETIC_WEIGHTS=1, seed "
                    f"{self.synthetic_seed}); outputs are meaningless on real pages", always_print=True)

    def load_yolo_speech_bubble(self, model_path=None, verbose: bool = False):
        mt = self._resolve_yolo_type(model_path)
        with self._lock:
            if self.models.get(mt) is not None:
                return self.models[mt]
            from mangatranslator_b200.yolo import YoloB200
            dev = self._require_cuda()
            # the path the caller configured, then the reference's default locations under the models directory
            cands = ([Path(model_path)] if model_path else []) + [self.model_paths[mt], self.model_paths[ModelType.YOLO_SPEECH_BUBBLE_2],
                                                                 self.model_paths[ModelType.YOLO_SPEECH_BUBBLE]]
            path = next((p for p in cands if p.is_file()), None)
            names = None
            if path is not None:
                try:
                    raw, names = W.load_ultralytics_state_dict(str(path))
                    sd, cfg = W.yolo_from_ultralytics(raw, names)
                except W.UnsupportedCheckpoint as e:
                    # not the hard-wired YOLOv8-seg graph (the default `yolo_2` file may be a YOLO11-seg): run the model
                    # from the module tree the file pickles (mangatranslator_b200/yolo_tree.py)
                    try:
                        from mangatranslator_b200.yolo_tree import YoloTreeB200
                        tree = W.load_ultralytics_tree(str(path))
                        self.models[mt] = YoloTreeB200(tree, dev, precision=self.precision)
                    except W.UnsupportedCheckpoint as e2:
                        raise ModelError(f"YOLO speech-bubble detector: cannot use {path}: {e}; as a module tree: {e2}") from e2
                    except Exception as e2:
                        raise ModelError(f"YOLO speech-bubble detector: failed to read {path}: {e2}") from e2
                    log_message(f"YOLO speech-bubble detector: module tree and weights from {path} "
                                f"({len(tree['layers'])} layers, {tree['layers'][-1]['t']} head, nc={self.models[mt].nc})",
                                always_print=True)
                    return self.models[mt]
                except Exception as e:
                    raise ModelError(f"YOLO speech-bubble detector: failed to read {path}: {e}") from e
                log_message(f"YOLO speech-bubble detector: weights from {path} (YOLOv8-seg, {len(sd)} tensors, nc={cfg['nc']})",
                            always_print=True)
            else:
                self._synthetic_or_raise("YOLO speech-bubble detector", cands[:2])
                cfg = W.yolo_cfg("m")
                sd = W.yolo_state_dict(self.synthetic_seed, cfg)
            if isinstance(names, (list, tuple)):
                names = dict(enumerate(names))
            self.models[mt] = YoloB200(sd, cfg, dev, precision=self.precision, names=names if isinstance(names, dict) else None)
            return self.models[mt]

    def _sam_checkpoint_dir(self) -> Optional[str]:
        """A `from_pretrained`-loadable directory: $MT_MODELS_DIR/sam (flat), or the snapshot inside the Hugging Face cache
        the reference fills (`cache_dir="models/sam"`, core/ml/model_manager.py:994-1003)."""
        roots = [self.models_dir / "sam", Path("models") / "sam"]
        for root in roots:
            if (root / "config.json").is_file():
                return str(root)
            for snap in sorted(root.glob("models--facebook--sam2*/snapshots/*")):
                if (snap / "config.json").is_file():
                    return str(snap)
        return None

    def load_sam2(self, verbose: bool = False):
        with self._lock:
            if self.models.get(ModelType.SAM2) is not None:
                return self.models[ModelType.SAM2]
            from mangatranslator_b200.sam2 import Sam2B200
            from mangatranslator_b200.sam2_api import Sam2ModelB200, Sam2ProcessorB200
            dev = self._require_cuda()
            ckpt = self._sam_checkpoint_dir()
            if ckpt is not None:
                try:
                    from transformers import Sam2Model
                    m = Sam2Model.from_pretrained(ckpt, local_files_only=True)
                    cfg, sd = m.config, m.state_dict()
                except Exception as e:
                    raise ModelError(f"SAM 2.1: failed to read {ckpt}: {e}") from e
                log_message(f"SAM 2.1: weights from {ckpt}", always_print=True)
            else:
                self._synthetic_or_raise("SAM 2.1", [self.models_dir / "sam"])
                cfg, sd = W.sam2_model_and_state(self.synthetic_seed, os.environ.get("MTB200_SAM_VARIANT", "tiny"))
            net = Sam2B200(sd, cfg, dev, precision=self.precision)
            self.models[ModelType.SAM2] = (Sam2ProcessorB200(net), Sam2ModelB200(net))
            return self.models[ModelType.SAM2]

    def _load_rcan(self, mt: ModelType, verbose: bool):
        with self._lock:
            if self.models.get(mt) is not None:
                return self.models[mt]
            from mangatranslator_b200.rcan import RcanB200
            dev = self._require_cuda()
            try:
                sd = self._load_file_state_dict(self.model_paths[mt])
            except Exception as e:
                raise ModelError(f"Upscaler: failed to read {self.model_paths[mt]}: {e}") from e
            if sd is not None:
                log_message(f"Upscaler ({mt.value}): weights from {self.model_paths[mt]}", always_print=True)
            else:
                self._synthetic_or_raise(f"Upscaler ({mt.value})", [self.model_paths[mt]])
                lite = mt == ModelType.UPSCALE_LITE        # "_PU": pixel-unshuffle x2 in front of a shallower body
                sd = W.rcan_state_dict(self.synthetic_seed, n_resblocks=6 if lite else 20, n_resgroups=4 if lite else 10,
                                       unshuffle=2 if lite else 1)
            # fp32-grade mode: the RCAN body runs in its own format (fp16 + e5m2 correction, MTB200_RCAN_PRECISION=bf16x3
            # selects the bf16 hi/lo A/B partner); MTB200_PRECISION=bf16 is the plain-bf16 (non-parity) network
            self.models[mt] = RcanB200(sd, dev, precision=None if self.precision == "bf16x3" else self.precision)
            return self.models[mt]

    def load_upscale(self, verbose: bool = False):
        return self._load_rcan(ModelType.UPSCALE, verbose)

    def load_upscale_lite(self, verbose: bool = False):
        return self._load_rcan(ModelType.UPSCALE_LITE, verbose)

    def _out_of_scope(self, what: str):
        raise ModelError(f"{what} is outside the B200 hot path of this build (SURVEY.md §8f)")

    def load_rtdetr_conjoined_bubble(self, verbose: bool = False):
        """RT-DETRv2 conjoined / fallback bubble detector (reference :745-778 returns an RTDetrYOLOAdapter; the B200
        object has the same call shape and `.names`)."""
        with self._lock:
            if self.models.get(ModelType.RTDETR_CONJOINED_BUBBLE) is not None:
                return self.models[ModelType.RTDETR_CONJOINED_BUBBLE]
            from mangatranslator_b200.rtdetr import RtDetrB200
            log_message("Loading RT-DETR conjoined bubble detection model...", verbose=verbose)
            # Unlike the primary detector, the secondary one is optional in the reference's flow (a load failure is
            # swallowed, core/image/detection.py:1541-1548), and random boxes from seeded weights would be merged into the
            # primaries as "missed bubbles": without a checkpoint it only loads when explicitly asked for.
            ckpt = self.model_paths[ModelType.RTDETR_CONJOINED_BUBBLE]
            if not ckpt.is_dir() and os.environ.get("MTB200_SYNTHETIC_RTDETR", "0") != "1":
                raise ModelError(f"Failed to load RT-DETR conjoined model: no checkpoint at {ckpt} "
                                 "(set MTB200_SYNTHETIC_RTDETR=1 for seeded synthetic weights)")
            dev = self._require_cuda()
            try:
                cfg, sd = W.rtdetr_model_and_state(self.synthetic_seed, checkpoint_dir=str(ckpt) if ckpt.is_dir() else None)
                model = RtDetrB200(sd, cfg, dev, precision=self.precision)
            except Exception as e:
                raise ModelError(f"Failed to load RT-DETR conjoined model: {e}") from e
            self.models[ModelType.RTDETR_CONJOINED_BUBBLE] = model
            return model

    def _load_tree_detector(self, mt: ModelType, what: str, family: str, scale: str, names: dict, env: str):
        """Panel (YOLO11-L, reference :809-833) and OSB-text (YOLO12x, :780-807) detectors: executed from the module tree
        the ultralytics `.pt` pickles (mangatranslator_b200/yolo_tree.py), no architecture assumed.  Like the RT-DETR
        model they are optional in the reference's flow, so without a checkpoint they only load when explicitly asked
        for (seeded synthetic tree of the published layout and scale)."""
        with self._lock:
            if self.models.get(mt) is not None:
                return self.models[mt]
            from mangatranslator_b200.yolo_tree import YoloTreeB200, synthetic_tree
            path = self.model_paths[mt]
            if path.is_file():
                try:
                    tree = W.load_ultralytics_tree(str(path))
                except W.UnsupportedCheckpoint as e:
                    raise ModelError(f"{what}: cannot use {path}: {e}") from e
                except Exception as e:
                    raise ModelError(f"{what}: failed to read {path}: {e}") from e
                log_message(f"{what}: module tree and weights from {path} ({len(tree['layers'])} layers)", always_print=True)
            elif os.environ.get(env, "0") == "1":
                log_message(f"{what}: NO CHECKPOINT ({path}) - seeded SYNTHETIC YOLO{family}{scale} tree ({env}=1); outputs are "
                            "meaningless on real pages", always_print=True)
                tree = synthetic_tree(family, scale, nc=len(names), seed=self.synthetic_seed, names=names)
            else:
                raise ModelError(f"{what}: no checkpoint at {path} (set {env}=1 for a seeded synthetic model)")
            dev = self._require_cuda()
            try:
                self.models[mt] = YoloTreeB200(tree, dev, precision=self.precision)
            except W.UnsupportedCheckpoint as e:
                raise ModelError(f"{what}: cannot use {path}: {e}") from e
            return self.models[mt]

    def load_yolo_osbtext(self, token: Optional[str] = None, verbose: bool = False):
        log_message("Loading YOLO OSB Text detection model...", verbose=verbose)
        return self._load_tree_detector(ModelType.YOLO_OSBTEXT, "OSB text detector", "12", "x", {0: "text"},
                                        "MTB200_SYNTHETIC_OSBTEXT")

    def load_yolo_panel(self, verbose: bool = False):
        log_message("Loading YOLO panel detection model...", verbose=verbose)
        # deepghs/manga109_yolo classes; `detect_panels` keeps the "frame" class (core/image/detection.py:1878-1895)
        return self._load_tree_detector(ModelType.YOLO_PANEL, "Panel detector", "11", "l",
                                        {0: "body", 1: "face", 2: "frame", 3: "text"}, "MTB200_SYNTHETIC_PANEL")

    def load_sam3(self, token: str = "", verbose: bool = False):
        self._out_of_scope("SAM3")

    # ---- lifecycle ---------------------------------------------------------------------------------------------
    def unload_model(self, model_type: ModelType, force_gc: bool = True, verbose: bool = False) -> None:
        with self._lock:
            self.models.pop(model_type, None)
        if force_gc:
            self.clear_cache()

    def unload_ocr_models(self, verbose: bool = False) -> None:
        """Same set as the reference (:1397-1432): what it calls "OCR-related" includes the detectors and SAM."""
        group = (ModelType.YOLO_SPEECH_BUBBLE, ModelType.YOLO_SPEECH_BUBBLE_2, ModelType.RTDETR_CONJOINED_BUBBLE,
                 ModelType.SAM2, ModelType.SAM3, ModelType.YOLO_OSBTEXT, ModelType.YOLO_PANEL, ModelType.MANGA_OCR,
                 ModelType.PADDLE_OCR_VL)
        had = [t for t in group if self.is_loaded(t)]
        for t in group:
            self.unload_model(t, force_gc=False, verbose=verbose)
        self.clear_cache()
        if had:
            log_message("OCR models unloaded.", verbose=verbose)

    def unload_upscale_models(self, verbose: bool = False) -> None:
        self.unload_model(ModelType.UPSCALE, verbose=verbose)
        self.unload_model(ModelType.UPSCALE_LITE, verbose=verbose)

    def unload_all(self, verbose: bool = False) -> None:
        with self._lock:
            self.models.clear()
        self.clear_cache()

    def clear_cache(self) -> None:
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def get_memory_stats(self) -> dict:
        """The reference's device report (:1495-1497 -> core/device.py get_device_info) plus the loaded model names."""
        from mangatranslator_b200.core.device import get_device_info
        stats = get_device_info(self.device if isinstance(self.device, torch.device) else None)
        stats["loaded_models"] = [m.value for m, v in self.models.items() if v is not None]
        return stats

    def print_memory_stats(self) -> None:
        stats = self.get_memory_stats()
        if stats.get("memory") == "N/A":
            log_message(f"Device: {stats['device']}", always_print=True)
        else:
            log_message(f"GPU Memory - Allocated: {stats['allocated_gb']} GB, Reserved: {stats['reserved_gb']} GB",
                        always_print=True)


def get_model_manager() -> ModelManager:
    return ModelManager()
