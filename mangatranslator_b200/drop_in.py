"""`install()` aliases this build's modules under the reference's package names (`core`, `utils`) so the reference's
entry points (`app.py`, `main.py`, `ui/*`) import the B200 hot path unchanged.  Same module objects under both names,
so singletons (ModelManager, cache) are shared."""
import importlib
import sys

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
