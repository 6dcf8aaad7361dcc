"""B200-native vision hot path of MangaTranslator (detect -> segment -> clean -> upscale).

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute runs in hand-written
sm_100a CUDA kernels reached through the C ABI in ``include/mtb200.h`` (``mangatranslator_b200/lib/libmtb200.so``).
There is no CPU fallback: importing :mod:`mangatranslator_b200._lib` without the built library raises.
"""
__version__ = "0.1.0"
