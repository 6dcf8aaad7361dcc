"""GPU: the device PNG encoder (csrc/png_kernels.cu + mangatranslator_b200/png_device.py) — every file must decode with
Pillow to exactly the pixels that went in, for RGB and opaque-RGBA output, odd sizes, flat / noisy / page-like content."""
import io
import zlib

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


def _cases():
    r = np.random.default_rng(5)
    yy, xx = np.mgrid[0:700, 0:520]
    page = np.full((700, 520, 3), 255, np.uint8)
    page[(xx // 7 + yy // 5) % 9 == 0] = 30
    page[100:300, 50:400] = (r.integers(90, 170, (200, 350, 1))).astype(np.uint8)          # screentone-like noise
    return {
        "noise": r.integers(0, 256, (123, 77, 3), dtype=np.uint8),
        "flat": np.full((64, 1000, 3), 255, np.uint8),
        "gradient": np.stack([np.add.outer(np.arange(300), 2 * np.arange(410)) % 256] * 3, -1).astype(np.uint8),
        "one_pixel": np.array([[[1, 2, 3]]], np.uint8),
        "thin": r.integers(0, 256, (1, 5000, 3), dtype=np.uint8),
        "page_like": page,
    }


@pytest.mark.parametrize("name", sorted(_cases()))
@pytest.mark.parametrize("oc", [3, 4])
def test_device_png_decodes_to_the_same_pixels(name, oc):
    from mangatranslator_b200.png_device import PngEncoderB200
    img = _cases()[name]
    enc = PngEncoderB200(torch.device("cuda:0"))
    data = enc.encode(torch.from_numpy(img).cuda(), out_channels=oc)
    back = Image.open(io.BytesIO(data))
    back.load()
    assert back.mode == ("RGBA" if oc == 4 else "RGB") and back.size == (img.shape[1], img.shape[0])
    got = np.asarray(back)
    assert np.array_equal(got[:, :, :3], img)
    if oc == 4:
        assert (got[:, :, 3] == 255).all()
    assert data == enc.encode(torch.from_numpy(img).cuda(), out_channels=oc)          # deterministic bytes


def test_upscaled_page_size_and_compression_against_pillow():
    """A 3072x2048 page: exact pixels, and a file no more than 1.6x Pillow's level-2 file (the device coder has no LZ77
    window beyond repeated-byte runs; the reference's default is level 2)."""
    from mangatranslator_b200 import synth
    from mangatranslator_b200.png_device import PngEncoderB200
    pg = synth.make_page(4, 1536, 1024, n_bubbles=12)
    up = np.asarray(Image.fromarray(pg.image_rgb).resize((2048, 3072), Image.BICUBIC))
    enc = PngEncoderB200(torch.device("cuda:0"))
    dev = torch.from_numpy(up).cuda()
    data = enc.encode(dev, out_channels=4)
    back = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(back, up)
    ref = io.BytesIO()
    Image.fromarray(up).convert("RGBA").save(ref, format="PNG", compress_level=2)
    ratio = len(data) / len(ref.getvalue())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        enc.encode(dev, out_channels=4)
    e1.record()
    torch.cuda.synchronize()
    print(f"3072x2048 RGBA: device PNG {len(data) / 1e6:.2f} MB vs Pillow level 2 {len(ref.getvalue()) / 1e6:.2f} MB "
          f"(ratio {ratio:.2f}); {e0.elapsed_time(e1) / 5:.2f} ms per page including the host table and container")
    assert ratio < 1.6
