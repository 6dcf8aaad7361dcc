"""Pre-processing parity: cv2.resize(INTER_LINEAR) letterbox (ultralytics LetterBox) and torchvision/ATen uint8
antialias-bilinear resize (Sam2 image processor) — the restated algorithms (CPU) and the CUDA kernels (GPU) are
bit-exact against the libraries themselves; likewise Pillow's LANCZOS resample (the exact-size step after the RCAN passes
and resize_to_min_side, core/image/image_utils.py:545,551-595)."""
import ctypes as C

import cv2
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _aa_tables(n_in, n_out):
    from mangatranslator_b200 import _lib
    l = _lib.lib()
    l.mtb_aa_weights_host.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    cap = 64
    st, ln = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
    w = np.zeros((n_out, cap), np.int16)
    k, p = C.c_int(), C.c_int()
    assert l.mtb_aa_weights_host(n_in, n_out, st.ctypes.data, ln.ctypes.data, w.ctypes.data, cap, C.byref(k), C.byref(p)) == 0
    return st, ln, w, p.value


def _aa_axis(img, axis, n_out):
    st, ln, w, p = _aa_tables(img.shape[axis], n_out)
    x = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((n_out,) + x.shape[1:], np.int64)
    for o in range(n_out):
        acc = np.full(x.shape[1:], 1 << (p - 1), np.int64)
        for j in range(ln[o]):
            acc += int(w[o, j]) * x[st[o] + j]
        out[o] = np.clip(acc >> p, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


@pytest.mark.parametrize("hw", [(1536, 1024), (1150, 800), (768, 1024), (300, 500)])
def test_aa_resize_algorithm_matches_torch_uint8(hw):
    h, w = hw
    img = np.random.default_rng(h).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0)
    ref = F.interpolate(t, size=(1024, 1024), mode="bilinear", antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    a = _aa_axis(img, 1, 1024) if w != 1024 else img
    b = _aa_axis(a, 0, 1024) if h != 1024 else a
    assert np.array_equal(b, ref)


def _cv_linear_numpy(img, nw, nh):
    """OpenCV 8-bit INTER_LINEAR restated (resize.cpp: 11-bit coefficients, two-pass fixed point)."""
    sh, sw = img.shape[:2]
    sx_scale, sy_scale = 1.0 / (nw / sw), 1.0 / (nh / sh)
    dx = np.arange(nw)
    fx = ((dx + 0.5) * sx_scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx.astype(np.float32)
    fx[sx < 0] = 0
    sx[sx < 0] = 0
    m = sx >= sw - 1
    fx[m] = 0
    sx[m] = sw - 1
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    sx1 = np.minimum(sx + 1, sw - 1)
    dy = np.arange(nh)
    fy = ((dy + 0.5) * sy_scale - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int64)
    fy = fy - sy.astype(np.float32)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    s0, s1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    im = img.astype(np.int64)
    hrow = im[:, sx] * a0[None, :, None] + im[:, sx1] * a1[None, :, None]
    v = (((b0[:, None, None] * (hrow[s0] >> 4)) >> 16) + ((b1[:, None, None] * (hrow[s1] >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("case", [((1536, 1024), (1067, 1600)), ((768, 1024), (1600, 1200)), ((1150, 800), (445, 640)),
                                  ((400, 300), (480, 640))])
def test_cv2_linear_algorithm_matches_cv2(case):
    (h, w), (nw, nh) = case
    img = np.random.default_rng(w).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(_cv_linear_numpy(img, nw, nh), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("hw_sz", [((1536, 1024), 1600), ((768, 1024), 1600), ((1150, 800), 640), ((640, 640), 640)])
def test_letterbox_kernel_matches_cv2(hw_sz):
    from mangatranslator_b200.preproc import letterbox_device, letterbox_geometry
    (h, w), imgsz = hw_sz
    img = np.random.default_rng(h + w).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    (nw, nh), (t, b, l, r), _, _ = letterbox_geometry(h, w, imgsz)
    ref = img if (nw, nh) == (w, h) else cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    ref = cv2.copyMakeBorder(ref, t, b, l, r, cv2.BORDER_CONSTANT, value=(114, 114, 114))[..., ::-1]
    got = letterbox_device(torch.from_numpy(img).cuda(), imgsz, swap_rb=True).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(1536, 1024), (1150, 800), (768, 1024)])
def test_aa_resize_kernel_matches_torch(hw):
    from mangatranslator_b200.preproc import resize_aa_device
    h, w = hw
    img = np.random.default_rng(h).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0)
    ref = F.interpolate(t, size=(1024, 1024), mode="bilinear", antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    got = resize_aa_device(torch.from_numpy(img).cuda(), 1024, 1024).cpu().numpy()
    assert np.array_equal(got, ref)


# ---- Pillow LANCZOS (Resample.c: 22-bit coefficients, uint8 intermediate) ---------------------------------------------
def _lanczos_tables(n_in, n_out):
    from mangatranslator_b200 import _lib
    l = _lib.lib()
    l.mtb_lanczos_weights_host.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    cap = int(np.ceil(3.0 * max(n_in / n_out, 1.0))) * 2 + 1
    st, ln = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
    w = np.zeros((n_out, cap), np.int32)
    k = C.c_int()
    assert l.mtb_lanczos_weights_host(n_in, n_out, st.ctypes.data, ln.ctypes.data, w.ctypes.data, cap, C.byref(k)) == 0
    assert k.value == cap
    return st, ln, w


def _lanczos_axis(img, axis, n_out):
    st, ln, w = _lanczos_tables(img.shape[axis], n_out)
    x = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((n_out,) + x.shape[1:], np.int64)
    for o in range(n_out):
        acc = np.full(x.shape[1:], 1 << 21, np.int64)
        for j in range(ln[o]):
            acc += int(w[o, j]) * x[st[o] + j]
        assert np.abs(acc).max() < 2 ** 31            # the kernel accumulates in int32 like Pillow
        out[o] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def _lanczos_numpy(img, oh, ow):
    a = _lanczos_axis(img, 1, ow) if ow != img.shape[1] else img
    return _lanczos_axis(a, 0, oh) if oh != img.shape[0] else a


LANCZOS_CASES = [((300, 200), (450, 300)), ((257, 391), (200, 304)), ((640, 480), (213, 160)), ((97, 131), (97, 400)),
                 ((120, 90), (311, 90)), ((64, 64), (1, 1)), ((5, 7), (50, 70)), ((1536 // 2, 1024 // 2), (1152, 768))]


@pytest.mark.parametrize("case", LANCZOS_CASES)
def test_lanczos_algorithm_matches_pillow(case):
    from PIL import Image
    (h, w), (oh, ow) = case
    img = np.random.default_rng(h * 7 + w).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(_lanczos_numpy(img, oh, ow), ref)


def test_lanczos_saturating_edges_match_pillow():
    """Black/white step edges drive the negative lobes below 0 and above 255: clip8 on both passes."""
    from PIL import Image
    img = np.zeros((90, 120, 3), np.uint8)
    img[:, 60:] = 255
    img[45:, :, 1] = 255 - img[45:, :, 1]
    for oh, ow in ((135, 180), (61, 77), (180, 77)):
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(_lanczos_numpy(img, oh, ow), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("case", LANCZOS_CASES + [((3072, 2048), (2304, 1536)), ((1536, 1024), (2000, 1333))])
def test_lanczos_kernel_matches_pillow(case):
    from PIL import Image
    from mangatranslator_b200.preproc import resize_lanczos_device
    (h, w), (oh, ow) = case
    img = np.random.default_rng(h * 7 + w).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.LANCZOS))
    got = resize_lanczos_device(torch.from_numpy(img).cuda(), oh, ow).cpu().numpy()
    assert np.array_equal(got, ref)
    # a 4-channel source reads its first three channels
    img4 = np.concatenate([img, np.full((h, w, 1), 7, np.uint8)], axis=2)
    got4 = resize_lanczos_device(torch.from_numpy(img4).cuda(), oh, ow).cpu().numpy()
    assert np.array_equal(got4, ref)


# ---- transparency flattening (Pillow paste with the alpha channel as mask) -------------------------------------------
def _flatten_numpy(rgba, bg=255):
    src, a = rgba[..., :3].astype(np.int64), rgba[..., 3:4].astype(np.int64)
    t = bg * (255 - a) + src * a + 128
    return ((t + (t >> 8)) >> 8).astype(np.uint8)


def _all_alpha_pairs():
    s, m = np.meshgrid(np.arange(256), np.arange(256))
    a = np.zeros((256, 256, 4), np.uint8)
    a[..., 0], a[..., 1], a[..., 2], a[..., 3] = s, 255 - s, (s * 7) % 256, m
    return a


def test_flatten_arithmetic_matches_pillow_exhaustively():
    from PIL import Image
    a = _all_alpha_pairs()                                       # every (value, alpha) pair
    im = Image.fromarray(a, "RGBA")
    bg = Image.new("RGB", im.size, (255, 255, 255))
    bg.paste(im, mask=im.split()[3])
    assert np.array_equal(_flatten_numpy(a), np.asarray(bg))


@pytest.mark.gpu
def test_flatten_kernel_and_convert_image_to_target_mode_match_pillow():
    from PIL import Image
    from mangatranslator_b200.core.image.image_utils import convert_image_to_target_mode
    from mangatranslator_b200.preproc import flatten_alpha_device
    a = _all_alpha_pairs()
    got = flatten_alpha_device(torch.from_numpy(a).cuda()).cpu().numpy()
    assert np.array_equal(got, _flatten_numpy(a))
    rgba = np.random.default_rng(3).integers(0, 256, size=(333, 217, 4), dtype=np.uint8)
    im = Image.fromarray(rgba, "RGBA")
    bg = Image.new("RGB", im.size, (255, 255, 255))
    bg.paste(im, mask=im.split()[3])
    out = convert_image_to_target_mode(im, "RGB")
    assert out.mode == "RGB" and np.array_equal(np.asarray(out), np.asarray(bg))
    la = im.convert("LA")
    bg2 = Image.new("RGB", la.size, (255, 255, 255))
    bg2.paste(la.convert("RGBA"), mask=la.split()[1])
    assert np.array_equal(np.asarray(convert_image_to_target_mode(la, "RGB")), np.asarray(bg2))
    assert convert_image_to_target_mode(im.convert("RGB"), "RGBA").mode == "RGBA"
