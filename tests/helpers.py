"""Shared test helpers: synthetic cases named in tests/golden/clean_golden.json, hashing, host-emulation loader."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_clean_golden():
    with open(os.path.join(ROOT, "tests", "golden", "clean_golden.json")) as f:
        return json.load(f)


def build_clean_case(g):
    """Recreate the inputs of one golden case (same seeded generator the fixture script used)."""
    from mangatranslator_b200 import synth
    page = synth.make_page(g["seed"], g["H"], g["W"], n_bubbles=g["n_bubbles"])
    dets = synth.detections_from_page(page)
    if g.get("conjoined"):
        dets[0]["conjoined_neighbor_bboxes"] = [dets[1]["bbox"]]
        dets[1]["conjoined_neighbor_bboxes"] = [dets[0]["bbox"]]
    bgr = np.ascontiguousarray(page.image_rgb[:, :, ::-1])
    if g["rgba"]:
        bgr = np.ascontiguousarray(np.concatenate([bgr, np.full(bgr.shape[:2] + (1,), 255, np.uint8)], axis=2))
    return bgr, dets


def check_bubbles_against_golden(bubbles, g):
    assert len(bubbles) == len(g["bubbles"]), (len(bubbles), len(g["bubbles"]))
    for b, gb in zip(bubbles, g["bubbles"]):
        assert [int(v) for v in b["bbox"]] == gb["bbox"]
        assert [int(v) for v in b["color"]] == gb["color"]
        assert [int(v) for v in b["text_bbox"]] == gb["text_bbox"]
        tc = None if b["text_color_bgr"] is None else [int(v) for v in b["text_color_bgr"]]
        assert tc == gb["text_color_bgr"]
        assert int((b["mask"] > 0).sum()) == gb["mask_pixels"]
        assert sha(b["mask"]) == gb["mask_sha256"]


_emul = None


def clean_emul_lib():
    """Builds (g++) and loads the sequential host build of the cleaning kernel logic (test infrastructure)."""
    global _emul
    if _emul is None:
        src = os.path.join(ROOT, "tests", "host_emul", "clean_emul.cpp")
        so = os.path.join(ROOT, "tests", "host_emul", "libclean_emul.so")
        hdr = os.path.join(ROOT, "mangatranslator_b200", "csrc", "clean_core.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src])
        _emul = C.CDLL(so)
        _emul.emul_workspace_words.restype = C.c_ulonglong
    return _emul


def emul_clean_bubble(img_bgr, mask, bbox, params, neighbors=None):
    """Run one bubble through the host emulation; returns (CleanResult, final-mask HxW uint8)."""
    from mangatranslator_b200 import clean_host as H
    E = clean_emul_lib()
    h, w, c = img_bgr.shape
    ys, xs = np.nonzero(mask)
    mb = (int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1) if len(xs) else (0, 0, 1, 1)
    wx0, wy0, cw, ch = H.plan_window(mb, w, h, params)
    J = H.CleanJob()
    img_c = np.ascontiguousarray(img_bgr)
    m_c = np.ascontiguousarray(mask)
    J.img, J.img_pitch, J.img_h, J.img_w, J.img_c = img_c.ctypes.data, w * c, h, w, c
    J.mask, J.mask_pitch, J.mask_x0, J.mask_y0, J.mask_w, J.mask_h = m_c.ctypes.data, w, 0, 0, w, h
    J.wx0, J.wy0, J.cw, J.ch = wx0, wy0, cw, ch
    for i in range(4):
        J.bbox[i] = int(bbox[i])
    nb = neighbors or []
    J.n_neighbors = len(nb)
    for k, b in enumerate(nb):
        for i in range(4):
            J.neighbors[k][i] = int(b[i])
    mr = H.default_max_runs(cw, ch)
    words = H.workspace_words(cw, ch, mr)
    assert words == E.emul_workspace_words(cw, ch, mr)
    work = np.zeros(words, np.uint32)
    J.work, J.max_runs, J.page_index = work.ctypes.data, mr, 0
    R = H.CleanResult()
    E.emul_clean_job(C.byref(params), C.byref(J), C.byref(R))
    cwords = (cw + 31) // 32
    p = work[H.PLANE_FINAL * cwords * ch:(H.PLANE_FINAL + 1) * cwords * ch].reshape(ch, cwords)
    bits = np.unpackbits(p.view(np.uint8).reshape(ch, cwords * 4), axis=1, bitorder="little")[:, :cw]
    full = np.zeros((h, w), np.uint8)
    full[wy0:wy0 + ch, wx0:wx0 + cw] = bits * 255
    return R, full
