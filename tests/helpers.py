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
        hdr2 = os.path.join(ROOT, "mangatranslator_b200", "csrc", "hd_emul.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in (src, hdr, hdr2)):
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


# ---- safe text box (calculate_centroid_expansion_box) cases -----------------------------------------------------------
SAFEBOX_KINDS = ["ellipse", "conjoined", "crescent", "border", "holes", "thin", "full", "speck", "bent", "dumbbell"]
SAFEBOX_PADDINGS = [4.0, 1.0, 2.5, 7.3, 10.0, 0.5, 15.0, 0.0]


def safebox_mask(seed: int, kind: str | None = None, size=None):
    """Seeded bubble-like uint8 {0,255} mask exercising one branch family of the reference function each: plain
    ellipse (centroid anchor), two blobs joined by a neck (pole of inaccessibility), crescent (centroid outside the
    safe area), bent strip (nearest safe pixel without the pole rule), shapes cut by the image border, pin-holes, thin strips (failures), a full frame,
    a few-pixel speck.  Returns (mask, padding_pixels)."""
    import cv2
    rng = np.random.default_rng(1000 + seed)
    kind = kind or SAFEBOX_KINDS[seed % len(SAFEBOX_KINDS)]
    h, w = size if size else (int(rng.integers(30, 420)), int(rng.integers(30, 420)))
    pad = SAFEBOX_PADDINGS[int(rng.integers(0, len(SAFEBOX_PADDINGS)))]
    m = np.zeros((h, w), np.uint8)
    if kind == "bent":      # a V-shaped strip barely wider than 2 x padding: the centroid falls beside the safe area
        t = int(2 * pad) + 1 + int(rng.integers(0, 4))
        pts = np.array([[t + 2, t + 2], [w // 2, h - t - 3], [w - t - 3, t + 2 + int(rng.integers(0, max(1, h // 3)))]], np.int32)
        cv2.polylines(m, [pts], False, 255, t)
    elif kind == "dumbbell":  # two discs of radius ~1.1 x padding joined by a neck just under 2 x padding wide: the
        pad = [7.3, 10.0, 15.0][int(rng.integers(0, 3))]      # centroid lies in the neck, unsafe but above 0.7 x max
        r, nk = int(round(1.1 * pad)), int(np.ceil(pad)) - 3
        h, w = max(h, 2 * r + 8), max(w, 6 * r + 12)
        m = np.zeros((h, w), np.uint8)
        y, xa, xb = h // 2, r + 3, w - r - 4 - int(rng.integers(0, 3))
        cv2.circle(m, (xa, y), r, 255, -1)
        cv2.circle(m, (xb, y + int(rng.integers(-2, 3))), r, 255, -1)
        cv2.line(m, (xa, y), (xb, y), 255, 2 * nk + 1)
    elif kind == "ellipse":
        cv2.ellipse(m, (w // 2 + int(rng.integers(-w // 6, w // 6 + 1)), h // 2), (max(3, w // 3), max(3, h // 3)),
                    float(rng.integers(0, 180)), 0, 360, 255, -1)
    elif kind == "conjoined":
        cv2.circle(m, (w // 4, h // 2), max(3, min(w, h) // 5), 255, -1)
        cv2.circle(m, (3 * w // 4, h // 2 + int(rng.integers(-h // 6, h // 6 + 1))), max(3, min(w, h) // 4), 255, -1)
        cv2.line(m, (w // 4, h // 2), (3 * w // 4, h // 2), 255, int(rng.integers(1, 12)))
    elif kind == "crescent":
        cv2.circle(m, (w // 2, h // 2), max(4, min(w, h) // 2 - 2), 255, -1)
        cv2.circle(m, (w // 2 + int(rng.integers(0, w // 5 + 1)), h // 2), max(2, min(w, h) // 3), 0, -1)
    elif kind == "border":
        cv2.ellipse(m, (int(rng.integers(0, w)), int(rng.integers(0, h))), (max(3, w // 2), max(3, h // 2)), 0, 0, 360,
                    255, -1)
    elif kind == "holes":
        cv2.ellipse(m, (w // 2, h // 2), (max(3, w // 2 - 2), max(3, h // 2 - 2)), 0, 0, 360, 255, -1)
        m[rng.random((h, w)) < 0.002] = 0
    elif kind == "thin":
        cv2.rectangle(m, (2, 2), (min(w - 1, 2 + int(rng.integers(1, 14))), min(h - 1, 2 + int(rng.integers(1, 40)))),
                      255, -1)
    elif kind == "full":
        m[:] = 255
    else:  # speck
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        m[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4))] = 255
    return m, pad


def safebox_page_masks(seed: int, h: int = 1536, w: int = 1024, n: int = 12):
    """Full-frame masks of `n` bubbles on a 3 x 4 grid of an h x w page (one conjoined pair among them)."""
    import cv2
    rng = np.random.default_rng(5000 + seed)
    out = []
    for i in range(n):
        m = np.zeros((h, w), np.uint8)
        cx, cy = int((i % 3 + 0.5) * w / 3 + rng.integers(-20, 21)), int((i // 3 + 0.5) * h / 4 + rng.integers(-20, 21))
        ax, ay = int(rng.integers(w // 10, w // 6)), int(rng.integers(h // 14, h // 9))
        cv2.ellipse(m, (cx, cy), (ax, ay), float(rng.integers(-20, 21)), 0, 360, 255, -1)
        if i == 5:
            cv2.ellipse(m, (cx + ax, cy + ay // 2), (ax // 2, ay // 2), 0, 0, 360, 255, -1)
        out.append(m)
    return out


_sb_emul = None


def safebox_emul_lib():
    """Builds (g++) and loads the sequential host build of the safe-box kernel logic (test infrastructure)."""
    global _sb_emul
    if _sb_emul is None:
        src = os.path.join(ROOT, "tests", "host_emul", "safebox_emul.cpp")
        so = os.path.join(ROOT, "tests", "host_emul", "libsafebox_emul.so")
        deps = [src] + [os.path.join(ROOT, "mangatranslator_b200", "csrc", f) for f in ("safebox_core.cuh", "hd_emul.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src])
        _sb_emul = C.CDLL(so)
    return _sb_emul


def emul_safe_box(mask, padding_pixels, t2=None, cap=None):
    """Run one mask through the host emulation of the kernel; returns the SafeBoxResult record."""
    from mangatranslator_b200 import safebox_host as S
    import safebox_oracle as O
    E = safebox_emul_lib()
    assert E.emul_safebox_sizeof(0) == C.sizeof(S.SafeBoxJob) and E.emul_safebox_sizeof(1) == C.sizeof(S.SafeBoxResult)
    m = np.ascontiguousarray(mask)
    h, w = m.shape
    J, R = S.SafeBoxJob(), S.SafeBoxResult()
    cap = S.window_cap(h, w) if cap is None else cap
    g, safe = np.zeros(cap, np.uint16), np.zeros(cap, np.uint8)
    J.mask, J.pitch, J.H, J.W = m.ctypes.data, w, h, w
    J.t2 = O.threshold_sq(padding_pixels) if t2 is None else t2
    J.cap, J.g, J.safe = cap, g.ctypes.data, safe.ctypes.data
    E.emul_safebox_job(C.byref(J), C.byref(R))
    return R


def check_page_against_cpu_pipeline(pipe, ref, out_u8, dets, batch, label=""):
    """End-to-end comparison of one page (device pipeline vs oracle/pipeline_oracle.CpuPipeline.run_page) with counts:
    * mask bits: identical outside the oracle's knife-edge band (|full-size logit| < 2e-3 for some prompt); the number of
      flipped bits inside the band is printed and bounded by a NUMBER;
    * cleaned page: byte-identical when no mask bit flipped; otherwise differing pixels must lie within the cleaning
      reach (ROI dilation) of a flipped bit's bubble — checked as: every differing pixel is inside some detection's box
      grown by 16 px;
    * upscale: the RCAN is judged in isolation on the ORACLE's cleaned page (its global average pooling couples every
      output pixel to every input pixel, so an end-to-end float bound would be void after a single flipped mask bit):
      float output within 1e-3 abs everywhere, uint8 within 1 LSB; and when the cleaned pages are identical the
      end-to-end uint8 page must satisfy the same 1-LSB bound."""
    import torch
    got_masks = np.stack([d["sam_mask"].cpu().numpy() for d in dets])
    assert got_masks.shape == ref["masks"].shape
    flips = got_masks != ref["masks"]
    n_flips = int(flips.sum())
    outside = int((flips & ~ref["band"][None]).sum())
    cleaned = batch.pages_out[0].cpu().numpy()
    n_clean_diff = int(np.any(cleaned != ref["cleaned"], axis=2).sum())
    src = torch.from_numpy(ref["cleaned"]).to(pipe.device)
    iso_u8, iso_f = pipe.rcan.upscale_u8(src, swap_rb=True, want_float=True)
    torch.cuda.synchronize()
    e_f = float((iso_f.permute(2, 0, 1).unsqueeze(0).cpu() - ref["upscaled_f"]).abs().max())
    d_iso = np.abs(iso_u8.cpu().numpy().astype(int) - ref["upscaled"].astype(int))
    print(f"{label}: mask bits flipped {n_flips} ({outside} outside the knife-edge band of {int(ref['band'].sum())} pixels); "
          f"cleaned pixels differing {n_clean_diff}; RCAN on the oracle's cleaned page: float max abs {e_f:.2e}, "
          f"uint8 values off by one {int((d_iso > 0).sum())}")
    assert outside == 0
    assert n_flips <= 64, n_flips
    if n_flips == 0:
        assert n_clean_diff == 0
    assert e_f < 1e-3 and d_iso.max() <= 1
    if n_clean_diff == 0:
        d = np.abs(out_u8.numpy().astype(int) - ref["upscaled"].astype(int))
        assert d.max() <= 1
    return n_flips, n_clean_diff
