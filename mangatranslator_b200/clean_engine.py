"""Batched launcher of the bubble-cleaning kernels (C ABI: mtb_clean_bubbles / mtb_clean_paint / mtb_clean_export_mask).

Plans one job per (page, detection): crop window, workspace slice, neighbour boxes; uploads the job table; runs ONE
kernel launch for all bubbles of all pages of the batch, then the colour-grouped fill.  Pixels never visit the host
here: pages and masks are device tensors (numpy inputs are uploaded once).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import clean_host as H
from ._lib import check, lib, ptr, stream_ptr

STATUS_NAMES = {0: "ok", 1: "empty mask", 2: "no valid contour", 3: "window too small", 4: "workspace overflow"}


def _declare(l) -> None:
    if getattr(l, "_clean_declared", False):
        return
    vp, i32 = C.c_void_p, C.c_int
    l.mtb_clean_workspace_words.argtypes = [i32, i32, i32]
    l.mtb_clean_workspace_words.restype = C.c_ulonglong
    l.mtb_clean_bubbles.argtypes = [C.POINTER(H.CleanParams), vp, vp, i32, vp]
    l.mtb_clean_bubbles.restype = i32
    l.mtb_clean_paint.argtypes = [vp, vp, i32, vp, vp, i32, vp]
    l.mtb_clean_paint.restype = i32
    l.mtb_clean_export_mask.argtypes = [vp, i32, vp, C.c_longlong, i32, vp]
    l.mtb_clean_export_mask.restype = i32
    l._clean_declared = True


def _mask_bbox_host(m: np.ndarray):
    rows = np.flatnonzero(m.any(axis=1))
    if rows.size == 0:
        return None
    cols = np.flatnonzero(m.any(axis=0))
    return int(cols[0]), int(rows[0]), int(cols[-1]) + 1, int(rows[-1]) + 1


@dataclass
class CleanBatchResult:
    pages_out: List[torch.Tensor]                      # cleaned pages (device, HxWxC uint8)
    results: List[List[Optional[H.CleanResult]]]       # per page, per detection (None = skipped before launch)
    jobs_dev: Optional[torch.Tensor] = None
    job_index: List[List[int]] = field(default_factory=list)   # per page, per detection -> job slot (-1 = none)
    keep: List[Any] = field(default_factory=list)

    def export_mask(self, page: int, det: int, plane: int = H.PLANE_FINAL) -> torch.Tensor:
        """Full-frame uint8 {0,255} mask of one job (device tensor)."""
        j = self.job_index[page][det]
        if j < 0:
            raise IndexError("detection was not submitted")
        h, w = self.pages_out[page].shape[:2]
        out = torch.empty((h, w), dtype=torch.uint8, device=self.pages_out[page].device)
        check(lib().mtb_clean_export_mask(ptr(self.jobs_dev), j, ptr(out), w, plane, stream_ptr()),
              "mtb_clean_export_mask")
        return out

    def text_boxes(self, padding_pixels: float = 4.0):
        """Safe text box of every cleaned bubble (what the reference's renderer computes per bubble from the `mask` entry
        of clean_speech_bubbles' records: core/pipeline.py:1657 -> core/text/text_renderer.py:162 ->
        core/image/image_utils.py:173-348).  All bubbles of all pages in ONE launch on the exported final masks.
        Returns, per page and detection: ((x, y, w, h), (cx, cy)), the reference's error message (str) when the safe-area
        calculation fails for that bubble (the renderer then falls back to the padded bbox), or None for detections
        whose cleaning failed / was skipped."""
        from . import safebox_host as S
        slots, masks = [], []
        for pi, row in enumerate(self.results):
            for di, r in enumerate(row):
                if r is not None and r.status == 0:
                    slots.append((pi, di))
                    masks.append(self.export_mask(pi, di))
        out = [[None] * len(row) for row in self.results]
        if not masks:
            return out
        for (pi, di), rec in zip(slots, S.safe_boxes_device(masks, padding_pixels)):
            try:
                out[pi][di] = S.decode(rec)
            except ValueError as e:
                out[pi][di] = e.args[0]
        return out


def clean_batch(pages: Sequence[torch.Tensor], detections: Sequence[Sequence[Dict[str, Any]]],
                params: H.CleanParams, *, in_place: bool = False, max_color_ranks: int = 2) -> CleanBatchResult:
    """pages[i]: device uint8 HxWxC (BGR or BGRA).  detections[i][k]: {'bbox': (x0,y0,x1,y1), 'sam_mask': HxW uint8
    numpy array or device tensor, optional 'mask_bbox', optional 'conjoined_neighbor_bboxes'}."""
    l = lib()
    _declare(l)
    dev = pages[0].device
    jobs: List[H.CleanJob] = []
    job_index: List[List[int]] = []
    keep: List[Any] = []
    work_sizes: List[int] = []
    for pi, (page, dets) in enumerate(zip(pages, detections)):
        assert page.dtype == torch.uint8 and page.dim() == 3 and page.is_contiguous()
        h, w, c = page.shape
        idx_row = []
        for det in dets:
            m = det.get("sam_mask")
            if m is None:
                idx_row.append(-1)
                continue
            mb = det.get("mask_bbox")
            if isinstance(m, np.ndarray):
                if mb is None:
                    mb = _mask_bbox_host(m)
                mt = torch.from_numpy(np.ascontiguousarray(m.astype(np.uint8, copy=False))).to(dev, non_blocking=True)
            else:
                mt = m.contiguous()
                if mb is None:
                    nz = torch.nonzero(mt, as_tuple=False)
                    mb = None if nz.numel() == 0 else (int(nz[:, 1].min()), int(nz[:, 0].min()),
                                                       int(nz[:, 1].max()) + 1, int(nz[:, 0].max()) + 1)
            if mb is None:
                mb = (0, 0, 1, 1)  # empty mask: the kernel reports status 1 like the reference raises
            keep.append(mt)
            wx0, wy0, cw, ch = H.plan_window(mb, w, h, params)
            J = H.CleanJob()
            J.img, J.img_pitch, J.img_h, J.img_w, J.img_c = page.data_ptr(), w * c, h, w, c
            J.mask, J.mask_pitch = mt.data_ptr(), mt.shape[1]
            J.mask_x0, J.mask_y0, J.mask_w, J.mask_h = 0, 0, mt.shape[1], mt.shape[0]
            J.wx0, J.wy0, J.cw, J.ch = wx0, wy0, cw, ch
            bb = det.get("bbox") or mb
            for i in range(4):
                J.bbox[i] = int(bb[i])
            nbs = list(det.get("conjoined_neighbor_bboxes") or [])
            if len(nbs) > H.MAX_NEIGHBORS:
                raise ValueError(f"bubble with {len(nbs)} conjoined neighbours (max {H.MAX_NEIGHBORS})")
            J.n_neighbors = len(nbs)
            for k, nb in enumerate(nbs):
                for i in range(4):
                    J.neighbors[k][i] = int(nb[i])
            J.max_runs = H.default_max_runs(cw, ch)
            J.page_index = pi
            work_sizes.append(H.workspace_words(cw, ch, J.max_runs))
            idx_row.append(len(jobs))
            jobs.append(J)
        job_index.append(idx_row)
    pages_out = list(pages) if in_place else [p.clone() for p in pages]
    n = len(jobs)
    if n == 0:
        return CleanBatchResult(pages_out, [[None] * len(d) for d in detections], None, job_index, keep)
    work = torch.empty(int(sum(work_sizes)), dtype=torch.int32, device=dev)
    off = 0
    for J, sz in zip(jobs, work_sizes):
        J.work = work.data_ptr() + 4 * off
        off += sz
    arr = (H.CleanJob * n)(*jobs)
    jobs_host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    jobs_dev = jobs_host.to(dev)
    res_dev = torch.zeros(n * C.sizeof(H.CleanResult), dtype=torch.uint8, device=dev)
    st = stream_ptr()
    check(l.mtb_clean_bubbles(C.byref(params), ptr(jobs_dev), ptr(res_dev), n, st), "mtb_clean_bubbles")
    page_ptrs = torch.tensor([p.data_ptr() for p in pages_out], dtype=torch.int64, device=dev)
    rank = torch.empty(n, dtype=torch.int32, device=dev)
    check(l.mtb_clean_paint(ptr(jobs_dev), ptr(res_dev), n, ptr(page_ptrs), ptr(rank), int(max_color_ranks), st),
          "mtb_clean_paint")
    res_host = res_dev.cpu().numpy().tobytes()
    rs = (H.CleanResult * n).from_buffer_copy(res_host)
    results: List[List[Optional[H.CleanResult]]] = []
    for row in job_index:
        results.append([rs[j] if j >= 0 else None for j in row])
    keep.extend([work, page_ptrs, rank, res_dev])
    return CleanBatchResult(pages_out, results, jobs_dev, job_index, keep)
