"""Timing of the safe-text-box path on one 1536x1024 page (12 full-frame bubble masks): CUDA events around
safe_boxes_device (job upload + memset + bounds + boxes + result D2H), and around the three stream operations alone.
    python tools/prof_safebox.py            # prints one JSON line"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import safebox_page_masks  # noqa: E402
from mangatranslator_b200 import safebox_host as S  # noqa: E402
from mangatranslator_b200._lib import lib, stream_ptr  # noqa: E402

dev = torch.device("cuda")
masks = [torch.from_numpy(m).to(dev) for m in safebox_page_masks(1)]
n = len(masks)
for _ in range(3):
    recs = S.safe_boxes_device(masks, 6.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    S.safe_boxes_device(masks, 6.0)
e1.record()
torch.cuda.synchronize()
call_ms = e0.elapsed_time(e1) / reps
# the launches alone on a prepared job table
h, w = masks[0].shape
cap = S.window_cap(h, w)
g = torch.empty(n * cap, dtype=torch.int16, device=dev)
safe = torch.empty(n * cap, dtype=torch.uint8, device=dev)
jobs = (S.SafeBoxJob * n)()
for i, m in enumerate(masks):
    j = jobs[i]
    j.mask, j.pitch, j.H, j.W, j.t2, j.cap = m.data_ptr(), m.stride(0), h, w, S.threshold_sq(6.0), cap
    j.g, j.safe = g.data_ptr() + 2 * i * cap, safe.data_ptr() + i * cap
jd = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(dev)
rd = torch.empty(n * C.sizeof(S.SafeBoxResult), dtype=torch.uint8, device=dev)
for _ in range(3):
    lib().mtb_safe_boxes(C.c_void_p(jd.data_ptr()), C.c_void_p(rd.data_ptr()), n, C.c_void_p(stream_ptr()))
e0.record()
for _ in range(reps):
    lib().mtb_safe_boxes(C.c_void_p(jd.data_ptr()), C.c_void_p(rd.data_ptr()), n, C.c_void_p(stream_ptr()))
e1.record()
torch.cuda.synchronize()
launch_ms = e0.elapsed_time(e1) / reps
print(json.dumps({"workload": f"{n} full-frame masks {h}x{w}, padding 6", "call_ms": round(call_ms, 3),
                  "launches_ms": round(launch_ms, 3), "mask_bytes": n * h * w,
                  "ok": int((recs["status"] == 0).sum())}))
