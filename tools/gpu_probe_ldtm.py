"""Which (TMEM lane, column) lands in which (thread, register) of tcgen05.ld.16x256b.x2?  Expected (the m16n8 accumulator
fragment): thread T, register 4g + 2h + e  ->  lane base + T/4 + 8h, column 8g + 2(T%4) + e.  Writes gpurun_out/ldtm_layout.json."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mangatranslator_b200 import _lib  # noqa: E402

lib = _lib.exp_lib()
lib.mtb_exp_ldtm_layout.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.mtb_exp_ldtm_layout.restype = C.c_int
res = {}
for off in (0, 16):
    out = torch.zeros(4, 32, 8, device="cuda")
    assert lib.mtb_exp_ldtm_layout(out.data_ptr(), off, _lib.stream_ptr()) == 0
    torch.cuda.synchronize()
    o = out.cpu().long()
    lane, col = o // 1000, o % 1000
    ok = True
    for w in range(4):
        for t in range(32):
            for g in range(2):
                for h in range(2):
                    for e in range(2):
                        r = 4 * g + 2 * h + e
                        ok &= int(lane[w, t, r]) == 32 * w + off + t // 4 + 8 * h and int(col[w, t, r]) == 8 * g + 2 * (t % 4) + e
    res[f"lane_off_{off}"] = {"matches_m16n8_fragment": bool(ok), "warp0_thread0": [[int(lane[0, 0, r]), int(col[0, 0, r])] for r in range(8)],
                             "warp1_thread5": [[int(lane[1, 5, r]), int(col[1, 5, r])] for r in range(8)]}
    print(off, res[f"lane_off_{off}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ldtm_layout.json"), "w"), indent=1)
