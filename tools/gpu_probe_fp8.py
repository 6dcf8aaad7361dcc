"""First GPU step of DESIGN.md §8.0 (fp16 + e5m2-correction body conv).  Needs a B200; writes gpurun_out/fp8_probe.json.

  1. functional: four kind::f16 MMAs (fp16 operands) + four kind::f8f6f4 MMAs (e5m2 operands, M = 128 or 64) accumulated in
     ONE fp32 TMEM tile equal the float64 product (checks both instruction descriptors, the byte layout of 8-bit K-major
     128B-swizzled operands, and that kinds may be mixed on one accumulator);
  2. where an M = 64 instruction puts its 64 rows in the 128 TMEM lanes;
  3. rates: clk per MMA for e5m2 M128/M64 x N240 alone and interleaved with the fp16 MMA on one accumulator.

    python tools/gpu_probe_fp8.py
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mangatranslator_b200 import _lib  # noqa: E402

lib = _lib.exp_lib()
main = _lib.lib()
vp = C.c_void_p
lib.mtb_exp_mixed_kind.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
lib.mtb_exp_mixed_kind.restype = C.c_int
lib.mtb_exp_mma_rate.argtypes = [vp] + [C.c_int] * 6 + [vp]
lib.mtb_exp_mma_rate.restype = C.c_int
dev = torch.device("cuda")
out = {}

torch.manual_seed(0)
a16 = torch.randn(128, 64, device=dev).half()
b16 = torch.randn(64, 64, device=dev).half()
a8 = (torch.randn(128, 128, device=dev) * 0.25).to(torch.float8_e5m2)
b8 = (torch.randn(64, 128, device=dev) * 0.25).to(torch.float8_e5m2)
p16 = a16.double() @ b16.double().T
p8 = a8.double() @ b8.double().T


def run(m8, which, init=0.0):
    d = torch.full((128, 64), init, dtype=torch.float32, device=dev)
    rc = lib.mtb_exp_mixed_kind(a16.data_ptr(), b16.data_ptr(), a8.view(torch.uint8).data_ptr(),
                                b8.view(torch.uint8).data_ptr(), d.data_ptr(), m8, which, _lib.stream_ptr())
    assert rc == 0, main.mtb_last_error()
    torch.cuda.synchronize()
    return d.double()


for name, m8, which, exp in (("fp16_only", 128, 1, p16), ("e5m2_only_m128", 128, 2, p8), ("mixed_m128", 128, 3, p16 + p8)):
    d = run(m8, which)
    out[name] = {"max_abs_err": float((d - exp).abs().max()), "ref_max": float(exp.abs().max())}
    print(name, out[name], flush=True)

# M = 64: which TMEM lane holds which row of the 64-row product?
d = run(64, 2)
lanes = {}
for lane in range(128):
    diff = (p8[:64] - d[lane].unsqueeze(0)).abs().max(dim=1).values
    r = int(diff.argmin())
    if float(diff[r]) < 1e-2:
        lanes[lane] = r
out["m64_lane_to_row"] = lanes
print("M=64 rows found in lanes:", lanes, flush=True)
d = run(64, 3)
rows = sorted(lanes.items())
if rows:
    err = max(float((d[lane] - (p16[lane] + p8[r])).abs().max()) for lane, r in rows)
    out["mixed_m64_on_its_lanes_max_abs_err"] = err      # fp16 M = 128 product + e5m2 M = 64 product, lane by lane
    print("mixed m64 err", err, flush=True)

# tap-shifted views of a 64-byte-swizzled 8-bit tile (the e5m2 activation plane: 64 channels x 1 B per pixel)
lib.mtb_exp_shifted_desc8.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
lib.mtb_exp_shifted_desc8.restype = C.c_int
for kind, kname in ((0, "e5m2"), (1, "fp16")):
    if kind == 0:
        ta = (torch.randn(512, 64, device=dev) * 0.25).to(torch.float8_e5m2)
        tb = (torch.randn(64, 64, device=dev) * 0.25).to(torch.float8_e5m2)
        pa, pb = ta.view(torch.uint8), tb.view(torch.uint8)
    else:     # a 64-byte row = 32 halves (one half of the channels of the fp16 activation plane)
        ta = torch.randn(512, 32, device=dev).half()
        tb = torch.randn(64, 32, device=dev).half()
        pa, pb = ta, tb
    out["shifted_sw64_" + kname] = {}
    for shift, sbo in ((0, 512), (8, 512), (1, 512), (3, 512), (0, 640), (1, 640), (10, 640), (11, 640), (21, 640), (22, 640)):
        rows = torch.tensor([shift + (r // 8) * (sbo // 64) + (r % 8) for r in range(128)], device=dev)
        exp = ta.double()[rows] @ tb.double().T
        best = None
        for base_offset in (0, (shift % 8)):
            d = torch.zeros((128, 64), dtype=torch.float32, device=dev)
            assert lib.mtb_exp_shifted_desc8(pa.data_ptr(), pb.data_ptr(), d.data_ptr(), shift, sbo, base_offset, kind,
                                             _lib.stream_ptr()) == 0
            torch.cuda.synchronize()
            err = float((d.double() - exp).abs().max())
            out["shifted_sw64_" + kname][f"shift{shift}_sbo{sbo}_base{base_offset}"] = err
            best = err if best is None else min(best, err)
        print(f"sw64 {kname} shift {shift} sbo {sbo}: max abs err {best:.3e} (ref max {float(exp.abs().max()):.2f})", flush=True)

names = {4: "M128N240_f16", 5: "M64N240_f16", 10: "M128N240_e5m2", 11: "M64N240_e5m2",
         12: "M128N240_f16+M64N240_e5m2_one_acc", 13: "M128N240_f16+M128N240_e5m2_one_acc",
         14: "2xM128N240_f16+1xM64N240_e5m2_sw128", 15: "2xM128N240_f16+1xM64N240_e5m2_sw64",
         16: "M128N240_f16_sw64", 17: "M64N240_e5m2_sw64"}
for pattern, nm in names.items():
    cyc = torch.zeros(148, dtype=torch.int64, device=dev)
    iters = 200
    for _ in range(2):
        sw64 = pattern in (15, 16, 17)
        assert lib.mtb_exp_mma_rate(cyc.data_ptr(), 148, pattern, iters, 640 if sw64 else 1280, 64 if sw64 else 128, 0,
                                    _lib.stream_ptr()) == 0
        torch.cuda.synchronize()
    n_mma = iters * 36 * (2 if pattern in (12, 13) else 3 if pattern in (14, 15) else 1)
    out["rate_" + nm] = {"clk_per_mma_mean": round(float(cyc.float().mean()) / n_mma, 2),
                         "clk_per_mma_max": round(float(cyc.float().max()) / n_mma, 2)}
    print(nm, out["rate_" + nm], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp8_probe.json"), "w"), indent=1)
