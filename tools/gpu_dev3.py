"""GPU dev check #3: layer timings after the epilogue rewrite + RCAN page time."""
import sys, json, os, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
from mangatranslator_b200 import planes as P
from mangatranslator_b200.ops import ConvPlan
dev = torch.device("cuda:0")
out = {}
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
H, W = 1536, 1024
torch.manual_seed(0)
x = torch.randn(1, 64, H, W, device=dev)
wt = torch.randn(64, 64, 3, 3, device=dev) / 24
for planes in (2, 1):
    xp, wp = P.nchw_to_planes(x, planes), P.conv_weight_to_planes(wt, planes)
    o = torch.zeros(planes, 1, H, W, 64, dtype=torch.bfloat16, device=dev)
    for mode, name in ((2, "halo"), (3, "halo_pixel_major"), (1, "per_tap")):
        plan = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", mode=mode)
        ms = timeit(plan.run)
        out[f"layer_{name}_planes{planes}"] = dict(ms=ms, tflops_alg=2.0 * H * W * 64 * 64 * 9 / ms / 1e9)
        print(name, planes, out[f"layer_{name}_planes{planes}"], flush=True)
del xp, wp, o, x
torch.cuda.empty_cache()
import rcan_oracle
from mangatranslator_b200.rcan import RcanB200
m = rcan_oracle.make_model(0)
for prec in ("bf16x3", "bf16"):
    net = RcanB200(m.state_dict(), dev, precision=prec)
    img = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev)
    net.upscale_u8(img); torch.cuda.synchronize()
    ms = timeit(lambda: net.upscale_u8(img), iters=3, warm=1)
    out[f"rcan_page_{prec}"] = dict(ms=ms, pages_per_s=1000.0 / ms, tflops_alg=48.6e3 / ms)
    print(prec, out[f"rcan_page_{prec}"], flush=True)
    del net
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dev3.json", "w"), indent=1)
