"""Halo-kernel bottleneck experiments: time the RCAN body layer with parts of the kernel disabled."""
import os, sys, json
sys.path.insert(0, ".")
import torch
from mangatranslator_b200 import planes as P
from mangatranslator_b200.ops import ConvPlan
dev = torch.device("cuda:0")
H, W = 1536, 1024
torch.manual_seed(0)
x = torch.randn(1, 64, H, W, device=dev)
wt = torch.randn(64, 64, 3, 3, device=dev) / 24
xp, wp = P.nchw_to_planes(x, 2), P.conv_weight_to_planes(wt, 2)
o = torch.zeros(2, 1, H, W, 64, dtype=torch.bfloat16, device=dev)
plan = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", mode=2)
def timeit(iters=10):
    for _ in range(3): plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): plan.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {}
for dbg, name in [(0, "full"), (1, "no_stores"), (2, "one_tap_mma"), (4, "no_tma"), (3, "no_stores+one_tap"), (5, "no_stores+no_tma"), (6, "one_tap+no_tma"), (7, "none")]:
    os.environ["MTB200_HALO_DEBUG"] = str(dbg)
    out[name] = timeit()
    print(name, round(out[name], 4), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dev4.json", "w"), indent=1)
