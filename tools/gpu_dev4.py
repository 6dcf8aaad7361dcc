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
plan = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", mode=int(os.environ.get("DEV4_MODE", "2")))
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

# stall accounting (debug bit 16: producer lane; bit 32: MMA warp and two epilogue warps), cycles per CTA
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
os.environ["MTB200_HALO_DEBUG_PTR"] = str(dbg.data_ptr())
for flags in ((16, 32, 32 | 1, 32 | 4) if os.environ.get("DEV4_MODE", "2") == "2" else (32, 32 | 1, 32 | 4, 32 | 5)):
    os.environ["MTB200_HALO_DEBUG"] = str(flags)
    dbg.zero_()
    plan.run(); torch.cuda.synchronize()
    d = dbg.view(148, 16).double().mean(0)
    if flags & 16:
        print("debug", flags, "tma_clk_avg", float(d[0] / d[2]), "empty_wait_clk_avg", float(d[1] / d[2]), flush=True)
    else:
        print("debug", flags, "per-CTA clk: total", int(d[7]), "mma wait full", int(d[4]), "mma wait tempty", int(d[5]), "tiles", float(d[6]),
              "| epi warp2 wait tfull", int(d[8]), "work", int(d[9]), "| epi warp17 wait", int(d[10]), "work", int(d[11]), flush=True)
os.environ["MTB200_HALO_DEBUG"] = "0"
