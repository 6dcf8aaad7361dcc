"""RCAN page (1536x1024, 10x20) timed per launch kind for both body formats in one process: fp16c (fp16 + e5m2 correction,
conv_halo_fp16c.cu) and bf16x3 (conv_halo_cm.cu).  Also the accuracy of each against a float64 torch evaluation on a
crop, and (optional, MTB200_PROF_DEBUG=1) the barrier-wait counters of the fp16c kernel.  Writes
gpurun_out/rcan_precisions.json.

    python tools/prof_rcan_precisions.py [H W]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mangatranslator_b200 import weights as W  # noqa: E402
from mangatranslator_b200.rcan import RcanB200  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 2 else 1536
Wd = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
sd = W.rcan_state_dict(0)
img = torch.randint(0, 256, (H, Wd, 3), dtype=torch.uint8, device=dev)
out = {"page": [H, Wd]}
res = {}
for prec in ("fp16c", "fp16c_v0", "bf16x3"):
    os.environ["MTB200_FP16C_VARIANT"] = "0" if prec == "fp16c_v0" else "1"
    net = RcanB200(sd, dev, precision=prec.split("_")[0])
    net.upscale_u8(img)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        steps = net.time_steps(img)
        kinds = {}
        for k, ms in steps:
            kinds.setdefault(k, []).append(ms)
        body = kinds["conv_body"]
        rec = {"total_ms": round(sum(ms for _, ms in steps), 3),
               "conv1_avg_ms": round(float(np.mean(body[0::2])), 4), "conv2_avg_ms": round(float(np.mean(body[1::2])), 4),
               "by_kind_ms": {k: round(float(np.sum(v)), 3) for k, v in kinds.items()},
               "by_kind_n": {k: len(v) for k, v in kinds.items()}}
        if best is None or rec["total_ms"] < best["total_ms"]:
            best = rec
    # graph replay of the whole page, the way the pipeline runs it
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net.upscale_u8(img)
    net.upscale_u8(img)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        net.upscale_u8(img)
    e1.record()
    torch.cuda.synchronize()
    best["page_ms_graph_replay"] = round(e0.elapsed_time(e1) / 5, 3)
    flops = 2.0 * H * Wd * 64 * 64 * 9
    avg = (best["conv1_avg_ms"] + best["conv2_avg_ms"]) / 2
    best["body_conv_algorithmic_tflops"] = round(flops / (avg * 1e-3) / 1e12, 1)
    out_u8, out_f = net.upscale_u8(img, want_float=True)
    res[prec] = out_f.clone()
    out[prec] = best
    print(prec, json.dumps(best), flush=True)
    del net
    torch.cuda.empty_cache()
out["fp16c_vs_bf16x3_max_abs"] = float((res["fp16c"] - res["bf16x3"]).abs().max())
out["fp16c_vs_bf16x3_mean_abs"] = float((res["fp16c"] - res["bf16x3"]).abs().mean())
out["output_range"] = [float(res["bf16x3"].min()), float(res["bf16x3"].max())]
print("fp16c vs bf16x3:", out["fp16c_vs_bf16x3_max_abs"], out["fp16c_vs_bf16x3_mean_abs"], out["output_range"], flush=True)

if os.environ.get("MTB200_PROF_DEBUG"):
    # barrier-wait counters of the fp16c kernel (DBG instantiation): one conv1 and one conv2 launch
    os.environ["MTB200_FP16C_VARIANT"] = "1"
    net = RcanB200(W.rcan_state_dict(0, n_resgroups=1, n_resblocks=2), dev, precision="fp16c")
    os.environ["MTB200_CUDA_GRAPHS"] = "0"
    net.upscale_u8(img)
    torch.cuda.synchronize()
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    os.environ["MTB200_HALO_DEBUG"] = "32"
    os.environ["MTB200_HALO_DEBUG_PTR"] = str(dbg.data_ptr())
    b = net._get(H, Wd)
    names = []
    for kind, arg in b["steps"]:
        if kind != "conv_body":
            continue
        dbg.zero_()
        arg.run()
        torch.cuda.synchronize()
        d = dbg.view(148, 16).double()
        names.append({"mma_wait_full": float(d[:, 4].mean()), "mma_wait_tempty": float(d[:, 5].mean()),
                      "tiles": float(d[:, 6].mean()), "mma_span": float(d[:, 7].mean()),
                      "epi_w2_wait_tfull": float(d[:, 8].mean()), "epi_w2_work": float(d[:, 9].mean()),
                      "epi_w17_wait_tfull": float(d[:, 10].mean()), "epi_w17_work": float(d[:, 11].mean())})
    del os.environ["MTB200_HALO_DEBUG"]
    out["fp16c_debug_counters_cycles_per_cta"] = names
    print(json.dumps(names, indent=1), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rcan_precisions.json"), "w"), indent=1)
