"""A/B sweep of the fp16c body conv's launch knobs inside ONE process, configurations interleaved so that they see the
same thermal / power state: MTB200_FP16C_EPI (2 = exchange epilogue, 1 = pair epilogue, 0 = per layer kind),
MTB200_FP16C_PF (L2 prefetch distance of the activation tile), MTB200_FP16C_RPF (TMA L2 prefetch of the residual tile).
Prints conv1 / conv2 average ms on a 1536x1024 frame and writes gpurun_out/fp16c_sweep.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mangatranslator_b200 import weights as W  # noqa: E402
from mangatranslator_b200.rcan import RcanB200  # noqa: E402

dev = torch.device("cuda:0")
net = RcanB200(W.rcan_state_dict(0, n_resgroups=int(os.environ.get("SWEEP_GROUPS", "3"))), dev, precision="fp16c")
img = torch.randint(0, 256, (1536, 1024, 3), dtype=torch.uint8, device=dev)
net.time_steps(img)
configs = [dict(MTB200_FP16C_STREAM=st, MTB200_FP16C_PF=pf) for st in (1, 0) for pf in (0, 2)]
acc = {i: [] for i in range(len(configs))}
for rep in range(int(os.environ.get("SWEEP_REPS", "6"))):
    for i, cfg in enumerate(configs):
        for k, v in cfg.items():
            os.environ[k] = str(v)
        body = [ms for k, ms in net.time_steps(img) if k == "conv_body"]
        acc[i].append((float(np.mean(body[0::2])), float(np.mean(body[1::2]))))
out = []
for i, cfg in enumerate(configs):
    a = np.array(acc[i])
    rec = dict(cfg, conv1_ms=round(float(np.median(a[:, 0])), 4), conv2_ms=round(float(np.median(a[:, 1])), 4),
               sum_ms=round(float(np.median(a.sum(1))), 4), sum_min=round(float(a.sum(1).min()), 4))
    out.append(rec)
    print(rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp16c_sweep.json"), "w"), indent=1)
