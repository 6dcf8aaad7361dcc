"""One RCAN body layer (1536x1024x64, bf16x3, halo kernel) for an ncu capture."""
import sys
sys.path.insert(0, ".")
import torch
from mangatranslator_b200 import planes as P
from mangatranslator_b200.ops import ConvPlan
dev = torch.device("cuda:0")
H, W = 1536, 1024
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
x = torch.randn(1, 64, H, W, device=dev)
wt = torch.randn(64, 64, 3, 3, device=dev) / 24
xp, wp = P.nchw_to_planes(x, planes), P.conv_weight_to_planes(wt, planes)
o = torch.zeros(planes, 1, H, W, 64, dtype=torch.bfloat16, device=dev)
plan = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", mode=mode)
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
