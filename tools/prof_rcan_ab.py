"""Two RCAN pages under ncu: first with MTB200_HALO_DEBUG=64 (residual L2 prefetch off), then with it on (A/B in one process)."""
import os, sys
os.environ["MTB200_CUDA_GRAPHS"] = "0"
sys.path.insert(0, ".")
import torch
from mangatranslator_b200 import weights as W
from mangatranslator_b200.rcan import RcanB200
dev = torch.device("cuda:0")
net = RcanB200(W.rcan_state_dict(0, n_resgroups=1, n_resblocks=6), dev)
img = torch.randint(0, 256, (1536, 1024, 3), dtype=torch.uint8, device=dev)
net.upscale_u8(img)
torch.cuda.synchronize()
for flag in ("64", "0", "64", "0"):
    os.environ["MTB200_HALO_DEBUG"] = flag
    net.upscale_u8(img)
    torch.cuda.synchronize()
print("done")
