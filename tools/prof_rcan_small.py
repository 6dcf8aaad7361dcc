"""A short RCAN (1 group x 3 blocks) on a 1536x1024 frame, eager launches: the target of `ncu -k regex:conv3x3_c64`."""
import os, sys
os.environ["MTB200_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mangatranslator_b200 import weights as W
from mangatranslator_b200.rcan import RcanB200
dev = torch.device("cuda:0")
net = RcanB200(W.rcan_state_dict(0, n_resgroups=1, n_resblocks=3), dev, precision=os.environ.get("PREC", "fp16c"))
img = torch.randint(0, 256, (1536, 1024, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
    net.upscale_u8(img)
torch.cuda.synchronize()
print("done")
