"""One page through detect + segment (YOLO graph, NMS, SAM encoder/decoder with 12 boxes) for an ncu launch list."""
import os, sys
os.environ["MTB200_CUDA_GRAPHS"] = "0"
sys.path.insert(0, ".")
import numpy as np, torch
from mangatranslator_b200 import synth
from mangatranslator_b200.core.image.detection import detect_pages_device
pg = synth.make_page(1, 1536, 1024)
page = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).cuda()
for _ in range(2):
    detect_pages_device([page], injected_boxes=[pg.boxes_xyxy])
torch.cuda.synchronize()
print("done")
