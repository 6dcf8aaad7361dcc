"""Runs the non-DNN kernels of the path once each on page-sized inputs (for an `ncu --set full` capture):
Pillow-exact LANCZOS resample, SAM antialias resize, conjoined split, RT-DETR deformable attention / stem pool, bubble cleaning."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from mangatranslator_b200 import conjoined as Cj, synth, weights as W
from mangatranslator_b200.core.image.cleaning import clean_pages_device
from mangatranslator_b200.preproc import resize_aa_device, resize_lanczos_device
from mangatranslator_b200.rtdetr import RtDetrB200
dev = torch.device("cuda:0")
H, Wd = 1536, 1024
pg = synth.make_page(1, H, Wd, n_bubbles=12)
page = torch.from_numpy(np.ascontiguousarray(pg.image_rgb[:, :, ::-1])).to(dev)
up = torch.randint(0, 256, (2 * H, 2 * Wd, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    resize_lanczos_device(up, int(1.5 * H), int(1.5 * Wd))
    resize_aa_device(page, 1024, 1024)
mask = torch.zeros((H, Wd), dtype=torch.uint8, device=dev)
mask[300:760, 150:860] = 255
boxes = torch.tensor([[150.0, 300.0, 520.0, 700.0], [440.0, 320.0, 820.0, 720.0]])
for _ in range(2):
    Cj.split_conjoined_device(mask, boxes, include_child_rects=False)
dets = synth.detections_from_page(pg)
dd = [dict(d, sam_mask=torch.from_numpy(d["sam_mask"]).to(dev)) for d in dets]
for _ in range(2):
    clean_pages_device([page], [dd], processing_scale=(H * Wd / 1e6) ** 0.5)
cfg, sd = W.rtdetr_model_and_state(0)
net = RtDetrB200(sd, cfg, dev)
from mangatranslator_b200 import graphs
graphs.ENABLED = False
x = resize_aa_device(page[:, :, [2, 1, 0]].contiguous(), 640, 640)
net.forward_u8(x)
net.forward_u8(x)
torch.cuda.synchronize()
print("done")
