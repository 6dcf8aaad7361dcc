"""GPU dev probe: RT-DETR B200 vs the transformers oracle, stage by stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rtdetr_oracle as R
from mangatranslator_b200 import synth
from mangatranslator_b200.rtdetr import RtDetrB200
cfg, m = R.make_model(0)
proc = R.make_processor()
net = RtDetrB200(m.state_dict(), cfg, torch.device("cuda:0"))
for seed, (h, w) in ((3, (640, 640)), (4, (900, 620)), (2, (640, 640))):
    rgb = synth.make_page(seed, h, w, n_bubbles=5).image_rgb
    ref = R.predict(m, proc, rgb, conf=0.35)
    dbg = {}
    from mangatranslator_b200.preproc import resize_aa_device
    x = torch.from_numpy(rgb).cuda()
    x = x if (h, w) == (640, 640) else resize_aa_device(x, 640, 640)
    logits, boxes = net.forward_u8(x, debug=dbg)
    score = ref["enc_cls"].max(-1).values
    srt = torch.sort(score, descending=True).values
    ours, theirs = set(dbg["topk"].cpu().tolist()), set(torch.topk(score, 300).indices.tolist())
    print("seed", seed, "enc_cls err", (dbg["enc_cls"].cpu() - ref["enc_cls"]).abs().max().item(), "gap@300", (srt[299] - srt[300]).item(),
          "set diff", len(ours ^ theirs))
    # match queries by topk index
    oi = dbg["topk"].cpu(); ti = torch.topk(score, 300).indices
    pos = {int(t): i for i, t in enumerate(ti)}
    common = [(i, pos[int(t)]) for i, t in enumerate(oi) if int(t) in pos]
    a = torch.tensor([c[0] for c in common]); b = torch.tensor([c[1] for c in common])
    print("  logits err", (logits.cpu()[a] - ref["logits"][b]).abs().max().item(), "boxes err", (boxes.cpu()[a] - ref["pred_boxes"][b]).abs().max().item(),
          "n det", len(ref["conf"]))
# timing: eager first call vs graph replays
import time
x = torch.from_numpy(synth.make_page(7, 640, 640, n_bubbles=5).image_rgb).cuda()
l0, b0 = [t.clone() for t in net.forward_u8(x)]
for _ in range(3):
    l1, b1 = net.forward_u8(x)
torch.cuda.synchronize()
print("graph == eager:", torch.equal(l0, l1), torch.equal(b0, b1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    net.forward_u8(x)
e1.record(); torch.cuda.synchronize()
print("ms per image (graph):", e0.elapsed_time(e1) / 10)
from mangatranslator_b200 import graphs
graphs.ENABLED = False
e0.record()
for _ in range(5):
    net.forward_u8(x)
e1.record(); torch.cuda.synchronize()
print("ms per image (eager):", e0.elapsed_time(e1) / 5)
