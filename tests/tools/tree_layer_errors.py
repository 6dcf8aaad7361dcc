"""GPU debug: per-layer error of the module-tree detector against its CPU oracle (where does the head error come from?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MTB200_SYNTHETIC_WEIGHTS", "1")
import numpy as np, torch, torch.nn.functional as F
import yolo_oracle as Y, yolo_tree_oracle as O
from test_yolo_tree_gpu import _image, _tree, FAMILIES
from mangatranslator_b200.preproc import letterbox_device
from mangatranslator_b200.yolo_tree import YoloTreeB200
dev = torch.device("cuda:0")
for family, kw in FAMILIES:
    img = _image(3, 300, 420)
    tree = _tree(family, kw, 3, img, 448)
    x = Y.preprocess(img, 448)
    # oracle per-layer outputs
    outs, cur = [], x
    for node in tree["layers"][:-1]:
        f = node.get("f", -1)
        srcs = [cur if j == -1 else outs[j] for j in (f if isinstance(f, (list, tuple)) else [f])]
        t = node["t"]
        cur = torch.cat(srcs, 1) if t == "Concat" else F.interpolate(srcs[0], scale_factor=2, mode="nearest") if t == "Upsample" else O.block(node, srcs[0])
        outs.append(cur)
    net = YoloTreeB200(tree, dev)
    lb = letterbox_device(torch.from_numpy(img).to(dev), 448, swap_rb=True)
    g = net.forward_letterboxed(lb)
    torch.cuda.synchronize()
    for i, (ref, (sl, hh, ww)) in enumerate(zip(outs, g["layer_outputs"])):
        got = sl.buf.float().sum(0)[0, :, :, sl.off:sl.off + sl.c].permute(2, 0, 1).cpu()
        got = got[:ref.shape[1]]
        err = (got - ref[0]).abs().max().item()
        print(f"yolo{family} layer {i:2d} {tree['layers'][i]['t']:9s} |ref|max {float(ref.abs().max()):8.3f} err {err:.2e} rel {err / float(ref.abs().max()):.1e}")
