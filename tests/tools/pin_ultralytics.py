"""Pin the restated ultralytics block rules against the real library — for a machine that HAS `ultralytics` and a
checkpoint (neither exists in the build container, which is why oracle/yolo_oracle.py and oracle/yolo_tree_oracle.py say
"parity unpinned").

    python tests/tools/pin_ultralytics.py path/to/model.pt [image.jpg] [--imgsz 640] [--conf 0.25] [--cpu-only]

What it does: runs `ultralytics.YOLO(path)` (the reference's call, core/ml/model_manager.py:740,804,830) and (a) the CPU
oracle on the module tree read from the same file by `weights.load_ultralytics_tree` (no GPU needed), (b) unless
--cpu-only, the CUDA executor `YoloTreeB200`; prints the largest differences of the raw head tensors, boxes and scores,
and whether the NMS picks the same anchors.  Exit code 0 when heads agree to 1e-3 relative and the detections match.
Any disagreement of (a) localises the block rule that was remembered wrongly (the per-layer outputs are printed with
--layers)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint")
    ap.add_argument("image", nargs="?")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--conf", type=float, default=0.25)
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--layers", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    try:
        from ultralytics import YOLO
    except ImportError:
        print("ultralytics is not installed here: nothing to pin against")
        return 2
    import yolo_oracle as Y
    import yolo_tree_oracle as O
    from mangatranslator_b200 import weights as W
    if a.image:
        import cv2
        img = cv2.imread(a.image)
    else:
        from mangatranslator_b200 import synth
        img = np.ascontiguousarray(synth.make_page(0, 1150, 800).image_rgb[:, :, ::-1])
    ref_model = YOLO(a.checkpoint)
    res = ref_model(img, conf=a.conf, imgsz=a.imgsz, verbose=False, device="cpu")[0]
    rb, rc = res.boxes.xyxy.cpu(), res.boxes.conf.cpu()
    tree = W.load_ultralytics_tree(a.checkpoint)
    x = Y.preprocess(img, a.imgsz)
    # raw head tensors of the library model on the same input (un-fused weights: eval mode, BatchNorm running stats)
    net = ref_model.model.float().eval()
    feats = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, m in enumerate(net.model)]
    with torch.no_grad():
        net(x)
    for h in hooks:
        h.remove()
    ours = O.predict(tree, img, a.conf, a.imgsz)
    if a.layers:
        outs, cur = [], x
        import torch.nn.functional as F
        for k, node in enumerate(tree["layers"][:-1]):
            f = node.get("f", -1)
            srcs = [cur if j == -1 else outs[j] for j in (f if isinstance(f, (list, tuple)) else [f])]
            t = node["t"]
            cur = torch.cat(srcs, 1) if t == "Concat" else F.interpolate(srcs[0], scale_factor=2, mode="nearest") \
                if t == "Upsample" else O.block(node, srcs[0])
            outs.append(cur)
            want = feats[k]
            print(f"layer {k:2d} {t:9s} max |diff| {float((cur - want).abs().max()):.3e}  (|ref|max {float(want.abs().max()):.3f})")
    ok = len(ours["conf"]) == len(rc)
    print(f"oracle vs ultralytics: {len(ours['conf'])} vs {len(rc)} detections")
    if ok and len(rc):
        print(f"  boxes max |diff| {float((ours['xyxy'] - rb).abs().max()):.4f} px, scores {float((ours['conf'] - rc).abs().max()):.2e}")
        ok = float((ours["xyxy"] - rb).abs().max()) < 0.05 and float((ours["conf"] - rc).abs().max()) < 1e-3
    if not a.cpu_only and torch.cuda.is_available():
        from mangatranslator_b200.yolo_tree import YoloTreeB200
        dev_res = YoloTreeB200(tree, torch.device("cuda:0"))(img, conf=a.conf, imgsz=a.imgsz)[0]
        n = 0 if dev_res.boxes is None else len(dev_res.boxes)
        print(f"CUDA executor: {n} detections")
        if n == len(rc) and n:
            print(f"  boxes max |diff| {float((dev_res.boxes.xyxy.cpu() - rb).abs().max()):.4f} px")
        ok = ok and n == len(rc)
    print("PINNED" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
