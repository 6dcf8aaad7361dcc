"""Weight sources for the hot-path models.

Real checkpoints are read from ``$MT_MODELS_DIR`` when present (same file names the reference's ModelManager uses,
core/ml/model_manager.py:108-119,183-204).  There is no network in the build / GPU containers and no checkpoint on
disk, so otherwise the state dicts below are generated: correct key names and shapes for the architectures, seeded
variance-preserving random values (so activations stay O(1) through the ~60-400 layer stacks and parity tests exercise
every layer), with the YOLO / SAM heads calibrated to produce a realistic number of detections / non-trivial masks.
The SAME tensors feed the CPU oracle (tests, bench cpu_baseline) and the CUDA path.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch


def models_dir() -> Optional[str]:
    d = os.environ.get("MT_MODELS_DIR")
    return d if d and os.path.isdir(d) else None


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def _conv(g, cout, cin, k, gain=1.0, bias_std=0.02):
    fan_in = cin * k * k
    w = torch.randn((cout, cin, k, k), generator=g) * (gain / math.sqrt(fan_in))
    b = torch.randn((cout,), generator=g) * bias_std
    return w, b


# ---- RCAN (2x-AnimeSharpV4 architecture family) -----------------------------------------------------------------
def rcan_state_dict(seed: int = 0, n_resgroups: int = 10, n_resblocks: int = 20, n_feats: int = 64,
                    reduction: int = 16, unshuffle: int = 1) -> Dict[str, torch.Tensor]:
    """`unshuffle` = 2 gives the "_PU" layout (12 input channels, two upsampler stages: tail.0.0 and tail.0.2)."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}

    def put(name, w, b):
        sd[name + ".weight"], sd[name + ".bias"] = w, b

    f = n_feats
    put("head.0", *_conv(g, f, 3 * unshuffle * unshuffle, 3, gain=1.4))
    for gi in range(n_resgroups):
        for bi in range(n_resblocks):
            base = f"body.{gi}.body.{bi}.body"
            put(base + ".0", *_conv(g, f, f, 3, gain=1.4))          # conv + ReLU
            put(base + ".2", *_conv(g, f, f, 3, gain=0.25))         # residual branch kept small, like trained nets: the
                                                                    # trunk stays O(1) over the 200 blocks (with 0.5 it grew
                                                                    # x46 and the output left [0,1] by an order of magnitude)
            put(base + ".3.conv_du.0", *_conv(g, f // reduction, f, 1, gain=1.0))
            put(base + ".3.conv_du.2", *_conv(g, f, f // reduction, 1, gain=1.0))
        put(f"body.{gi}.body.{n_resblocks}", *_conv(g, f, f, 3, gain=0.25))
    put(f"body.{n_resgroups}", *_conv(g, f, f, 3, gain=0.3))
    put("tail.0.0", *_conv(g, 4 * f, f, 3, gain=1.0))
    up, i = 2 * unshuffle, 2
    while up > 2:
        put(f"tail.0.{i}", *_conv(g, 4 * f, f, 3, gain=1.0))
        up, i = up // 2, i + 2
    put("tail.1", *_conv(g, 3, f, 3, gain=0.2))
    sd["tail.1.bias"] = torch.full((3,), 0.5)                        # mid-grey output so pixels are not clipped
    return sd


# ---- YOLOv8-seg ---------------------------------------------------------------------------------------------------
def yolo_cfg(variant: str = "m", nc: int = 1) -> dict:
    depth, width, max_ch = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
                            "l": (1.0, 1.0, 512), "x": (1.0, 1.25, 512)}[variant]
    return dict(nc=nc, depth=depth, width=width, max_ch=max_ch, nm=32, npr=256)


def yolo_state_dict(seed: int = 0, cfg: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Keys follow the layer naming of mangatranslator_b200.yolo / oracle.yolo_oracle (l0..l21, head.*)."""
    cfg = cfg or yolo_cfg("m")
    g = _gen(seed)
    depth, width, max_ch = cfg["depth"], cfg["width"], cfg["max_ch"]
    d = lambda n: max(round(n * depth), 1)
    c = lambda x: int(math.ceil(min(x, max_ch) * width / 8) * 8)
    sd: Dict[str, torch.Tensor] = {}

    # Gains: 1.6 keeps the second moment through conv -> SiLU chains; the bottlenecks' residual branches (0.8) and the
    # convs that read a concat (1.4) are damped so that the 8 C2f stages of the "m" layout keep activations at O(1-10)
    # like a trained, BN-folded detector (with 1.6 everywhere they reached 1e4 and the prototypes 1e4: any absolute
    # tolerance on head tensors was meaningless).
    def conv(name, cin, cout, k, gain=1.6):
        w, b = _conv(g, cout, cin, k, gain=gain, bias_std=0.05)
        sd[name + ".weight"], sd[name + ".bias"] = w, b

    def c2f(name, c1, c2, n):
        hc = c2 // 2
        conv(f"{name}.cv1.conv", c1, 2 * hc, 1)
        conv(f"{name}.cv2.conv", (2 + n) * hc, c2, 1, gain=1.4)
        for i in range(n):
            conv(f"{name}.m.{i}.cv1.conv", hc, hc, 3)
            conv(f"{name}.m.{i}.cv2.conv", hc, hc, 3, gain=0.8)

    c64, c128, c256, c512, c1024 = c(64), c(128), c(256), c(512), c(1024)
    conv("l0.conv", 3, c64, 3)
    conv("l1.conv", c64, c128, 3)
    c2f("l2", c128, c128, d(3))
    conv("l3.conv", c128, c256, 3)
    c2f("l4", c256, c256, d(6))
    conv("l5.conv", c256, c512, 3)
    c2f("l6", c512, c512, d(6))
    conv("l7.conv", c512, c1024, 3)
    c2f("l8", c1024, c1024, d(3))
    conv("l9.cv1.conv", c1024, c1024 // 2, 1)
    conv("l9.cv2.conv", c1024 // 2 * 4, c1024, 1, gain=1.4)
    c2f("l12", c1024 + c512, c512, d(3))
    c2f("l15", c512 + c256, c256, d(3))
    conv("l16.conv", c256, c256, 3)
    c2f("l18", c256 + c512, c512, d(3))
    conv("l19.conv", c512, c512, 3)
    c2f("l21", c512 + c1024, c1024, d(3))
    nc, nm = cfg["nc"], cfg["nm"]
    ch = (c256, c512, c1024)
    c2h, c3h, c4h = max(16, ch[0] // 4, 64), max(ch[0], min(nc, 100)), max(ch[0] // 4, nm)
    for i, x in enumerate(ch):
        for br, hcx, oc in (("cv2", c2h, 64), ("cv3", c3h, nc), ("cv4", c4h, nm)):
            conv(f"head.{br}.{i}.0.conv", x, hcx, 3)
            conv(f"head.{br}.{i}.1.conv", hcx, hcx, 3)
            conv(f"head.{br}.{i}.2", hcx, oc, 1, gain=0.2 if br == "cv2" else 0.05)
        # class head: logits of about -4 +- 1.5, so that (like a trained detector on a clean page) a handful of anchors,
        # not thousands, clear conf=0.6 and the scores differ from anchor to anchor by far more than the numerical noise;
        # stage-level runs inject the page's ground-truth boxes anyway
        sd[f"head.cv3.{i}.2.weight"] *= 5.0
        # the stride-16 level of the seeded "m" network sits ~1.4 above the others on the synthetic pages: with the extra
        # shift about twenty of the 35 700 anchors of a 1600-pixel page clear conf = 0.6 (a page's worth of bubbles)
        sd[f"head.cv3.{i}.2.bias"] = torch.full((nc,), -5.4 if i == 1 else -4.0)
    npr = c(cfg["npr"])
    conv("head.proto.cv1.conv", c256, npr, 3)
    w = torch.randn((npr, npr, 2, 2), generator=g) * (1.6 / math.sqrt(npr))
    sd["head.proto.upsample.weight"], sd["head.proto.upsample.bias"] = w, torch.randn((npr,), generator=g) * 0.1
    conv("head.proto.cv2.conv", npr, npr, 3)
    conv("head.proto.cv3.conv", npr, nm, 1)
    return sd


# ---- ultralytics checkpoints (the reference's detector files: core/ml/model_manager.py:183-190, 711-743) -----------------
class UnsupportedCheckpoint(ValueError):
    pass


def synthetic_allowed() -> bool:
    """Seeded synthetic weights are an explicit opt-in (tests, bench, smoke): a missing checkpoint is an error otherwise."""
    return os.environ.get("MTB200_SYNTHETIC_WEIGHTS", "0") == "1"


def _load_pickle_inert(path: str):
    """torch.load of a file that pickles objects of packages that are not installed (ultralytics): torch / numpy / builtin
    classes resolve normally, every other class becomes an inert attribute bag.  Nothing from the file is executed."""
    import pickle

    class _Bag:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            if isinstance(state, dict):
                self.__dict__.update(state)
            elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
                self.__dict__.update(state[0] or {})
                self.__dict__.update(state[1])

        def __call__(self, *a, **k):
            return self

    _safe_roots = ("torch", "collections", "numpy", "builtins", "_codecs", "copyreg", "pathlib", "__builtin__")

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in _safe_roots and not (module == "builtins" and name in ("eval", "exec", "compile", "open",
                                                                                          "__import__", "getattr", "setattr")):
                try:
                    return super().find_class(module, name)
                except Exception:
                    pass
            return type(name, (_Bag,), {"__module__": module})

    class _PickleModule:
        Unpickler = _Unpickler
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
        __name__ = "pickle"

    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)


def load_ultralytics_state_dict(path: str):
    """State dict (+ class names) of an ultralytics `.pt` without ultralytics installed.

    `YOLO(path)` unpickles `ckpt["model"]`, an `ultralytics.nn.tasks.SegmentationModel` object.  Here the pickle is read
    with an Unpickler that resolves torch / numpy / builtin classes normally and replaces every `ultralytics.*` (and any
    other unknown) class by an inert attribute bag, then the module tree (`_modules` / `_parameters` / `_buffers`) is
    walked into the flat `model.N....` names a state dict has.  Nothing from the file is executed.  Plain state-dict
    files (a dict of tensors, or {"model": state_dict}) are accepted as they are."""
    obj = _load_pickle_inert(path)
    names = None
    root = obj
    if isinstance(obj, dict):
        root = obj.get("ema") or obj.get("model") or obj
        if isinstance(root, dict) and all(torch.is_tensor(v) for v in root.values()):
            return {k: v.float() for k, v in root.items()}, obj.get("names")
        if isinstance(obj, dict) and all(torch.is_tensor(v) for v in obj.values()):
            return {k: v.float() for k, v in obj.items()}, None
    names = getattr(root, "names", None)
    sd: Dict[str, torch.Tensor] = {}

    def walk(mod, prefix):
        d = getattr(mod, "__dict__", {})
        for store in ("_parameters", "_buffers"):
            for k, v in (d.get(store) or {}).items():
                if v is not None and torch.is_tensor(v):
                    sd[prefix + k] = v.detach().float()
        for k, sub in (d.get("_modules") or {}).items():
            if sub is not None:
                walk(sub, prefix + k + ".")

    walk(root, "")
    if not sd:
        raise UnsupportedCheckpoint(f"{path}: no tensors found in the checkpoint")
    return sd, (dict(names) if isinstance(names, dict) else names)


def yolo_from_ultralytics(sd: Dict[str, torch.Tensor], names=None):
    """ultralytics YOLOv8-seg state dict (`model.N.conv.weight`, `model.N.bn.*`, `model.22.cv2/cv3/cv4/proto`) ->
    (state dict in this package's layout with BatchNorm folded into the convs, cfg).  Raises UnsupportedCheckpoint for
    other families (YOLO11's C3k2 / C2PSA blocks, detection-only heads, ...).  What ultralytics itself does at first
    predict (`model.fuse()`): w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps), eps = 1e-3."""
    if any(k.startswith("model.model.") for k in sd):
        sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    keys = set(sd)
    if "model.0.conv.weight" not in keys:
        raise UnsupportedCheckpoint("not an ultralytics detection checkpoint (no model.0.conv.weight)")
    if any(".attn." in k or ".m.0.m." in k or k.startswith("model.23.") for k in keys):
        raise UnsupportedCheckpoint("YOLO11 / YOLO12 layout (C3k2 / C2PSA / A2C2f blocks): only YOLOv8-seg checkpoints are "
                                    "supported by the speech-bubble detector of this build")
    if "model.22.proto.cv1.conv.weight" not in keys or "model.22.cv4.0.2.weight" not in keys:
        raise UnsupportedCheckpoint("not a YOLOv8 *segmentation* checkpoint (no model.22.proto / cv4 mask branch)")
    eps = 1e-3

    def fold(src: str):
        w = sd[src + ".conv.weight"].float()
        if src + ".bn.weight" in keys:
            g, b = sd[src + ".bn.weight"].float(), sd[src + ".bn.bias"].float()
            m, v = sd[src + ".bn.running_mean"].float(), sd[src + ".bn.running_var"].float()
            k = g / torch.sqrt(v + eps)
            return w * k.view(-1, 1, 1, 1), b - m * k
        bias = sd.get(src + ".conv.bias")                    # an already fused export
        return w, (bias.float() if bias is not None else torch.zeros(w.shape[0]))

    out: Dict[str, torch.Tensor] = {}

    def put(dst: str, src: str):
        out[dst + ".conv.weight"], out[dst + ".conv.bias"] = fold(src)

    def c2f(i: int):
        put(f"l{i}.cv1", f"model.{i}.cv1")
        put(f"l{i}.cv2", f"model.{i}.cv2")
        j = 0
        while f"model.{i}.m.{j}.cv1.conv.weight" in keys:
            put(f"l{i}.m.{j}.cv1", f"model.{i}.m.{j}.cv1")
            put(f"l{i}.m.{j}.cv2", f"model.{i}.m.{j}.cv2")
            j += 1
        return j

    for i in (0, 1, 3, 5, 7, 16, 19):
        put(f"l{i}", f"model.{i}")
    nb = {i: c2f(i) for i in (2, 4, 6, 8, 12, 15, 18, 21)}
    put("l9.cv1", "model.9.cv1")
    put("l9.cv2", "model.9.cv2")
    for br in ("cv2", "cv3", "cv4"):
        for lvl in range(3):
            for k in (0, 1):
                put(f"head.{br}.{lvl}.{k}", f"model.22.{br}.{lvl}.{k}")
            out[f"head.{br}.{lvl}.2.weight"] = sd[f"model.22.{br}.{lvl}.2.weight"].float()
            out[f"head.{br}.{lvl}.2.bias"] = sd[f"model.22.{br}.{lvl}.2.bias"].float()
    for k in ("cv1", "cv2", "cv3"):
        put(f"head.proto.{k}", f"model.22.proto.{k}")
    out["head.proto.upsample.weight"] = sd["model.22.proto.upsample.weight"].float()
    out["head.proto.upsample.bias"] = sd["model.22.proto.upsample.bias"].float()
    # scale: width from the stem, depth from the bottleneck counts, as ultralytics' yaml scales define them
    c0 = out["l0.conv.weight"].shape[0]
    table = {16: "n", 32: "s", 48: "m", 64: "l", 80: "x"}
    if c0 not in table:
        raise UnsupportedCheckpoint(f"unknown YOLOv8 width (stem has {c0} channels)")
    cfg = yolo_cfg(table[c0], nc=int(out["head.cv3.0.2.weight"].shape[0]))
    cfg["nm"] = int(out["head.cv4.0.2.weight"].shape[0])
    # npr is stored un-scaled in the cfg (YoloB200 / the oracle apply the width multiple): invert c(x) = ceil(x*w/8)*8
    npr_scaled = int(out["head.proto.cv1.conv.weight"].shape[0])
    cfg["npr"] = next((n for n in (256, 128, 64, 32, 512) if int(math.ceil(min(n, cfg["max_ch"]) * cfg["width"] / 8) * 8) == npr_scaled),
                      None)
    if cfg["npr"] is None:
        raise UnsupportedCheckpoint(f"unexpected prototype width {npr_scaled}")
    d = lambda n: max(round(n * cfg["depth"]), 1)
    expect = {2: d(3), 4: d(6), 6: d(6), 8: d(3), 12: d(3), 15: d(3), 18: d(3), 21: d(3)}
    if nb != expect:
        raise UnsupportedCheckpoint(f"bottleneck counts {nb} do not match YOLOv8{table[c0]}-seg {expect}")
    if int(out["head.cv2.0.2.weight"].shape[0]) != 64:
        raise UnsupportedCheckpoint("DFL with reg_max != 16 is not supported")
    return out, cfg


# ---- ultralytics model OBJECT -> node tree (panel YOLO11 / OSB-text YOLO12 detectors; mangatranslator_b200/yolo_tree.py) -----
def _mods(obj) -> dict:
    return getattr(obj, "__dict__", {}).get("_modules") or {}


def _cls(obj) -> str:
    return type(obj).__name__


def _attr(obj, name, default=None):
    d = getattr(obj, "__dict__", {})
    if name in d:
        return d[name]
    for store in ("_parameters", "_buffers", "_modules"):
        if name in (d.get(store) or {}):
            return d[store][name]
    return default


def _child(obj, name, where="module"):
    m = _mods(obj).get(name)
    if m is None:
        raise UnsupportedCheckpoint(f"{_cls(obj)} has no `{name}` {where}")
    return m


def _first(v) -> int:
    return int(v[0] if isinstance(v, (tuple, list)) else v)


def _tree_conv(obj) -> dict:
    """ultralytics Conv / DWConv (conv + bn + act) or a bare torch Conv2d -> Conv node with BatchNorm folded."""
    name = _cls(obj)
    if name == "Conv2d":
        conv, bn, act = obj, None, False
    elif name in ("Conv", "DWConv"):
        conv, bn = _child(obj, "conv"), _mods(obj).get("bn")
        a = _mods(obj).get("act")
        a = a if a is not None else _attr(obj, "act")
        an = _cls(a) if a is not None else "Identity"
        if an not in ("SiLU", "Identity"):
            raise UnsupportedCheckpoint(f"activation {an} (SiLU or none supported)")
        act = an == "SiLU"
    else:
        raise UnsupportedCheckpoint(f"expected a convolution, found {name}")
    w = _attr(conv, "weight")
    if w is None or w.dim() != 4:
        raise UnsupportedCheckpoint("convolution without a 4-d weight")
    w = w.detach().float()
    k, s, p, g = _attr(conv, "kernel_size"), _attr(conv, "stride"), _attr(conv, "padding"), int(_attr(conv, "groups", 1))
    if int(w.shape[2]) != int(w.shape[3]) or _first(_attr(conv, "dilation", 1)) != 1 or isinstance(p, str):
        raise UnsupportedCheckpoint("non-square, dilated or string-padded convolution")
    cb = _attr(conv, "bias")
    b = cb.detach().float() if cb is not None else torch.zeros(w.shape[0])
    if bn is not None and _cls(bn) != "Identity":
        gam, bet = _attr(bn, "weight").detach().float(), _attr(bn, "bias").detach().float()
        mean, var = _attr(bn, "running_mean").detach().float(), _attr(bn, "running_var").detach().float()
        kk = gam / torch.sqrt(var + float(_attr(bn, "eps", 1e-3)))
        w, b = w * kk.view(-1, 1, 1, 1), bet + (b - mean) * kk
    return {"t": "Conv", "w": w.contiguous(), "b": b.contiguous(), "k": int(w.shape[2]), "s": _first(s), "p": _first(p),
            "g": g, "act": bool(act)}


def _flat_convs(obj) -> list:
    if _cls(obj) in ("Sequential", "ModuleList"):
        return [c for m in _mods(obj).values() for c in _flat_convs(m)]
    return [_tree_conv(obj)]


def _tree_block(obj) -> dict:
    name = _cls(obj)
    kids = lambda n: list(_mods(_child(obj, n)).values())
    if name in ("Conv", "DWConv"):
        return _tree_conv(obj)
    if name == "Bottleneck":
        return {"t": "Bottleneck", "cv1": _tree_conv(_child(obj, "cv1")), "cv2": _tree_conv(_child(obj, "cv2")),
                "add": bool(_attr(obj, "add", False))}
    if name in ("C2f", "C3k2"):
        return {"t": "C2f", "cv1": _tree_conv(_child(obj, "cv1")), "cv2": _tree_conv(_child(obj, "cv2")),
                "m": [_tree_block(m) for m in kids("m")]}
    if name in ("C3", "C3k"):
        return {"t": "C3", "cv1": _tree_conv(_child(obj, "cv1")), "cv2": _tree_conv(_child(obj, "cv2")),
                "cv3": _tree_conv(_child(obj, "cv3")), "m": [_tree_block(m) for m in kids("m")]}
    if name == "SPPF":
        return {"t": "SPPF", "cv1": _tree_conv(_child(obj, "cv1")), "cv2": _tree_conv(_child(obj, "cv2")),
                "k": _first(_attr(_child(obj, "m"), "kernel_size", 5))}
    if name == "Attention":
        return {"t": "Attention", "num_heads": int(_attr(obj, "num_heads")), "head_dim": int(_attr(obj, "head_dim")),
                "key_dim": int(_attr(obj, "key_dim")), "scale": float(_attr(obj, "scale")),
                "qkv": _tree_conv(_child(obj, "qkv")), "proj": _tree_conv(_child(obj, "proj")), "pe": _tree_conv(_child(obj, "pe"))}
    if name == "PSABlock":
        return {"t": "PSABlock", "attn": _tree_block(_child(obj, "attn")), "ffn": _flat_convs(_child(obj, "ffn")),
                "add": bool(_attr(obj, "add", True))}
    if name == "C2PSA":
        return {"t": "C2PSA", "c": int(_attr(obj, "c")), "cv1": _tree_conv(_child(obj, "cv1")),
                "cv2": _tree_conv(_child(obj, "cv2")), "m": [_tree_block(m) for m in kids("m")]}
    if name == "AAttn":
        if "qkv" not in _mods(obj):
            raise UnsupportedCheckpoint("area attention with separate qk / v projections (the yolov12 fork's layout)")
        return {"t": "AAttn", "area": int(_attr(obj, "area", 1)), "num_heads": int(_attr(obj, "num_heads")),
                "head_dim": int(_attr(obj, "head_dim")), "qkv": _tree_conv(_child(obj, "qkv")),
                "proj": _tree_conv(_child(obj, "proj")), "pe": _tree_conv(_child(obj, "pe"))}
    if name == "ABlock":
        return {"t": "ABlock", "attn": _tree_block(_child(obj, "attn")), "mlp": _flat_convs(_child(obj, "mlp"))}
    if name == "A2C2f":
        m = []
        for sub in kids("m"):
            m.append([_tree_block(a) for a in _mods(sub).values()] if _cls(sub) == "Sequential" else _tree_block(sub))
        gamma = _attr(obj, "gamma")
        return {"t": "A2C2f", "cv1": _tree_conv(_child(obj, "cv1")), "cv2": _tree_conv(_child(obj, "cv2")),
                "gamma": gamma.detach().float() if torch.is_tensor(gamma) else None, "m": m}
    raise UnsupportedCheckpoint(f"module {name} is not supported by this build")


def tree_from_ultralytics_model(root) -> dict:
    """Unpickled `DetectionModel` (inert bags, see `_load_pickle_inert`) -> the node tree YoloTreeB200 executes."""
    seq = _mods(root).get("model")
    if seq is None:
        raise UnsupportedCheckpoint("no `model` Sequential in the checkpoint object")
    layers = []
    for idx, (key, m) in enumerate(_mods(seq).items()):
        name = _cls(m)
        f = _attr(m, "f", -1)
        f = [int(v) for v in f] if isinstance(f, (list, tuple)) else int(f)
        if name == "Concat":
            node = {"t": "Concat", "d": int(_attr(m, "d", 1))}
        elif name == "Upsample":
            sf, mode = _attr(m, "scale_factor", 2), _attr(m, "mode", "nearest")
            if mode != "nearest" or float(_first(sf)) != 2.0:
                raise UnsupportedCheckpoint(f"Upsample(scale_factor={sf}, mode={mode})")
            node = {"t": "Upsample", "scale": 2}
        elif name in ("Detect", "Segment"):
            if bool(_attr(m, "end2end", False)):
                raise UnsupportedCheckpoint("end-to-end (NMS-free) Detect head")
            stride = _attr(m, "stride")
            node = {"t": name, "nc": int(_attr(m, "nc")), "reg_max": int(_attr(m, "reg_max", 16)),
                    "stride": [int(round(float(v))) for v in (stride.tolist() if torch.is_tensor(stride) else stride)],
                    "cv2": [_flat_convs(b) for b in _mods(_child(m, "cv2")).values()],
                    "cv3": [_flat_convs(b) for b in _mods(_child(m, "cv3")).values()]}
            if name == "Segment":
                pr = _child(m, "proto")
                up = _child(pr, "upsample")
                uw, ub = _attr(up, "weight"), _attr(up, "bias")
                if _cls(up) != "ConvTranspose2d" or uw is None or tuple(uw.shape[2:]) != (2, 2) or _first(_attr(up, "stride")) != 2:
                    raise UnsupportedCheckpoint("prototype upsample is not ConvTranspose2d(k=2, s=2)")
                node.update(nm=int(_attr(m, "nm")), cv4=[_flat_convs(b) for b in _mods(_child(m, "cv4")).values()],
                            proto={"cv1": _tree_conv(_child(pr, "cv1")),
                                   "upsample": {"w": uw.detach().float().contiguous(),
                                                "b": (ub.detach().float() if ub is not None else torch.zeros(uw.shape[1]))},
                                   "cv2": _tree_conv(_child(pr, "cv2")), "cv3": _tree_conv(_child(pr, "cv3"))})
        elif name in ("Pose", "OBB", "Classify", "RTDETRDecoder", "v10Detect", "YOLOEDetect", "WorldDetect"):
            raise UnsupportedCheckpoint(f"{name} head: this loader handles `Detect` and `Segment` models")
        else:
            node = _tree_block(m)
        node["f"] = f
        layers.append(node)
    if not layers or layers[-1]["t"] not in ("Detect", "Segment"):
        raise UnsupportedCheckpoint("the model does not end in a Detect / Segment head")
    names = _attr(root, "names")
    return {"layers": layers, "names": dict(names) if isinstance(names, dict) else None}


def load_ultralytics_tree(path: str) -> dict:
    """Node tree of the detection model OBJECT an ultralytics `.pt` holds (`ckpt["ema"] or ckpt["model"]`, what
    `YOLO(path)` runs: core/ml/model_manager.py:804,830).  Plain state-dict files carry no architecture and are refused."""
    obj = _load_pickle_inert(path)
    root = obj
    if isinstance(obj, dict):
        root = obj.get("ema") or obj.get("model")
        if root is None or isinstance(root, dict):
            raise UnsupportedCheckpoint(f"{path}: a plain state dict carries no module tree; the pickled model object is needed")
    tree = tree_from_ultralytics_model(root)
    if tree["names"] is None and isinstance(obj, dict) and isinstance(obj.get("names"), dict):
        tree["names"] = dict(obj["names"])
    return tree


# ---- SAM 2.1 ----------------------------------------------------------------------------------------------------------
def sam2_model_and_state(seed: int = 0, variant: str = "tiny"):
    """(config, state_dict) of seeded SYNTHETIC weights, initialised by the library the reference uses (`transformers`,
    core/ml/model_manager.py:996-1005); no forward pass happens here.  Real checkpoints are resolved by
    ModelManager.load_sam2."""
    from transformers import Sam2Config, Sam2Model
    torch.manual_seed(seed)
    if variant == "tiny":
        cfg = Sam2Config()
    else:  # hiera-large: the checkpoint the reference actually loads (core/ml/model_manager.py:202-204)
        cfg = Sam2Config()
        bc = cfg.vision_config.backbone_config
        bc.hidden_size, bc.embed_dim_per_stage = 144, [144, 288, 576, 1152]
        bc.blocks_per_stage, bc.num_attention_heads_per_stage = [2, 6, 36, 4], [2, 4, 8, 16]
        bc.global_attention_blocks, bc.window_size_per_stage = [23, 33, 43], [8, 4, 16, 8]
        bc.window_positional_embedding_background_size = [7, 7]
        cfg.vision_config.backbone_channel_list = [1152, 576, 288, 144]
    m = Sam2Model(cfg).eval()
    with torch.no_grad():
        for mlp in m.mask_decoder.output_hypernetworks_mlps:   # default init gives |logit| ~ 1e-2: widen for real masks
            mlp.proj_out.weight.mul_(100.0)
    return cfg, m.state_dict()


# ---- RT-DETRv2 (secondary conjoined / fallback bubble detector) ------------------------------------------------------
RTDETR_NAMES = {0: "bubble", 1: "text_bubble", 2: "text_free"}     # classes the stage code looks for (detection.py:1430-1437)


def rtdetr_model_and_state(seed: int = 0, checkpoint_dir: Optional[str] = None, **config_overrides):
    """(config, state_dict) of `RTDetrV2ForObjectDetection`.  Like the SAM weights, initialisation and (when
    `checkpoint_dir` or `$MT_MODELS_DIR/rtdetr/comic-text-and-bubble-detector` exists) loading go through the library the
    reference uses
    (core/ml/model_manager.py:758-766); no forward pass happens here.  Synthetic weights: seeded default init, non-trivial
    frozen-batch-norm statistics, and class biases lifted from the focal-loss prior so a few dozen queries pass conf 0.35."""
    from transformers import RTDetrV2Config, RTDetrV2ForObjectDetection
    d = models_dir()
    ckpt = checkpoint_dir or (os.path.join(d, "rtdetr", "comic-text-and-bubble-detector") if d else None)
    if ckpt and os.path.isdir(ckpt):
        m = RTDetrV2ForObjectDetection.from_pretrained(ckpt).eval()
        return m.config, m.state_dict()
    cfg = RTDetrV2Config(num_labels=len(RTDETR_NAMES), id2label=dict(RTDETR_NAMES),
                         label2id={v: k for k, v in RTDETR_NAMES.items()}, **config_overrides)
    torch.manual_seed(seed)
    model = RTDetrV2ForObjectDetection(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) * 0.5 + 0.75)
        for name, p in model.named_parameters():
            is_bn = "normalization." in name or ".norm." in name or ("input_proj" in name and ".1." in name)
            if is_bn and name.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) * 0.4 + 0.8)
            elif is_bn and name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        for head in list(model.class_embed) + [model.model.enc_score_head]:
            head.bias.copy_(torch.tensor([-0.2, -1.0, -0.8]))
            head.weight.mul_(3.0)
        # A randomly initialised decoder gives all queries nearly the same class logits.  Calibrate the last class head on
        # one seeded noise image (a CPU forward of the library model, synthetic weights only) so that ~20 % of the queries
        # are "bubble" detections above conf 0.35, ~5 % "text_free", and no "text_bubble".
        last = model.class_embed[-1]
        last.weight.mul_(4.0)
        probe = torch.rand((1, 3, 640, 640), generator=g)
        logits = model(pixel_values=probe).logits[0]
        thr = math.log(0.35 / 0.65)
        for c, q in ((0, 0.80), (1, 1.0), (2, 0.95)):
            cut = torch.quantile(logits[:, c], q).item() if q < 1.0 else logits[:, c].max().item() + 1.0
            last.bias[c] -= cut - thr
    return cfg, model.state_dict()
