"""CPU: the panel (YOLO11) / OSB-text (YOLO12) detector trees — reference call sites core/image/detection.py:1817-1921,
:120-201; loaders core/ml/model_manager.py:780-833.

* an ultralytics-style checkpoint (pickled DetectionModel OBJECT, un-fused BatchNorm, classes of a package that is then
  forgotten) is read back into exactly the node tree it was made from (BatchNorm folded);
* heads other than Detect and plain state dicts are refused with a clear error;
* the oracle runs both families and the post-processing yields boxes inside the image."""
import os
import sys
import textwrap

import numpy as np
import pytest
import torch

from mangatranslator_b200 import synth, weights as W
from mangatranslator_b200 import yolo_tree as T

FAKE = textwrap.dedent('''
    import torch, torch.nn as nn
    G = torch.Generator().manual_seed(7)

    class Conv(nn.Module):
        default_act = nn.SiLU()
        def __init__(self, node):
            super().__init__()
            w, b = node["w"], node["b"]
            co, cig, k, _ = w.shape
            self.conv = nn.Conv2d(cig * node["g"], co, k, node["s"], node["p"], groups=node["g"], bias=False)
            self.bn = nn.BatchNorm2d(co, eps=1e-3)
            self.act = self.default_act if node["act"] else nn.Identity()
            gamma, var = torch.rand(co, generator=G) + 0.5, torch.rand(co, generator=G) + 0.5
            mean = torch.randn(co, generator=G) * 0.1
            kk = gamma / torch.sqrt(var + 1e-3)
            with torch.no_grad():
                self.conv.weight.copy_(w / kk.view(-1, 1, 1, 1))
                self.bn.weight.copy_(gamma); self.bn.running_var.copy_(var); self.bn.running_mean.copy_(mean)
                self.bn.bias.copy_(b + mean * kk)

    class DWConv(Conv):
        pass

    def plain(node):
        co, ci, k, _ = node["w"].shape
        m = nn.Conv2d(ci, co, k)
        with torch.no_grad():
            m.weight.copy_(node["w"]); m.bias.copy_(node["b"])
        return m

    def conv(node):
        return (DWConv if node["g"] > 1 else Conv)(node)

    class Bottleneck(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.cv1, self.cv2, self.add = conv(n["cv1"]), conv(n["cv2"]), n["add"]

    class C3k(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.cv1, self.cv2, self.cv3 = conv(n["cv1"]), conv(n["cv2"]), conv(n["cv3"])
            self.m = nn.Sequential(*[make(b) for b in n["m"]])

    class C3k2(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.c = n["cv1"]["w"].shape[0] // 2
            self.cv1, self.cv2 = conv(n["cv1"]), conv(n["cv2"])
            self.m = nn.ModuleList(make(b) for b in n["m"])

    class SPPF(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.cv1, self.cv2 = conv(n["cv1"]), conv(n["cv2"])
            self.m = nn.MaxPool2d(kernel_size=n["k"], stride=1, padding=n["k"] // 2)

    class Attention(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.num_heads, self.head_dim, self.key_dim, self.scale = n["num_heads"], n["head_dim"], n["key_dim"], n["scale"]
            self.qkv, self.proj, self.pe = conv(n["qkv"]), conv(n["proj"]), Conv(n["pe"])

    class PSABlock(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.attn = Attention(n["attn"])
            self.ffn = nn.Sequential(*[conv(c) for c in n["ffn"]])
            self.add = n["add"]

    class C2f(C3k2):                 # the YOLOv8 name of the same block
        pass

    V8_NAMES = False

    class C2PSA(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.c = n["c"]
            self.cv1, self.cv2 = conv(n["cv1"]), conv(n["cv2"])
            self.m = nn.Sequential(*[PSABlock(b) for b in n["m"]])

    class AAttn(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.area, self.num_heads, self.head_dim = n["area"], n["num_heads"], n["head_dim"]
            self.qkv, self.proj, self.pe = conv(n["qkv"]), conv(n["proj"]), Conv(n["pe"])

    class ABlock(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.attn = AAttn(n["attn"])
            self.mlp = nn.Sequential(*[conv(c) for c in n["mlp"]])

    class A2C2f(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.cv1, self.cv2 = conv(n["cv1"]), conv(n["cv2"])
            self.gamma = nn.Parameter(n["gamma"].clone()) if n["gamma"] is not None else None
            self.m = nn.ModuleList(nn.Sequential(*[ABlock(a) for a in m]) if isinstance(m, list) else C3k(m) for m in n["m"])

    class Concat(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.d = n["d"]

    class DFL(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(16, 1, 1, bias=False)

    class Detect(nn.Module):
        end2end = False
        def __init__(self, n):
            super().__init__()
            self.nc, self.reg_max, self.nl = n["nc"], n["reg_max"], len(n["cv2"])
            self.stride = torch.tensor([float(v) for v in n["stride"]])
            self.cv2 = nn.ModuleList(nn.Sequential(conv(b[0]), conv(b[1]), plain(b[2])) for b in n["cv2"])
            self.cv3 = nn.ModuleList((nn.Sequential(conv(b[0]), conv(b[1]), plain(b[2])) if len(b) == 3 else       # YOLOv8 (legacy)
                                      nn.Sequential(nn.Sequential(conv(b[0]), conv(b[1])), nn.Sequential(conv(b[2]), conv(b[3])),
                                                    plain(b[4]))) for b in n["cv3"])
            self.dfl = DFL()

    class Proto(nn.Module):
        def __init__(self, n):
            super().__init__()
            self.cv1, self.cv2, self.cv3 = conv(n["cv1"]), conv(n["cv2"]), conv(n["cv3"])
            ci, co = n["upsample"]["w"].shape[:2]
            self.upsample = nn.ConvTranspose2d(ci, co, 2, 2, 0, bias=True)
            with torch.no_grad():
                self.upsample.weight.copy_(n["upsample"]["w"]); self.upsample.bias.copy_(n["upsample"]["b"])

    class Segment(Detect):
        def __init__(self, n):
            super().__init__(n)
            self.nm = n["nm"]
            self.proto = Proto(n["proto"])
            self.cv4 = nn.ModuleList(nn.Sequential(conv(b[0]), conv(b[1]), plain(b[2])) for b in n["cv4"])

    class Pose(Detect):
        pass

    def make(n):
        return {"Conv": conv, "Bottleneck": Bottleneck, "C3": C3k, "C2f": C2f if V8_NAMES else C3k2, "SPPF": SPPF, "C2PSA": C2PSA, "A2C2f": A2C2f,
                "Concat": Concat, "Detect": Detect, "Segment": Segment,
                "Upsample": lambda n: nn.Upsample(None, 2, "nearest")}[n["t"]](n)

    class DetectionModel(nn.Module):
        def __init__(self, tree, head=None):
            super().__init__()
            mods = []
            for i, n in enumerate(tree["layers"]):
                m = head(n) if (head and n["t"] == "Detect") else make(n)
                m.f, m.i, m.type = n["f"], i, n["t"]
                mods.append(m)
            self.model = nn.Sequential(*mods)
            self.names = tree["names"]
            self.yaml = {"nc": len(tree["names"])}
''')


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) - {"family", "scale"} == set(b) - {"family", "scale"}, (path, set(a) ^ set(b))
        for k in a:
            if k not in ("family", "scale"):
                _same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif torch.is_tensor(a):
        assert torch.is_tensor(b) and a.shape == b.shape, path
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, float(a.abs().max())), path
    elif isinstance(a, float):
        assert abs(a - b) < 1e-9, path
    else:
        assert a == b, (path, a, b)


def _save_with_fake_package(tmp_path, tree, head=None, as_state_dict=False, v8_names=False):
    pkg = tmp_path / "fakepkg" / "ultralytics" / "nn"
    pkg.mkdir(parents=True, exist_ok=True)
    (tmp_path / "fakepkg" / "ultralytics" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "tasks.py").write_text(FAKE)
    stash = {n: sys.modules.pop(n) for n in [n for n in sys.modules if n == "ultralytics" or n.startswith("ultralytics.")]}
    sys.path.insert(0, str(tmp_path / "fakepkg"))
    path = tmp_path / "model.pt"
    try:
        import ultralytics.nn.tasks as tasks
        tasks.V8_NAMES = v8_names
        model = tasks.DetectionModel(tree, getattr(tasks, head) if head else None)
        torch.save({"model": model.state_dict() if as_state_dict else model, "epoch": 1}, str(path))
    finally:
        sys.path.remove(str(tmp_path / "fakepkg"))
        for name in [n for n in sys.modules if n == "ultralytics" or n.startswith("ultralytics.")]:
            del sys.modules[name]
    return path, stash


@pytest.mark.parametrize("family,kw", [("11", {}), ("12", dict(a2_residual=True, mlp_ratio=1.2)), ("11", dict(segment=True))],
                         ids=["yolo11", "yolo12", "yolo11-seg"])
def test_checkpoint_object_is_read_back_into_the_tree_it_was_made_from(tmp_path, family, kw):
    tree = T.synthetic_tree(family, "s", nc=2, seed=3, names={0: "frame", 1: "text"}, **kw)
    path, stash = _save_with_fake_package(tmp_path, tree)
    try:
        with pytest.raises(ImportError):
            import ultralytics  # noqa: F401
        got = W.load_ultralytics_tree(str(path))
    finally:
        sys.modules.update(stash)
    assert got["names"] == {0: "frame", 1: "text"}
    _same(tree["layers"], got["layers"])
    if family == "12":
        assert any(n["t"] == "A2C2f" and n["gamma"] is not None for n in got["layers"])
        assert any(int(n["mlp"][0]["w"].shape[0]) % 16 for l in got["layers"] if l["t"] == "A2C2f" for m in l["m"]
                   if isinstance(m, list) for n in m)                      # the 1.2x MLP width is not a multiple of 16


def test_other_heads_and_plain_state_dicts_are_refused(tmp_path):
    tree = T.synthetic_tree("11", "s", nc=1, seed=0)
    path, stash = _save_with_fake_package(tmp_path, tree, head="Pose")
    try:
        with pytest.raises(W.UnsupportedCheckpoint, match="Pose"):
            W.load_ultralytics_tree(str(path))
    finally:
        sys.modules.update(stash)
    path, stash = _save_with_fake_package(tmp_path, tree, as_state_dict=True)
    try:
        with pytest.raises(W.UnsupportedCheckpoint, match="state dict"):
            W.load_ultralytics_tree(str(path))
    finally:
        sys.modules.update(stash)


@pytest.mark.parametrize("family", ["11", "12"])
def test_oracle_runs_both_families(family):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import yolo_tree_oracle as O
    pg = synth.make_page(3, 450, 320)
    bgr = np.ascontiguousarray(pg.image_rgb[:, :, ::-1])
    tree = T.synthetic_tree(family, "s", nc=2, seed=1)
    O.calibrate(tree, bgr, 320, cls_mean=-5.0, cls_std=1.5)
    r = O.predict(tree, bgr, 0.25, 320)
    assert [tuple(b.shape[1:]) for b, _ in r["heads"]] == [(64, 40, 32), (64, 20, 16), (64, 10, 8)]
    n = len(r["conf"])
    assert 0 < n <= 300
    b = r["xyxy"].numpy()
    assert b[:, 0].min() >= 0 and b[:, 1].min() >= 0 and b[:, 2].max() <= 320 and b[:, 3].max() <= 450
    assert np.all(np.diff(r["conf"].numpy()) <= 0)
    # published parameter counts of the "s" scale (YOLO11s 9.4 M, YOLO12s 9.3 M): the restated layouts have the right size
    cnt = [0]

    def count(v):
        if isinstance(v, dict):
            if v.get("t") == "Conv":
                cnt[0] += v["w"].numel() + v["b"].numel()
            for x in v.values():
                count(x)
        elif isinstance(v, list):
            for x in v:
                count(x)
    count(T.synthetic_tree(family, "s", nc=80, seed=0)["layers"])
    assert abs(cnt[0] / 1e6 - {"11": 9.4, "12": 9.3}[family]) < 0.25, cnt[0]


@pytest.mark.parametrize("family,scale,published", [("11", "l", 25.3), ("12", "x", 59.1)])
def test_layouts_of_the_reference_checkpoint_scales_have_the_published_size(family, scale, published):
    """The panel model is a YOLO11-L and the OSB-text model a YOLO12x (core/ml/model_manager.py:129-132): the restated
    yaml layouts at those scales hold the parameter counts ultralytics publishes for them (80 classes)."""
    total = [0]

    def count(v):
        if isinstance(v, dict):
            if v.get("t") == "Conv":
                total[0] += v["w"].numel() + v["b"].numel()
            if v.get("gamma") is not None:
                total[0] += v["gamma"].numel()
            for x in v.values():
                count(x)
        elif isinstance(v, list):
            for x in v:
                count(x)
    count(T.synthetic_tree(family, scale, nc=80, seed=0)["layers"])
    assert abs(total[0] / 1e6 - published) < 0.06, total[0]


class _FakeBoxes:
    def __init__(self, xyxy, conf, cls):
        self.xyxy, self.conf, self.cls = xyxy, conf, cls

    def __len__(self):
        return len(self.xyxy)


class _FakeResults:
    def __init__(self, boxes):
        self.boxes = boxes


class _FakeDetector:
    names = {0: "body", 1: "face", 2: "frame", 3: "text"}

    def __init__(self, xyxy, cls=None):
        n = len(xyxy)
        self.out = _FakeResults(_FakeBoxes(torch.tensor(xyxy, dtype=torch.float32).reshape(n, 4),
                                           torch.linspace(0.9, 0.5, n), torch.tensor(cls if cls is not None else [0] * n).float()))
        self.calls = []

    def __call__(self, image, **kw):
        self.calls.append(kw)
        return [self.out]


def _random_case(rng):
    n, m = int(rng.integers(1, 8)), int(rng.integers(0, 10))
    xy = rng.uniform(0, 500, size=(n, 2))
    bubbles = np.concatenate([xy, xy + rng.uniform(40, 200, size=(n, 2))], 1).astype(np.float32)
    texts = []
    for _ in range(m):
        if rng.random() < 0.7:                       # near a bubble: inside, sticking out, or just touching
            b = bubbles[int(rng.integers(0, n))]
            cx, cy = rng.uniform(b[0] - 20, b[2] + 20), rng.uniform(b[1] - 20, b[3] + 20)
        else:
            cx, cy = rng.uniform(0, 700, size=2)
        w, h = rng.uniform(5, 120, size=2)
        texts.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    return bubbles, np.asarray(texts, np.float32).reshape(m, 4)


def test_osb_text_expansion_and_panels_match_the_live_reference():
    """`_expand_boxes_with_osb_text` (:120-201) and `detect_panels` (:1817-1921) with the same fake detector injected into
    both model managers: identical boxes, identical call arguments (imgsz 640, the caller's confidence)."""
    import _refimport
    if not _refimport.available():
        pytest.skip("reference tree not present")
    _refimport.import_reference()
    import core.image.detection as ref
    from core.caching import get_cache as ref_cache
    from core.ml.model_manager import ModelType as RefType, get_model_manager as ref_mm
    from PIL import Image
    from mangatranslator_b200.core.caching import get_cache
    from mangatranslator_b200.core.image import detection as ours
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    rng = np.random.default_rng(5)
    img = Image.fromarray(rng.integers(0, 255, size=(700, 700, 3), dtype=np.uint8))
    cv = np.ascontiguousarray(np.asarray(img)[:, :, ::-1])
    mm, rmm = get_model_manager(), ref_mm()
    try:
        for trial in range(40):
            bubbles, texts = _random_case(rng)
            fake_o, fake_r = _FakeDetector(texts), _FakeDetector(texts)
            mm.models[ModelType.YOLO_OSBTEXT], rmm.models[RefType.YOLO_OSBTEXT] = fake_o, fake_r
            get_cache().clear_all()
            ref_cache().clear_all() if hasattr(ref_cache(), "clear_all") else None
            conf = 0.3 + 0.01 * trial                   # a new cache key per trial on both sides
            pb = torch.from_numpy(bubbles)
            want = ref._expand_boxes_with_osb_text(cv, img, pb.clone(), ref_cache(), rmm, torch.device("cpu"), conf, "", False)
            got = ours._expand_boxes_with_osb_text(cv, img, pb.clone(), get_cache(), mm, torch.device("cpu"), conf, "", False)
            assert torch.equal(want, got), (trial, want, got)
            assert fake_o.calls == fake_r.calls and fake_o.calls[0]["imgsz"] == 640 and fake_o.calls[0]["conf"] == conf
            # a second call is served from the detection cache on both sides
            ours._expand_boxes_with_osb_text(cv, img, pb.clone(), get_cache(), mm, torch.device("cpu"), conf, "", False)
            assert len(fake_o.calls) == 1
        for trial in range(10):
            n = int(rng.integers(0, 9))
            xy = rng.uniform(0, 500, size=(n, 2))
            boxes = np.concatenate([xy, xy + rng.uniform(30, 190, size=(n, 2))], 1)
            cls = rng.integers(0, 4, size=n).tolist()
            fake_o, fake_r = _FakeDetector(boxes, cls), _FakeDetector(boxes, cls)
            mm.models[ModelType.YOLO_PANEL], rmm.models[RefType.YOLO_PANEL] = fake_o, fake_r
            want = ref.detect_panels(None, 0.25, torch.device("cpu"), False, image_override=img)
            got = ours.detect_panels(None, 0.25, torch.device("cpu"), False, image_override=img)
            assert want == got and all(isinstance(v, int) for b in got for v in b)
            assert fake_o.calls == fake_r.calls
            assert len(got) == sum(1 for c in cls if c == 2)
    finally:
        for t in (ModelType.YOLO_OSBTEXT, ModelType.YOLO_PANEL):
            mm.models.pop(t, None)
        for t in (RefType.YOLO_OSBTEXT, RefType.YOLO_PANEL):
            rmm.models.pop(t, None)


def test_yolov8_named_detection_checkpoint_reads_into_the_same_tree(tmp_path):
    """A YOLOv8 detection model (class names `C2f`, the legacy class branch of plain 3x3 convs) is the same tree: whatever
    the hard-wired YOLOv8-seg reader refuses (a detection-only `yolo_1`-family file) still runs from its module tree."""
    s = T._Synth(5, "s", 2)
    ch = s.ch
    layers = []

    def add(node, f=-1):
        node = dict(node)
        node["f"] = f
        layers.append(node)
        return len(layers) - 1
    add(s.conv(3, ch(64), 3, 2))
    add(s.conv(ch(64), ch(128), 3, 2))
    p3 = add(s.c3k2(ch(128), ch(256), 1, False, 0.5))                # C2f(c, c, n=1, shortcut=True)
    add(s.conv(ch(256), ch(512), 3, 2))
    p4 = add(s.c3k2(ch(512), ch(512), 2, False, 0.5))
    add(s.conv(ch(512), ch(1024), 3, 2))
    p5 = add(s.sppf(ch(1024), ch(1024)))
    add(s.detect([ch(256), ch(512), ch(1024)], legacy=True), f=[p3, p4, p5])
    tree = {"layers": layers, "names": {0: "a", 1: "b"}}
    path, stash = _save_with_fake_package(tmp_path, tree, v8_names=True)
    try:
        got = W.load_ultralytics_tree(str(path))
    finally:
        sys.modules.update(stash)
    _same(tree["layers"], got["layers"])
    assert [len(b) for b in got["layers"][-1]["cv3"]] == [3, 3, 3]
