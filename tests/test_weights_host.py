"""CPU: checkpoint ingestion and the weight policy of the loaders (reference: core/ml/model_manager.py:183-190, 711-743).

* an ultralytics-style YOLOv8-seg checkpoint — a pickled model OBJECT with un-fused BatchNorm under `model.N.*` names — is
  read without ultralytics installed and converted (BN folded) to exactly the weights the oracle network holds;
* YOLO11-style layouts and non-segmentation heads are refused with a clear error;
* a missing checkpoint is an error unless MTB200_SYNTHETIC_WEIGHTS=1."""
import os
import sys
import textwrap

import pytest
import torch

from mangatranslator_b200 import weights as W


def _ultralytics_named(sd, seed=1):
    """Our folded layout -> ultralytics names with a random (invertible) BatchNorm in front of every folded conv."""
    g = torch.Generator().manual_seed(seed)
    ul = {}
    for k, v in sd.items():
        parts = k.split(".")
        if k.startswith("head.proto.upsample") or (k.startswith("head.cv") and parts[3] == "2"):
            ul["model.22." + k[len("head."):]] = v.clone()
            continue
        if not k.endswith(".conv.weight"):
            continue
        base = k[:-len(".conv.weight")]
        src = ("model.22." + base[len("head."):]) if base.startswith("head.") else ("model." + base[1:])
        co = v.shape[0]
        gamma, var = torch.rand(co, generator=g) + 0.5, torch.rand(co, generator=g) + 0.5
        mean = torch.randn(co, generator=g) * 0.1
        kk = gamma / torch.sqrt(var + 1e-3)
        ul[src + ".conv.weight"] = v / kk.view(-1, 1, 1, 1)
        ul[src + ".bn.weight"], ul[src + ".bn.running_var"], ul[src + ".bn.running_mean"] = gamma, var, mean
        ul[src + ".bn.bias"] = sd[base + ".conv.bias"] + mean * kk
        ul[src + ".bn.num_batches_tracked"] = torch.tensor(0)
    ul["model.22.dfl.conv.weight"] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    return ul


@pytest.mark.parametrize("variant", ["n", "m"])
def test_ultralytics_names_fold_to_the_oracle_weights(variant):
    cfg = W.yolo_cfg(variant, nc=2)
    sd = W.yolo_state_dict(5, cfg)
    out, cfg2 = W.yolo_from_ultralytics(_ultralytics_named(sd))
    assert cfg2 == cfg
    assert set(out) == set(sd)
    for k in sd:
        assert out[k].shape == sd[k].shape
        assert (out[k] - sd[k]).abs().max().item() <= 1e-6 * max(1.0, float(sd[k].abs().max())), k
    # and the oracle network accepts the converted dict as it is
    import yolo_oracle
    m = yolo_oracle.YoloV8Seg(**cfg2)
    m.load_state_dict(out)


def test_pickled_model_object_is_read_without_ultralytics(tmp_path):
    """`YOLO(path)` files hold the model OBJECT (classes from the ultralytics package).  Write one with a stand-in package,
    forget the package, and read the tensors back through the inert-class unpickler."""
    pkg = tmp_path / "fakepkg" / "ultralytics" / "nn"
    pkg.mkdir(parents=True)
    (tmp_path / "fakepkg" / "ultralytics" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "tasks.py").write_text(textwrap.dedent("""
        import torch.nn as nn
        class Conv(nn.Module):
            def __init__(self, c1, c2, k):
                super().__init__()
                self.conv = nn.Conv2d(c1, c2, k, bias=False)
                self.bn = nn.BatchNorm2d(c2, eps=1e-3)
        class SegmentationModel(nn.Module):
            def __init__(self):
                super().__init__()
                self.model = nn.Sequential(Conv(3, 16, 3), Conv(16, 32, 3))
                self.names = {0: "bubble"}
                self.yaml = {"nc": 1}
    """))
    # other tests put a MagicMock "ultralytics" into sys.modules to import the reference: set it aside
    stash = {n: sys.modules.pop(n) for n in [n for n in sys.modules if n == "ultralytics" or n.startswith("ultralytics.")]}
    sys.path.insert(0, str(tmp_path / "fakepkg"))
    try:
        from ultralytics.nn.tasks import SegmentationModel
        torch.manual_seed(0)
        model = SegmentationModel()
        expect = {k: v.clone() for k, v in model.state_dict().items()}
        path = tmp_path / "best.pt"
        torch.save({"model": model, "epoch": 3, "train_args": {"imgsz": 1600}}, str(path))
    finally:
        sys.path.remove(str(tmp_path / "fakepkg"))
        for name in [n for n in sys.modules if n == "ultralytics" or n.startswith("ultralytics.")]:
            del sys.modules[name]
    with pytest.raises(ImportError):
        import ultralytics  # noqa: F401  (really gone)
    try:
        sd, names = W.load_ultralytics_state_dict(str(path))
    finally:
        sys.modules.update(stash)
    assert names == {0: "bubble"}
    assert set(sd) == set(expect)
    for k in expect:
        assert torch.equal(sd[k], expect[k].float())
    # a plain state-dict file is accepted too
    torch.save(expect, str(tmp_path / "sd.pt"))
    sd2, _ = W.load_ultralytics_state_dict(str(tmp_path / "sd.pt"))
    assert set(sd2) == set(expect)


def test_other_families_are_refused_with_a_clear_error():
    cfg = W.yolo_cfg("n")
    ul = _ultralytics_named(W.yolo_state_dict(1, cfg))
    y11 = dict(ul)
    y11["model.10.m.0.attn.qkv.conv.weight"] = torch.zeros(1)            # C2PSA attention of YOLO11
    with pytest.raises(W.UnsupportedCheckpoint, match="YOLO11"):
        W.yolo_from_ultralytics(y11)
    det_only = {k: v for k, v in ul.items() if ".proto." not in k and ".cv4." not in k}
    with pytest.raises(W.UnsupportedCheckpoint, match="segmentation"):
        W.yolo_from_ultralytics(det_only)
    with pytest.raises(W.UnsupportedCheckpoint):
        W.yolo_from_ultralytics({"foo": torch.zeros(1)})


def test_missing_checkpoint_is_an_error_without_the_explicit_opt_in(monkeypatch, tmp_path):
    from mangatranslator_b200.core.ml.model_manager import ModelManager
    from mangatranslator_b200.utils.exceptions import ModelError
    mm = ModelManager()
    monkeypatch.setenv("MTB200_SYNTHETIC_WEIGHTS", "0")
    with pytest.raises(ModelError, match="MTB200_SYNTHETIC_WEIGHTS"):
        mm._synthetic_or_raise("YOLO speech-bubble detector", [tmp_path / "nothing.pt"])
    monkeypatch.setenv("MTB200_SYNTHETIC_WEIGHTS", "1")
    mm._synthetic_or_raise("YOLO speech-bubble detector", [tmp_path / "nothing.pt"])     # logs the source, does not raise
