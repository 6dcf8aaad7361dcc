"""CPU: the weight transformations the B200 RT-DETRv2 applies before it builds its conv plans — frozen-BatchNorm
folding, the AvgPool2d(2)->conv1x1 shortcut as one 2x2/stride-2 conv, RepVGG re-parameterisation — against the unfused
torch modules; the synthetic-weight generator; and the loader's refusal to invent a secondary detector."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


def test_weight_folds_equal_the_unfused_modules():
    from mangatranslator_b200.rtdetr import fold_avgpool2, fold_conv_bn, fold_repvgg
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 12, 10, generator=g, dtype=torch.float64)

    def bn_params(c):
        return (torch.rand(c, generator=g, dtype=torch.float64) + 0.5, torch.randn(c, generator=g, dtype=torch.float64),
                torch.randn(c, generator=g, dtype=torch.float64), torch.rand(c, generator=g, dtype=torch.float64) + 0.5)

    w3 = torch.randn(6, 8, 3, 3, generator=g, dtype=torch.float64)
    ga, be, mu, va = bn_params(6)
    ref = F.batch_norm(F.conv2d(x, w3, padding=1), mu, va, ga, be, False, 0.0, 1e-5)
    wf, bf = fold_conv_bn(w3, ga, be, mu, va, 1e-5)
    assert (F.conv2d(x, wf, bf, padding=1) - ref).abs().max() < 1e-12
    # ResNet-vd shortcut
    w1 = torch.randn(6, 8, 1, 1, generator=g, dtype=torch.float64)
    ref = F.batch_norm(F.conv2d(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True), w1), mu, va, ga, be, False, 0.0, 1e-5)
    wf, bf = fold_conv_bn(w1, ga, be, mu, va, 1e-5)
    assert (F.conv2d(x, fold_avgpool2(wf), bf, stride=2) - ref).abs().max() < 1e-12
    # RepVGG block: conv3x3+BN and conv1x1+BN summed
    w3b = torch.randn(8, 8, 3, 3, generator=g, dtype=torch.float64)
    w1b = torch.randn(8, 8, 1, 1, generator=g, dtype=torch.float64)
    p3, p1 = bn_params(8), bn_params(8)
    ref = (F.batch_norm(F.conv2d(x, w3b, padding=1), p3[2], p3[3], p3[0], p3[1], False, 0.0, 1e-5) +
           F.batch_norm(F.conv2d(x, w1b), p1[2], p1[3], p1[0], p1[1], False, 0.0, 1e-5))
    w, b = fold_repvgg(fold_conv_bn(w3b, p3[0], p3[1], p3[2], p3[3], 1e-5), fold_conv_bn(w1b, p1[0], p1[1], p1[2], p1[3], 1e-5))
    assert (F.conv2d(x, w, b, padding=1) - ref).abs().max() < 1e-12


def test_synthetic_weights_are_seeded_and_calibrated():
    from mangatranslator_b200 import weights as W
    cfg, sd = W.rtdetr_model_and_state(0)
    cfg2, sd2 = W.rtdetr_model_and_state(0)
    assert set(sd) == set(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert dict(cfg.id2label) == W.RTDETR_NAMES and sd["class_embed.0.weight"].shape[0] == 3
    assert sd["model.backbone.model.embedder.embedder.0.normalization.running_var"].std() > 0     # non-trivial BN stats


def test_loader_does_not_invent_a_secondary_detector(monkeypatch, tmp_path):
    """Without a checkpoint the reference's load fails and detection keeps the primaries (detection.py:1541-1548); seeded
    random weights would be merged in as 'missed bubbles', so the loader refuses unless explicitly asked."""
    from mangatranslator_b200.core.ml.model_manager import ModelType, get_model_manager
    from mangatranslator_b200.utils.exceptions import ModelError
    mm = get_model_manager()
    mm.models.pop(ModelType.RTDETR_CONJOINED_BUBBLE, None)
    monkeypatch.delenv("MTB200_SYNTHETIC_RTDETR", raising=False)
    monkeypatch.setitem(mm.model_paths, ModelType.RTDETR_CONJOINED_BUBBLE, tmp_path / "absent")
    with pytest.raises(ModelError):
        mm.load_rtdetr_conjoined_bubble()


def _deform_attn_numpy(value, shapes, offsets, logits, ref, n_points, offset_scale):
    """The arithmetic of csrc/rtdetr_kernels.cu::deform_attn_kernel, in NumPy float32 (one image).
    value [tokens][heads][hd], offsets [Q][heads][L*P][2], logits [Q][heads][L*P], ref [Q][4] -> [Q][heads*hd]."""
    import numpy as np
    Q, heads, LP, _ = offsets.shape
    hd = value.shape[2]
    out = np.zeros((Q, heads, hd), np.float32)
    starts = np.cumsum([0] + [h * w for h, w in shapes])
    e = np.exp(logits - logits.max(-1, keepdims=True))
    wts = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    for q in range(Q):
        rx, ry, rw, rh = ref[q]
        for h in range(heads):
            for l, (H, W) in enumerate(shapes):
                for pnt in range(n_points):
                    i = l * n_points + pnt
                    lx = rx + offsets[q, h, i, 0] * (1.0 / n_points) * rw * offset_scale
                    ly = ry + offsets[q, h, i, 1] * (1.0 / n_points) * rh * offset_scale
                    ix, iy = ((2 * lx - 1 + 1) * W - 1) * 0.5, ((2 * ly - 1 + 1) * H - 1) * 0.5
                    ix = ix if -4.0 < ix < 1.0e6 else -4.0
                    iy = iy if -4.0 < iy < 1.0e6 else -4.0
                    x0, y0 = int(np.floor(ix)), int(np.floor(iy))
                    ax, ay = ix - x0, iy - y0
                    sv = np.zeros(hd, np.float32)
                    for yy, xx, w8 in ((y0, x0, (1 - ax) * (1 - ay)), (y0, x0 + 1, ax * (1 - ay)),
                                       (y0 + 1, x0, (1 - ax) * ay), (y0 + 1, x0 + 1, ax * ay)):
                        if 0 <= yy < H and 0 <= xx < W:
                            sv += np.float32(w8) * value[starts[l] + yy * W + xx, h]
                    out[q, h] += sv * wts[q, h, i]
    return out.reshape(Q, heads * hd)


def test_deformable_attention_arithmetic_matches_transformers():
    """The kernel's restatement of `multi_scale_deformable_attention_v2` (softmax over levels x points, reference-box
    scaled offsets, bilinear grid_sample with zero padding, weighted sum) against the library function itself on CPU,
    with sampling points inside, on the border of, and far outside the feature maps."""
    import numpy as np
    from transformers.models.rt_detr_v2.modeling_rt_detr_v2 import multi_scale_deformable_attention_v2
    g = torch.Generator().manual_seed(0)
    shapes, heads, hd, Q, P_ = [(10, 12), (5, 6), (3, 3)], 2, 32, 7, 4
    n_tok = sum(h * w for h, w in shapes)
    value = torch.randn(1, n_tok, heads, hd, generator=g)
    offsets = torch.randn(1, Q, heads, len(shapes) * P_, 2, generator=g) * 3.0
    offsets[0, 0] *= 40.0                                               # far outside
    logits = torch.randn(1, Q, heads, len(shapes) * P_, generator=g)
    ref = torch.rand(1, Q, 4, generator=g)
    ref[0, 1] = torch.tensor([0.0, 1.0, 0.3, 0.3])                      # on the border
    attn = torch.softmax(logits, -1)
    scale = torch.tensor([1.0 / P_] * (len(shapes) * P_)).unsqueeze(-1)
    loc = ref[:, :, None, None, :2] + offsets * scale * ref[:, :, None, None, 2:] * 0.5
    exp = multi_scale_deformable_attention_v2(value, shapes, loc, attn, [P_] * len(shapes), "default")[0].numpy()
    got = _deform_attn_numpy(value[0].numpy(), shapes, offsets[0].numpy(), logits[0].numpy(), ref[0].numpy(), P_, 0.5)
    assert np.abs(got - exp).max() < 2e-5
