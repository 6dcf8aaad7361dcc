"""B200 RT-DETRv2 — the object `ModelManager.load_rtdetr_conjoined_bubble()` returns (secondary "conjoined / fallback"
bubble detector of the reference: core/ml/model_manager.py:745-778 builds `RTDetrYOLOAdapter(RTDetrV2ForObjectDetection,
RTDetrImageProcessor)`, core/image/detection.py:1392-1548 calls it like a YOLO model with conf=0.35, imgsz=640).

Call contract (core/ml/rtdetr_adapter.py:61-113): `m(image_bgr_or_pil, conf=..., device=..., verbose=False, imgsz=640)
-> [Results]` with `Results.boxes.xyxy / .conf / .cls` (device tensors, original-pixel boxes, descending score),
`Results.names`, and `m.names`.

Network (transformers `RTDetrV2ForObjectDetection`, restated layer for layer from the state dict; arithmetic lives in
the third-party library, so parity is against that library with seeded weights — no checkpoint offline):
  * processor: uint8 bilinear antialias resize to imgsz x imgsz (torchvision backend) and /255 -> `mtb_resize_aa_u8` +
    `mtb_image_to_planes`
  * ResNet-vd backbone: conv + frozen BN folded into tcgen05 conv plans (bias + ReLU in the epilogue, residual add +
    ReLU of a bottleneck in its last conv's epilogue); the `AvgPool2d(2) -> conv1x1` shortcut is ONE 2x2 stride-2 conv
    with the 1x1 weights / 4 on every tap; stem max-pool = `mtb_maxpool2d`
  * hybrid encoder: 1x1 projections, AIFI transformer layer on the stride-32 map (sin-cos positions added to q and k,
    dense attention kernel, erf-GELU FFN, post-norm), CCFM top-down / bottom-up fusion where every RepVGG block
    (3x3 + 1x1, summed) is re-parameterised into one 3x3 conv and concats are channel-slice writes
  * query selection: encoder head on all 8400 tokens (invalid anchors zeroed), top-300 by best class score
  * 6 decoder layers: self-attention over the queries, multi-scale deformable attention (`mtb_deform_attn`), FFN,
    iterative box refinement; class head of the last layer
  * post-processing: sigmoid, top-300 over (query, class), cxcywh -> xyxy in original pixels, `score > conf`.
Every matrix product runs as bf16x3 (fp32-grade) conv plans; what is left in torch is index plumbing on a few hundred
values (top-k, gathers, the 300x4 box refinement).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import planes as P
from ._lib import check, lib, ptr, stream_ptr
from .ops import ConvPlan
from .preproc import resize_aa_device
from .sam2 import AttnDesc
from .sam2 import _declare as _declare_sam


def _declare(l) -> None:
    _declare_sam(l)
    if getattr(l, "_rtdetr_declared", False):
        return
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    l.mtb_maxpool2d.argtypes = [vp, vp] + [i32] * 8 + [vp]
    l.mtb_deform_attn.argtypes = [vp, i64, i32, i32, i32, i32, i32, C.POINTER(i32), i32, vp, i32, vp, i32, vp, i32, f32, vp,
                                  i64, vp]
    l.mtb_image_to_planes.argtypes = [vp, i32, i32, i32, i32, f32, C.POINTER(f32), vp, i32, i32, vp]
    for n in ("mtb_maxpool2d", "mtb_deform_attn", "mtb_image_to_planes"):
        getattr(l, n).restype = i32
    l._rtdetr_declared = True


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


# ---- weight folding (pure tensor functions; checked on CPU against the unfused torch modules) -----------------------
def fold_conv_bn(w, gamma, beta, mean, var, eps: float):
    """bias-free conv followed by an eval-mode / frozen BatchNorm -> (weight, bias) of one conv."""
    scale = gamma * (var + eps).rsqrt()
    return w * scale.view(-1, 1, 1, 1), beta - mean * scale


def fold_avgpool2(w1x1):
    """AvgPool2d(2, 2) followed by a 1x1 conv == a 2x2 stride-2 conv with a quarter of the 1x1 weights on every tap."""
    return (w1x1 / 4.0).expand(-1, -1, 2, 2).contiguous()


def fold_repvgg(conv3, conv1):
    """(w3, b3), (w1, b1) of the two folded branches of a RepVGG block -> one 3x3 conv (1x1 into the centre tap)."""
    return conv3[0] + torch.nn.functional.pad(conv1[0], (1, 1, 1, 1)), conv3[1] + conv1[1]


class _Boxes:
    def __init__(self, xyxy, conf, cls):
        self.xyxy, self.conf, self.cls = xyxy, conf, cls

    def __len__(self) -> int:
        return int(self.xyxy.shape[0])


class RtDetrB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config, device: torch.device, *, precision: str = "bf16x3",
                 names: Optional[Dict[int, str]] = None):
        self.l = lib()
        _declare(self.l)
        self.device = device
        self.planes = 2 if precision == "bf16x3" else 1
        self.sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()
                   if v.is_floating_point()}
        bc = _get(config, "backbone_config")
        self.depths = list(_get(bc, "depths"))
        self.hidden_sizes = list(_get(bc, "hidden_sizes"))
        self.embedding_size = int(_get(bc, "embedding_size"))
        if _get(bc, "layer_type") != "bottleneck" or _get(bc, "downsample_in_bottleneck") or _get(bc, "downsample_in_first_stage"):
            raise ValueError("RtDetrB200: only the bottleneck ResNet-vd backbone layout is supported")
        self.out_stages = [int(i) - 1 for i in _get(bc, "out_indices")]           # stage indices feeding the encoder
        self.d = int(_get(config, "d_model"))
        self.hid = int(_get(config, "encoder_hidden_dim"))
        self.strides = list(_get(config, "feat_strides"))
        self.enc_layers = int(_get(config, "encoder_layers"))
        self.enc_heads = int(_get(config, "encoder_attention_heads"))
        self.proj_layers = list(_get(config, "encode_proj_layers"))
        self.pe_temp = float(_get(config, "positional_encoding_temperature"))
        self.nq = int(_get(config, "num_queries"))
        self.dec_layers = int(_get(config, "decoder_layers"))
        self.dec_heads = int(_get(config, "decoder_attention_heads"))
        self.n_points = int(_get(config, "decoder_n_points"))
        self.n_levels = int(_get(config, "decoder_n_levels"))
        self.offset_scale = float(_get(config, "decoder_offset_scale"))
        self.ln_eps = float(_get(config, "layer_norm_eps"))
        self.bn_eps = float(_get(config, "batch_norm_eps"))
        self.nc = int(self.sd["class_embed.0.weight"].shape[0])
        ok = (self.d == 256 and self.hid == 256 and len(self.strides) == 3 and self.n_levels == 3 and
              _get(config, "num_feature_levels") == 3 and self.d // self.dec_heads == 32 and
              _get(config, "activation_function") == "silu" and _get(config, "encoder_activation_function") == "gelu" and
              _get(config, "decoder_activation_function") == "relu" and _get(config, "decoder_method") == "default" and
              not _get(config, "normalize_before") and not _get(config, "learn_initial_query") and
              float(_get(config, "hidden_expansion")) == 1.0 and self.enc_layers == 1 and
              _get(config, "anchor_image_size") is None and _get(config, "eval_size") is None)
        if not ok:
            raise ValueError("RtDetrB200: configuration outside the supported RT-DETRv2 layout")
        id2label = names if names is not None else (_get(config, "id2label") or {})
        self.names = {int(k): str(v) for k, v in id2label.items()}
        self.use_focal_loss = bool(_get(config, "use_focal_loss", True))
        self._w: Dict[str, tuple] = {}
        self._plans: Dict[Tuple[int, int], dict] = {}

    # ---- weights -------------------------------------------------------------------------------------------------
    def _fold(self, conv: str, bn: str, eps: float, avg_taps: bool = False):
        """conv (no bias) + BatchNorm -> (weight planes, bias); `avg_taps`: 1x1 conv behind AvgPool2d(2) = 2x2/s2 conv."""
        key = conv + ("/avg" if avg_taps else "")
        if key not in self._w:
            sd = self.sd
            w, b = fold_conv_bn(sd[conv + ".weight"], sd[bn + ".weight"], sd[bn + ".bias"], sd[bn + ".running_mean"],
                                sd[bn + ".running_var"], eps)
            if avg_taps:
                w = fold_avgpool2(w)
            self._w[key] = (P.conv_weight_to_planes(w, self.planes), P.pad_bias(b, w.shape[0]))
        return self._w[key]

    def _fold_repvgg(self, pre: str):
        """RepVGG block: conv3x3+BN and conv1x1+BN summed -> one 3x3 conv (the 1x1 goes into the centre tap)."""
        if pre not in self._w:
            sd, eps = self.sd, self.bn_eps
            parts = [fold_conv_bn(sd[f"{pre}.{n}.conv.weight"], sd[f"{pre}.{n}.norm.weight"], sd[f"{pre}.{n}.norm.bias"],
                                  sd[f"{pre}.{n}.norm.running_mean"], sd[f"{pre}.{n}.norm.running_var"], eps)
                     for n in ("conv1", "conv2")]
            w, b = fold_repvgg(parts[0], parts[1])
            self._w[pre] = (P.conv_weight_to_planes(w, self.planes), P.pad_bias(b, w.shape[0]))
        return self._w[pre]

    def _lin(self, name: str):
        if name not in self._w:
            w = self.sd[name + ".weight"]
            self._w[name] = (P.conv_weight_to_planes(w[:, :, None, None], self.planes), P.pad_bias(self.sd[name + ".bias"], w.shape[0]))
        return self._w[name]

    # ---- graph ---------------------------------------------------------------------------------------------------
    def _build(self, H: int, W: int) -> dict:
        pl, dev, hid, d = self.planes, self.device, self.hid, self.d
        keep: List[torch.Tensor] = []
        steps: List[tuple] = []

        def buf(h, w, c):
            t = torch.zeros((pl, 1, h, w, c), dtype=torch.bfloat16, device=dev)
            keep.append(t)
            return t

        def f32(rows, c):
            t = torch.zeros((1, 1, rows, c), dtype=torch.float32, device=dev)
            keep.append(t)
            return t

        def conv(x, wb, out, **kw):
            steps.append(("conv", ConvPlan(x, wb[0], wb[1], out, **kw)))
            return out

        g = dict(keep=keep)
        g["x_in"] = buf(H, W, 64)
        bb = "model.backbone.model"
        # stem: three 3x3 convs (the first with stride 2) and MaxPool2d(3, 2, 1)
        h, w = (H + 1) // 2, (W + 1) // 2
        e = self.embedding_size
        x = conv(g["x_in"], self._fold(f"{bb}.embedder.embedder.0.convolution", f"{bb}.embedder.embedder.0.normalization", 1e-5),
                 buf(h, w, e // 2), k=3, stride=2, pad=1, act="relu")
        x = conv(x, self._fold(f"{bb}.embedder.embedder.1.convolution", f"{bb}.embedder.embedder.1.normalization", 1e-5),
                 buf(h, w, e // 2), k=3, pad=1, act="relu")
        x = conv(x, self._fold(f"{bb}.embedder.embedder.2.convolution", f"{bb}.embedder.embedder.2.normalization", 1e-5),
                 buf(h, w, e), k=3, pad=1, act="relu")
        hp, wp = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        xp = buf(hp, wp, e)
        steps.append(("maxpool", (x, xp, h, w, e)))
        x, h, w, cin = xp, hp, wp, e
        feats = []
        for si, (depth, cout) in enumerate(zip(self.depths, self.hidden_sizes)):
            for bi in range(depth):
                s = 2 if (si > 0 and bi == 0) else 1
                pre = f"{bb}.encoder.stages.{si}.layers.{bi}"
                ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
                mid = cout // 4
                if bi == 0:
                    if s == 2:          # AvgPool2d(2, 2, ceil_mode=True) -> conv1x1 -> BN  ==  one 2x2 stride-2 conv
                        if h % 2 or w % 2:
                            raise ValueError("RtDetrB200: feature maps must stay even (imgsz a multiple of 32)")
                        sc = conv(x, self._fold(f"{pre}.shortcut.1.convolution", f"{pre}.shortcut.1.normalization", 1e-5, True),
                                  buf(ho, wo, cout), k=2, stride=2, pad=0)
                    else:
                        sc = conv(x, self._fold(f"{pre}.shortcut.convolution", f"{pre}.shortcut.normalization", 1e-5),
                                  buf(ho, wo, cout), k=1)
                else:
                    sc = x
                a = conv(x, self._fold(f"{pre}.layer.0.convolution", f"{pre}.layer.0.normalization", 1e-5), buf(h, w, mid),
                         k=1, act="relu")
                b = conv(a, self._fold(f"{pre}.layer.1.convolution", f"{pre}.layer.1.normalization", 1e-5), buf(ho, wo, mid),
                         k=3, stride=s, pad=1, act="relu")
                x = conv(b, self._fold(f"{pre}.layer.2.convolution", f"{pre}.layer.2.normalization", 1e-5), buf(ho, wo, cout),
                         k=1, act="relu", residual=sc, act_after_res=True)
                h, w, cin = ho, wo, cout
            if si in self.out_stages:
                feats.append((x, h, w, cout))
        # encoder input projections (conv1x1 + BN)
        proj = []
        for lv, (f, fh, fw, fc) in enumerate(feats):
            proj.append([conv(f, self._fold(f"model.encoder_input_proj.{lv}.0", f"model.encoder_input_proj.{lv}.1", 1e-5),
                              buf(fh, fw, hid), k=1), fh, fw])
        # AIFI on the listed levels
        for ai, lv in enumerate(self.proj_layers):
            t, th, tw = proj[lv]
            proj[lv][0] = self._aifi(steps, buf, f"model.encoder.aifi.{ai}.layers.0", t, th, tw, g)
        # CCFM: top-down FPN ...
        enc = "model.encoder"
        fpn = [proj[-1]]
        n_stage = len(proj) - 1
        for idx in range(n_stage):
            low, lh, lw = proj[n_stage - idx - 1]
            top, th, tw = fpn[-1]
            lat = conv(top, self._fold(f"{enc}.lateral_convs.{idx}.conv", f"{enc}.lateral_convs.{idx}.norm", self.bn_eps),
                       buf(th, tw, hid), k=1, act="silu")
            fpn[-1] = [lat, th, tw]
            cat = buf(lh, lw, 2 * hid)                         # [upsampled top | backbone level]
            steps.append(("up", (lat, cat, th, tw, hid, 2 * hid, 0)))
            steps.append(("copy_slice", (low, cat, hid)))
            fpn.append([self._csp(steps, buf, f"{enc}.fpn_blocks.{idx}", cat, lh, lw), lh, lw])
        fpn.reverse()
        # ... and bottom-up PAN
        pan = [fpn[0]]
        for idx in range(n_stage):
            top, th, tw = pan[-1]
            nxt, nh, nw = fpn[idx + 1]
            cat = buf(nh, nw, 2 * hid)                         # [downsampled | fpn level]
            steps.append(("conv", ConvPlan(top, *self._fold(f"{enc}.downsample_convs.{idx}.conv", f"{enc}.downsample_convs.{idx}.norm",
                                                            self.bn_eps), cat, k=3, stride=2, pad=1, act="silu", out_coff=0)))
            steps.append(("copy_slice", (nxt, cat, hid)))
            pan.append([self._csp(steps, buf, f"{enc}.pan_blocks.{idx}", cat, nh, nw), nh, nw])
        # decoder input projections, flattened level after level into one token tensor
        n_tok = sum(ph * pw for _, ph, pw in pan)
        tokens = buf(1, n_tok, d)
        start = 0
        levels = []
        for lv, (t, th, tw) in enumerate(pan):
            o = conv(t, self._fold(f"model.decoder_input_proj.{lv}.0", f"model.decoder_input_proj.{lv}.1", self.bn_eps),
                     buf(th, tw, d), k=1)
            steps.append(("copy_rows", (o, tokens, start, th * tw)))
            levels.append((th, tw, start))
            start += th * tw
        g.update(tokens=tokens, n_tok=n_tok, levels=levels)
        # anchors / valid mask (RTDetrV2Model.generate_anchors), host once per geometry
        anchors = []
        for lv, (th, tw, _) in enumerate(levels):
            gy, gx = torch.meshgrid(torch.arange(th, dtype=torch.float32), torch.arange(tw, dtype=torch.float32), indexing="ij")
            gxy = torch.stack([gx, gy], -1).unsqueeze(0) + 0.5
            gxy[..., 0] /= tw
            gxy[..., 1] /= th
            wh = torch.ones_like(gxy) * 0.05 * (2.0 ** lv)
            anchors.append(torch.concat([gxy, wh], -1).reshape(-1, th * tw, 4))
        anchors = torch.concat(anchors, 1)
        valid = ((anchors > 1e-2) * (anchors < 1 - 1e-2)).all(-1, keepdim=True)
        anchors = torch.log(anchors / (1 - anchors))
        anchors = torch.where(valid, anchors, torch.tensor(torch.finfo(torch.float32).max))
        g["anchors"] = anchors[0].to(dev)                                        # [n_tok][4]
        g["valid"] = valid[0].to(dev).to(torch.bfloat16).view(1, 1, 1, n_tok, 1)  # multiplies both planes exactly
        # encoder head on all tokens
        mem = buf(1, n_tok, d)
        steps.append(("mask_rows", (tokens, g["valid"], mem)))
        eo = conv(mem, self._lin("model.enc_output.0"), buf(1, n_tok, d), k=1)
        om = buf(1, n_tok, d)
        steps.append(("ln", (eo, n_tok, d, self.sd["model.enc_output.1.weight"], self.sd["model.enc_output.1.bias"], om)))
        ncp = P.pad_to(self.nc, 16)
        g["enc_cls"] = f32(n_tok, ncp)
        conv(om, self._lin("model.enc_score_head"), g["enc_cls"], k=1)
        b1 = conv(om, self._lin("model.enc_bbox_head.layers.0"), buf(1, n_tok, d), k=1, act="relu")
        b2 = conv(b1, self._lin("model.enc_bbox_head.layers.1"), buf(1, n_tok, d), k=1, act="relu")
        g["enc_box"] = f32(n_tok, 16)
        conv(b2, self._lin("model.enc_bbox_head.layers.2"), g["enc_box"], k=1)
        g["out_mem"] = om
        g["steps_encoder"] = steps
        # ---- decoder ------------------------------------------------------------------------------------------------
        Q = self.nq
        dsteps: List[List[tuple]] = []
        tgt = buf(1, Q, d)
        g["tgt0"] = tgt
        g["ref_in"] = torch.zeros((pl, 1, 1, Q, 64), dtype=torch.bfloat16, device=dev)      # sigmoid(ref) as planes (4 of 64 ch)
        g["ref32"] = torch.zeros((Q, 4), dtype=torch.float32, device=dev)
        lvl = (C.c_int * (3 * len(levels)))(*[v for t in levels for v in t])
        g["lvl"] = lvl
        heads = self.dec_heads
        LP = self.n_levels * self.n_points
        g["box_out"], g["cls_out"] = [], None
        x = tgt
        for li in range(self.dec_layers):
            st: List[tuple] = []
            pre = f"model.decoder.layers.{li}"

            def dconv(xx, wb, out, **kw):
                st.append(("conv", ConvPlan(xx, wb[0], wb[1], out, **kw)))
                return out

            # query position embedding from the current reference boxes
            qp1 = dconv(g["ref_in"], self._lin("model.decoder.query_pos_head.layers.0"), buf(1, Q, 2 * d), k=1, act="relu")
            qpos = dconv(qp1, self._lin("model.decoder.query_pos_head.layers.1"), buf(1, Q, d), k=1)
            # self attention: q = k = x + pos, v = x
            xq = buf(1, Q, d)
            st.append(("add", (x, qpos, xq, Q, d, Q)))
            qb = dconv(xq, self._lin(f"{pre}.self_attn.q_proj"), buf(1, Q, d), k=1)
            kb = dconv(xq, self._lin(f"{pre}.self_attn.k_proj"), buf(1, Q, d), k=1)
            vb = dconv(x, self._lin(f"{pre}.self_attn.v_proj"), buf(1, Q, d), k=1)
            ab = buf(1, Q, d)
            st.append(("attn", dict(q=qb, k=kb, v=vb, out=ab, heads=heads, hd=d // heads, nq=Q, nk=Q)))
            x1 = dconv(ab, self._lin(f"{pre}.self_attn.o_proj"), buf(1, Q, d), k=1, residual=x)
            x1n = buf(1, Q, d)
            st.append(("ln", (x1, Q, d, self.sd[f"{pre}.self_attn_layer_norm.weight"], self.sd[f"{pre}.self_attn_layer_norm.bias"], x1n)))
            # multi-scale deformable attention over the encoder tokens
            xq2 = buf(1, Q, d)
            st.append(("add", (x1n, qpos, xq2, Q, d, Q)))
            val = dconv(tokens, self._lin(f"{pre}.encoder_attn.value_proj"), buf(1, n_tok, d), k=1)
            off = f32(Q, heads * LP * 2)
            dconv(xq2, self._lin(f"{pre}.encoder_attn.sampling_offsets"), off, k=1)
            lg = f32(Q, P.pad_to(heads * LP, 16))
            dconv(xq2, self._lin(f"{pre}.encoder_attn.attention_weights"), lg, k=1)
            da = buf(1, Q, d)
            st.append(("deform", (val, off, lg, da, n_tok)))
            x2 = dconv(da, self._lin(f"{pre}.encoder_attn.output_proj"), buf(1, Q, d), k=1, residual=x1n)
            x2n = buf(1, Q, d)
            st.append(("ln", (x2, Q, d, self.sd[f"{pre}.encoder_attn_layer_norm.weight"], self.sd[f"{pre}.encoder_attn_layer_norm.bias"], x2n)))
            # FFN
            hm = dconv(x2n, self._lin(f"{pre}.mlp.fc1"), buf(1, Q, self.sd[f"{pre}.mlp.fc1.weight"].shape[0]), k=1, act="relu")
            x3 = dconv(hm, self._lin(f"{pre}.mlp.fc2"), buf(1, Q, d), k=1, residual=x2n)
            x3n = buf(1, Q, d)
            st.append(("ln", (x3, Q, d, self.sd[f"{pre}.final_layer_norm.weight"], self.sd[f"{pre}.final_layer_norm.bias"], x3n)))
            # box refinement head (applied to the reference boxes by the caller), class head on the last layer
            bb1 = dconv(x3n, self._lin(f"model.decoder.bbox_embed.{li}.layers.0"), buf(1, Q, d), k=1, act="relu")
            bb2 = dconv(bb1, self._lin(f"model.decoder.bbox_embed.{li}.layers.1"), buf(1, Q, d), k=1, act="relu")
            bo = f32(Q, 16)
            dconv(bb2, self._lin(f"model.decoder.bbox_embed.{li}.layers.2"), bo, k=1)
            g["box_out"].append(bo)
            if li == self.dec_layers - 1:
                g["cls_out"] = f32(Q, ncp)
                dconv(x3n, self._lin(f"model.decoder.class_embed.{li}"), g["cls_out"], k=1)
            dsteps.append(st)
            x = x3n
        g["steps_decoder"] = dsteps
        g["in_u8"] = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        return g

    def _aifi(self, steps, buf, pre, t, th, tw, g):
        """RTDetrV2AIFILayer with one post-norm encoder layer on a th x tw map (tokens = rows of the same buffer)."""
        hid, n = self.hid, th * tw
        heads = self.enc_heads
        # 2D sin-cos position embedding (RTDetrV2SinePositionEmbedding), constant per geometry
        gw, gh = torch.meshgrid(torch.arange(tw, dtype=torch.float32), torch.arange(th, dtype=torch.float32), indexing="xy")
        pos_dim = hid // 4
        omega = 1.0 / (self.pe_temp ** (torch.arange(pos_dim, dtype=torch.float32) / pos_dim))
        ow, oh = gw.flatten()[..., None] @ omega[None], gh.flatten()[..., None] @ omega[None]
        pe = torch.concat([oh.sin(), oh.cos(), ow.sin(), ow.cos()], dim=1)                      # [n][hid]
        pos = P.split_planes(pe.view(1, th, tw, hid).to(self.device), self.planes)
        g["keep"].append(pos)
        xq = buf(th, tw, hid)
        steps.append(("add", (t, pos, xq, n, hid, n)))
        lin = self._lin
        qb, kb, vb, ab = buf(th, tw, hid), buf(th, tw, hid), buf(th, tw, hid), buf(th, tw, hid)
        for name, src, dst in (("q_proj", xq, qb), ("k_proj", xq, kb), ("v_proj", t, vb)):
            steps.append(("conv", ConvPlan(src, *lin(f"{pre}.self_attn.{name}"), dst, k=1)))
        steps.append(("attn", dict(q=qb, k=kb, v=vb, out=ab, heads=heads, hd=hid // heads, nq=n, nk=n)))
        x1 = buf(th, tw, hid)
        steps.append(("conv", ConvPlan(ab, *lin(f"{pre}.self_attn.o_proj"), x1, k=1, residual=t)))
        x1n = buf(th, tw, hid)
        steps.append(("ln", (x1, n, hid, self.sd[f"{pre}.self_attn_layer_norm.weight"], self.sd[f"{pre}.self_attn_layer_norm.bias"], x1n)))
        hm = buf(th, tw, self.sd[f"{pre}.mlp.fc1.weight"].shape[0])
        steps.append(("conv", ConvPlan(x1n, *lin(f"{pre}.mlp.fc1"), hm, k=1, act="gelu")))
        x2 = buf(th, tw, hid)
        steps.append(("conv", ConvPlan(hm, *lin(f"{pre}.mlp.fc2"), x2, k=1, residual=x1n)))
        x2n = buf(th, tw, hid)
        steps.append(("ln", (x2, n, hid, self.sd[f"{pre}.final_layer_norm.weight"], self.sd[f"{pre}.final_layer_norm.bias"], x2n)))
        return x2n

    def _csp(self, steps, buf, pre, cat, h, w):
        """RTDetrV2CSPRepLayer (hidden_expansion 1.0: conv3 is the identity): silu(conv1(x)) -> 3 RepVGG blocks, plus
        silu(conv2(x)) added in the last block's epilogue."""
        hid, eps = self.hid, self.bn_eps
        a = buf(h, w, hid)
        steps.append(("conv", ConvPlan(cat, *self._fold(f"{pre}.conv1.conv", f"{pre}.conv1.norm", eps), a, k=1, act="silu")))
        c2 = buf(h, w, hid)
        steps.append(("conv", ConvPlan(cat, *self._fold(f"{pre}.conv2.conv", f"{pre}.conv2.norm", eps), c2, k=1, act="silu")))
        x = a
        for bi in range(3):
            o = buf(h, w, hid)
            kw = dict(residual=c2) if bi == 2 else {}
            steps.append(("conv", ConvPlan(x, *self._fold_repvgg(f"{pre}.bottlenecks.{bi}"), o, k=3, pad=1, act="silu", **kw)))
            x = o
        return x

    # ---- execution -----------------------------------------------------------------------------------------------
    def _run(self, steps) -> None:
        l, st, pl = self.l, stream_ptr(), self.planes
        for kind, a in steps:
            if kind == "conv":
                a.run()
            elif kind == "ln":
                x, rows, c, gm, bt, out = a
                check(l.mtb_layernorm(ptr(x), rows, c, c, 0, pl, ptr(gm), ptr(bt), self.ln_eps, ptr(out), c, 0, pl, 0, st),
                      "mtb_layernorm")
            elif kind == "add":
                x, y, o, rows, c, brows = a
                check(l.mtb_add_planes(ptr(x), ptr(y), ptr(o), rows, c, brows, pl, st), "mtb_add_planes")
            elif kind == "maxpool":
                x, y, h, w, c = a
                check(l.mtb_maxpool2d(ptr(x), ptr(y), 1, h, w, c, 3, 2, 1, pl, st), "mtb_maxpool2d")
            elif kind == "up":
                src, dst, h, w, c, ct_out, co = a
                check(l.mtb_upsample2x(ptr(src), ptr(dst), 1, h, w, c, 0, ct_out, co, c, pl, st), "mtb_upsample2x")
            elif kind == "copy_slice":
                src, dst, co = a
                dst[..., co:co + src.shape[-1]].copy_(src)
            elif kind == "copy_rows":
                src, dst, start, n = a
                dst[:, 0, 0, start:start + n].copy_(src.view(pl, n, -1))
            elif kind == "mask_rows":
                src, mask, dst = a
                torch.mul(src, mask, out=dst)
            elif kind == "attn":
                self._attn(a)
            elif kind == "deform":
                val, off, lg, out, n_tok = a
                g = self._cur
                check(l.mtb_deform_attn(ptr(val), val[0].numel(), pl, self.d, self.dec_heads, self.d // self.dec_heads,
                                        self.n_levels, g["lvl"], self.n_points, ptr(off), off.shape[-1], ptr(lg), lg.shape[-1],
                                        ptr(g["ref32"]), self.nq, self.offset_scale, ptr(out), out[0].numel(), st),
                      "mtb_deform_attn")

    def _attn(self, a: dict) -> None:
        dsc = AttnDesc()
        q, k, v, o = a["q"], a["k"], a["v"], a["out"]
        dsc.heads, dsc.hd = a["heads"], a["hd"]
        dsc.scale = float(a["hd"]) ** -0.5
        dsc.q, dsc.k, dsc.v, dsc.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
        dsc.q_ct, dsc.k_ct, dsc.v_ct, dsc.o_ct = q.shape[-1], k.shape[-1], v.shape[-1], o.shape[-1]
        dsc.q_ps, dsc.k_ps, dsc.v_ps, dsc.o_ps = q[0].numel(), k[0].numel(), v[0].numel(), o[0].numel()
        dsc.planes = self.planes
        dsc.mode, dsc.B, dsc.nq, dsc.nk = 0, 1, a["nq"], a["nk"]
        check(self.l.mtb_attention(C.byref(dsc), stream_ptr()), "mtb_attention")

    def _plan(self, H: int, W: int) -> dict:
        if (H, W) not in self._plans:
            self._plans[(H, W)] = self._build(H, W)
        return self._plans[(H, W)]

    def _forward_static(self, g: dict, debug: Optional[dict] = None) -> None:
        """g['in_u8'] -> g['logits'], g['boxes']; device work on static buffers only (no host sync), graph-capturable."""
        self._cur = g
        l, st = self.l, stream_ptr()
        H, W = int(g["in_u8"].shape[0]), int(g["in_u8"].shape[1])
        zero = (C.c_float * 3)(0.0, 0.0, 0.0)
        check(l.mtb_image_to_planes(ptr(g["in_u8"]), H, W, 3, 0, 1.0 / 255.0, zero, ptr(g["x_in"]), 64, self.planes, st),
              "mtb_image_to_planes")
        self._run(g["steps_encoder"])
        # query selection: top-Q tokens by their best class score
        enc_cls = g["enc_cls"][0, 0, :, :self.nc]
        enc_box = g["enc_box"][0, 0, :, :4] + g["anchors"]
        _, topk = torch.topk(enc_cls.max(-1).values, self.nq, dim=0)
        ref_unact = enc_box[topk]
        g["tgt0"].copy_(g["out_mem"][:, :, :, topk])
        ref = torch.sigmoid(ref_unact)
        if debug is not None:
            debug.update(enc_cls=enc_cls.clone(), enc_box=enc_box.clone(), topk=topk.clone(),
                         tokens=P.merge_planes(g["tokens"])[0, 0].clone())
        for li, st_l in enumerate(g["steps_decoder"]):
            g["ref32"].copy_(ref)
            g["ref_in"][..., :4].copy_(P.split_planes(ref.view(1, 1, self.nq, 4), self.planes))
            self._run(st_l)
            pred = g["box_out"][li][0, 0, :, :4]
            x = ref.clamp(0, 1)
            inv = torch.log(x.clamp(min=1e-5) / (1 - x).clamp(min=1e-5))
            ref = torch.sigmoid(pred + inv)
        g["logits"].copy_(g["cls_out"][0, 0, :, :self.nc])
        g["boxes"].copy_(ref)

    def forward_u8(self, img_rgb_u8: torch.Tensor, *, debug: Optional[dict] = None):
        """img_rgb_u8: device uint8 HxWx3 already at the network size.  Returns (logits [Q][nc], boxes cxcywh [Q][4]) —
        static output buffers of the plan, valid until the next call.  From the second call of a size on, the ~450
        launches (conv plans, glue kernels and the few torch index ops) replay as one CUDA graph."""
        from . import graphs
        H, W = int(img_rgb_u8.shape[0]), int(img_rgb_u8.shape[1])
        g = self._plan(H, W)
        if "logits" not in g:
            g["logits"] = torch.zeros((self.nq, self.nc), dtype=torch.float32, device=self.device)
            g["boxes"] = torch.zeros((self.nq, 4), dtype=torch.float32, device=self.device)
            g["uses"] = 0
        g["in_u8"].copy_(img_rgb_u8[:, :, :3])
        g["uses"] += 1
        if debug is None and graphs.ENABLED and ("graph" in g or g["uses"] >= 2):
            if "graph" not in g:
                g["graph"] = graphs.CapturedGraph(lambda: self._forward_static(g))
            g["graph"].replay()
        else:
            self._forward_static(g, debug)
        return g["logits"], g["boxes"]

    def predict(self, img_rgb_u8: torch.Tensor, conf: float, imgsz: Optional[int]):
        """Processor + model + `post_process_object_detection` for one device uint8 HxWx3 RGB image.  Returns
        (xyxy [n][4] original pixels, scores [n], labels [n]) in descending score order, all on the device."""
        oh, ow = int(img_rgb_u8.shape[0]), int(img_rgb_u8.shape[1])
        size = int(imgsz) if imgsz is not None else 640
        x = img_rgb_u8 if (oh, ow) == (size, size) else resize_aa_device(img_rgb_u8.contiguous(), size, size)
        logits, boxes = self.forward_u8(x)
        cx, cy, w, h = boxes.unbind(-1)
        xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
        xyxy = xyxy * torch.tensor([ow, oh, ow, oh], dtype=torch.float32, device=self.device)
        if self.use_focal_loss:
            scores = torch.sigmoid(logits)
            scores, index = torch.topk(scores.flatten(), self.nq)
            labels = index % self.nc
            xyxy = xyxy[index // self.nc]
        else:
            scores, labels = torch.softmax(logits, -1)[:, :-1].max(-1)
        keep = scores > float(conf)
        return xyxy[keep], scores[keep], labels[keep]

    def __call__(self, source, conf: float = 0.35, device=None, verbose: bool = False, imgsz: Optional[int] = None, **_kw):
        """YOLO-style call of the reference's adapter (core/ml/rtdetr_adapter.py:61-113)."""
        from PIL import Image
        if isinstance(source, torch.Tensor) and source.is_cuda:
            rgb = source[:, :, [2, 1, 0]].contiguous()                   # device BGR page
        else:
            if isinstance(source, str):
                source = Image.open(source)
            if isinstance(source, Image.Image):
                arr = np.asarray(source.convert("RGB") if source.mode != "RGB" else source)
            else:
                arr = np.asarray(source)
                arr = np.repeat(arr[:, :, None], 3, 2) if arr.ndim == 2 else arr[:, :, [2, 1, 0]]   # BGR(A) -> RGB
            rgb = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        xyxy, scores, labels = self.predict(rgb, conf, imgsz)
        return [SimpleNamespace(boxes=_Boxes(xyxy.float(), scores.float(), labels.float()), names=self.names)]
