"""B200 SAM 2.1 (Hiera encoder + FPN neck + box-prompted mask decoder) — what `ModelManager.load_sam2()` returns.

Mirrors the computation the reference obtains from `transformers.Sam2Model` / `Sam2Processor`
(core/ml/model_manager.py:982-1010, core/image/detection.py:475-511): resize to 1024x1024 (uint8 antialias bilinear),
normalise, Hiera image encoder, neck, prompt encoder (boxes), two-way-transformer mask decoder, dynamic multimask
selection, bilinear upsampling to the page, `> 0`, clip to the floor/ceil box, uint8 {0,255} masks.

Every linear / 1x1 conv / transposed conv is a tcgen05 conv plan (bf16x3 by default so the float path stays within
1e-3 of the fp32 oracle); layer norms, window attention (partition, padding and query pooling folded into the kernel's
addressing), the hyper-network mask product and the mask writer are the kernels in csrc/sam_kernels.cu.
"""
from __future__ import annotations

import collections
import os
import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import planes as P
from ._lib import check, lib, ptr, stream_ptr
from .ops import ConvPlan
from .preproc import resize_aa_device

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class AttnDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("heads", C.c_int), ("hd", C.c_int), ("nq", C.c_int), ("nk", C.c_int),
                ("scale", C.c_float),
                ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
                ("q_ct", C.c_int), ("q_off", C.c_int), ("k_ct", C.c_int), ("k_off", C.c_int),
                ("v_ct", C.c_int), ("v_off", C.c_int), ("o_ct", C.c_int), ("o_off", C.c_int),
                ("q_ps", C.c_longlong), ("k_ps", C.c_longlong), ("v_ps", C.c_longlong), ("o_ps", C.c_longlong),
                ("planes", C.c_int), ("mode", C.c_int),
                ("grid_h", C.c_int), ("grid_w", C.c_int), ("ws", C.c_int), ("pool", C.c_int),
                ("pad_q", C.c_void_p), ("pad_k", C.c_void_p), ("pad_v", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong)]


def _declare(l) -> None:
    if getattr(l, "_sam_declared", False):
        return
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    l.mtb_layernorm.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp, f32, vp, i32, i32, i32, i32, vp]
    l.mtb_maxpool2x2.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    l.mtb_add_planes.argtypes = [vp, vp, vp, i64, i32, i64, i32, vp]
    l.mtb_attention.argtypes = [C.POINTER(AttnDesc), vp]
    l.mtb_attention_workspace_bytes.argtypes = [i32, i32, i32, i32]
    l.mtb_attention_workspace_bytes.restype = i64
    l.mtb_sam_prompt_boxes.argtypes = [vp, f32, f32, i32, vp, i32, vp, vp, vp, f32, vp, vp]
    l.mtb_sam_hyper_masks.argtypes = [vp, i32, vp, i32, i32, i32, i64, vp, vp]
    l.mtb_sam_select_mask.argtypes = [vp, vp, i32, i32, i64, f32, f32, vp, vp]
    l.mtb_sam_mask_write.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, vp, vp, vp]
    l.mtb_sam_patch_embed.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]
    l.mtb_upsample2x.argtypes = [vp, vp] + [i32] * 9 + [vp]
    for n in ("mtb_layernorm", "mtb_maxpool2x2", "mtb_add_planes", "mtb_attention", "mtb_sam_prompt_boxes",
              "mtb_sam_hyper_masks", "mtb_sam_select_mask", "mtb_sam_mask_write", "mtb_sam_patch_embed", "mtb_upsample2x"):
        getattr(l, n).restype = i32
    l._sam_declared = True


def _cfg_get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


class Sam2B200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config, device: torch.device, *, precision: str = "bf16x3"):
        self.l = lib()
        _declare(self.l)
        self.device = device
        self.planes = 2 if precision == "bf16x3" else 1
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self.sd = sd
        vc = _cfg_get(config, "vision_config")
        bc = _cfg_get(vc, "backbone_config")
        self.dims = list(_cfg_get(bc, "embed_dim_per_stage"))
        self.heads = list(_cfg_get(bc, "num_attention_heads_per_stage"))
        self.blocks_per_stage = list(_cfg_get(bc, "blocks_per_stage"))
        self.windows = list(_cfg_get(bc, "window_size_per_stage"))
        self.global_blocks = set(_cfg_get(bc, "global_attention_blocks"))
        self.pool_stages = int(_cfg_get(bc, "num_query_pool_stages"))
        self.ln_eps = float(_cfg_get(bc, "layer_norm_eps", 1e-6))
        self.img = 1024
        md = _cfg_get(config, "mask_decoder_config")
        self.dec_heads = int(_cfg_get(md, "num_attention_heads", 8))
        self.dec_layers = int(_cfg_get(md, "num_hidden_layers", 2))
        self.stab_delta = float(_cfg_get(md, "dynamic_multimask_stability_delta", 0.05))
        self.stab_thresh = float(_cfg_get(md, "dynamic_multimask_stability_thresh", 0.98))
        self.hidden = 256
        self._w: Dict[str, tuple] = {}
        self._enc = None
        # decoder plans per prompt count (buffers + launch descriptors + one CUDA graph and mask buffer per page size),
        # least recently used first: pages with ragged bubble counts / sizes must not grow device memory without bound
        self._dec: "collections.OrderedDict[int, dict]" = collections.OrderedDict()
        self.max_decoders = int(os.environ.get("MTB200_SAM_MAX_DECODERS", "8"))
        self.max_sizes_per_decoder = int(os.environ.get("MTB200_SAM_MAX_SIZES", "4"))
        self._prep_constants()

    # ---- weights / constants -----------------------------------------------------------------------------------
    def _lin(self, name: str, extra_bias: Optional[torch.Tensor] = None):
        key = name + ("+" if extra_bias is not None else "")
        if key not in self._w:
            w = self.sd[name + ".weight"]
            if w.dim() == 2:
                w = w[:, :, None, None]
            b = self.sd[name + ".bias"]
            if extra_bias is not None:
                b = b + extra_bias
            self._w[key] = (P.conv_weight_to_planes(w, self.planes), P.pad_bias(b, w.shape[0]))
        return self._w[key]

    def _deconv(self, name: str):
        if name not in self._w:
            w = self.sd[name + ".weight"]            # [Cin][Cout][2][2]
            cin, cout = w.shape[0], w.shape[1]
            w1 = w.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1).contiguous()
            self._w[name] = (P.conv_weight_to_planes(w1, self.planes), P.pad_bias(self.sd[name + ".bias"].repeat(4), 4 * cout))
        return self._w[name]

    def _prep_constants(self) -> None:
        sd, dev = self.sd, self.device
        # window + background positional embedding of the patch grid (weights only; Sam2HieraDetModel._get_pos_embed)
        g = self.img // 4
        pe = F.interpolate(sd["vision_encoder.backbone.pos_embed"].cpu(), size=(g, g), mode="bicubic")
        we = sd["vision_encoder.backbone.pos_embed_window"].cpu()
        pe = pe + we.tile([x // y for x, y in zip(pe.shape, we.shape)])
        self.pos_embed = pe.permute(0, 2, 3, 1).contiguous().to(dev)             # [1][g][g][C0] fp32
        # rescale+normalize constants exactly as the HF fast image processor fuses them
        mean = torch.tensor(IMAGENET_MEAN) * (1.0 / (1.0 / 255.0))
        std = torch.tensor(IMAGENET_STD) * (1.0 / (1.0 / 255.0))
        self.norm_mean, self.norm_std = mean.to(dev), std.to(dev)
        # image-wide positional embedding of the 64x64 embedding grid (Sam2Model.get_image_wide_positional_embeddings)
        s = self.img // 16
        grid = torch.ones((s, s))
        y = (grid.cumsum(0) - 0.5) / s
        x = (grid.cumsum(1) - 0.5) / s
        coords = 2 * torch.stack([x, y], -1) - 1
        gauss = sd["shared_image_embedding.positional_embedding"].cpu()
        proj = 2 * np.pi * (coords @ gauss)
        kpe = torch.cat([proj.sin(), proj.cos()], -1).reshape(1, 1, s * s, self.hidden)
        self.key_pe = P.split_planes(kpe.to(dev), self.planes)                      # [pl][1][1][4096][256]
        self.out_tokens = torch.cat([sd["mask_decoder.obj_score_token.weight"], sd["mask_decoder.iou_token.weight"],
                                     sd["mask_decoder.mask_tokens.weight"]], 0).contiguous()  # [6][256]
        self.gauss_prompt = sd["prompt_encoder.shared_embedding.positional_embedding"].contiguous()
        pe_w = sd["prompt_encoder.point_embed.weight"]
        self.pe2, self.pe3 = pe_w[2].contiguous(), pe_w[3].contiguous()
        self.not_a_point = sd["prompt_encoder.not_a_point_embed.weight"][0].contiguous()

    # ---- small launch helpers ----------------------------------------------------------------------------------
    def _ln(self, x, name, out, gelu=False, eps=None):
        g, b = self.sd[name + ".weight"], self.sd[name + ".bias"]
        rows = x.shape[1] * x.shape[2] * x.shape[3]
        c = x.shape[4]
        return ("ln", (x, rows, c, g, b, out, int(gelu), float(self.ln_eps if eps is None else eps)))

    def _buf(self, n, h, w, c, keep):
        t = torch.zeros((self.planes, n, h, w, c), dtype=torch.bfloat16, device=self.device)
        keep.append(t)
        return t

    # ---- encoder graph -----------------------------------------------------------------------------------------
    def _build_encoder(self) -> dict:
        keep: List[torch.Tensor] = []
        steps: List[tuple] = []
        g = self.img // 4
        x = self._buf(1, g, g, self.dims[0], keep)
        enc = dict(x0=x)
        stage_out = []
        bi = 0
        H = g
        for s, nb in enumerate(self.blocks_per_stage):
            for b in range(nb):
                pre = f"vision_encoder.backbone.blocks.{bi}"
                first = s > 0 and b == 0
                d_in = self.dims[s - 1] if first else self.dims[s]
                d_out = self.dims[s]
                ws = self.windows[s - 1] if first else self.windows[s]
                if bi in self.global_blocks:
                    ws = 0
                pool = 0 < s <= self.pool_stages and b == 0
                heads = self.heads[s]
                hd = d_out // heads
                Ho = H // 2 if pool else H
                xn = self._buf(1, H, H, d_in, keep)
                steps.append(self._ln(x, pre + ".layer_norm1", xn))
                if d_in != d_out:
                    rfull = self._buf(1, H, H, d_out, keep)
                    w = self._lin(pre + ".proj")
                    steps.append(("conv", ConvPlan(xn, w[0], w[1], rfull, k=1)))
                    resid = self._buf(1, Ho, Ho, d_out, keep)
                    steps.append(("pool", (rfull, resid, H, d_out)))
                else:
                    resid = x
                qkv = self._buf(1, H, H, 3 * d_out, keep)
                w = self._lin(pre + ".attn.qkv")
                steps.append(("conv", ConvPlan(xn, w[0], w[1], qkv, k=1)))
                att = self._buf(1, Ho, Ho, d_out, keep)
                bq = self.sd[pre + ".attn.qkv.bias"]
                pads = (bq[:d_out].contiguous(), bq[d_out:2 * d_out].contiguous(), bq[2 * d_out:].contiguous())
                keep.extend(pads)
                steps.append(("attn", dict(q=qkv, k=qkv, v=qkv, out=att, q_off=0, k_off=d_out, v_off=2 * d_out, o_off=0,
                                           heads=heads, hd=hd, grid=H, ws=ws, pool=int(pool), pads=pads)))
                x1 = self._buf(1, Ho, Ho, d_out, keep)
                w = self._lin(pre + ".attn.proj")
                steps.append(("conv", ConvPlan(att, w[0], w[1], x1, k=1, residual=resid)))
                xn2 = self._buf(1, Ho, Ho, d_out, keep)
                steps.append(self._ln(x1, pre + ".layer_norm2", xn2))
                hm = self._buf(1, Ho, Ho, self.sd[pre + ".mlp.proj_in.weight"].shape[0], keep)
                w = self._lin(pre + ".mlp.proj_in")
                steps.append(("conv", ConvPlan(xn2, w[0], w[1], hm, k=1, act="gelu")))
                x2 = self._buf(1, Ho, Ho, d_out, keep)
                w = self._lin(pre + ".mlp.proj_out")
                steps.append(("conv", ConvPlan(hm, w[0], w[1], x2, k=1, residual=x1)))
                x, H = x2, Ho
                bi += 1
            stage_out.append((x, H))
        # neck (FPN): convs[0] <-> last stage; top-down add only into level 2 (fpn_top_down_levels = [2, 3])
        (f0, h0), (f1, h1), (f2, h2), (f3, h3) = stage_out
        hid = self.hidden
        lat3 = self._buf(1, h3, h3, hid, keep)
        w = self._lin("vision_encoder.neck.convs.0")
        steps.append(("conv", ConvPlan(f3, w[0], w[1], lat3, k=1)))
        up3 = self._buf(1, h2, h2, hid, keep)
        steps.append(("up", (lat3, up3, h3, hid)))
        emb = self._buf(1, h2, h2, hid, keep)
        extra = self.sd["no_memory_embedding"].reshape(-1) + self.sd["prompt_encoder.no_mask_embed.weight"].reshape(-1)
        w = self._lin("vision_encoder.neck.convs.1", extra_bias=extra)   # + no_memory_embedding + dense no-mask prompt
        steps.append(("conv", ConvPlan(f2, w[0], w[1], emb, k=1, residual=up3)))
        lat1 = self._buf(1, h1, h1, hid, keep)
        w = self._lin("vision_encoder.neck.convs.2")
        steps.append(("conv", ConvPlan(f1, w[0], w[1], lat1, k=1)))
        lat0 = self._buf(1, h0, h0, hid, keep)
        w = self._lin("vision_encoder.neck.convs.3")
        steps.append(("conv", ConvPlan(f0, w[0], w[1], lat0, k=1)))
        s1 = self._buf(1, h1, h1, 64, keep)
        w = self._lin("mask_decoder.conv_s1")
        steps.append(("conv", ConvPlan(lat1, w[0], w[1], s1, k=1)))
        s0 = self._buf(1, h0, h0, 32, keep)
        w = self._lin("mask_decoder.conv_s0")
        steps.append(("conv", ConvPlan(lat0, w[0], w[1], s0, k=1)))
        enc.update(steps=steps, keep=keep, emb=emb, s0=s0, s1=s1, emb_hw=h2)
        return enc

    # ---- decoder graph for P boxes ------------------------------------------------------------------------------
    def _build_decoder(self, Pn: int, enc: dict) -> dict:
        keep: List[torch.Tensor] = []
        steps: List[tuple] = []
        hid, pl, dev = self.hidden, self.planes, self.device
        T = 9
        S = enc["emb_hw"]
        nk = S * S
        tok = lambda c: self._buf(1, 1, Pn * T, c, keep)            # token tensors: rows = p*9 + t
        img = lambda c: self._buf(Pn, S, S, c, keep)                # per-box image tokens
        pe_q = tok(hid)                                             # = initial point embeddings
        queries = tok(hid)
        keys = img(hid)
        d = dict(pe_q=pe_q, queries0=queries, keys0=keys, T=T, S=S)

        def attn_block(prefix, q_src, k_src, v_src, out_rows_like_q, nq, nkeys, internal):
            """Sam2Attention: q/k/v projections, attention, output projection (returned tensor = o_proj output)."""
            heads = self.dec_heads
            hd = internal // heads
            q_is_img = nq == nk
            k_is_img = nkeys == nk
            qb = img(internal) if q_is_img else tok(internal)
            kb = img(internal) if k_is_img else tok(internal)
            vb = img(internal) if k_is_img else tok(internal)
            for name, src, dst in (("q_proj", q_src, qb), ("k_proj", k_src, kb), ("v_proj", v_src, vb)):
                w = self._lin(f"{prefix}.{name}")
                steps.append(("conv", ConvPlan(src, w[0], w[1], dst, k=1)))
            ob = img(internal) if q_is_img else tok(internal)
            steps.append(("attn", dict(q=qb, k=kb, v=vb, out=ob, q_off=0, k_off=0, v_off=0, o_off=0, heads=heads, hd=hd,
                                       B=Pn, nq=nq, nk=nkeys, ws=-1)))
            return ob

        def o_proj(prefix, ob, dst, residual):
            w = self._lin(f"{prefix}.o_proj")
            steps.append(("conv", ConvPlan(ob, w[0], w[1], dst, k=1, residual=residual)))

        q_cur, k_cur = queries, keys
        for li in range(self.dec_layers):
            pre = f"mask_decoder.transformer.layers.{li}"
            # (1) self attention on the tokens
            if li == 0:
                ob = attn_block(pre + ".self_attn", q_cur, q_cur, q_cur, True, T, T, hid)
                q1 = tok(hid)
                o_proj(pre + ".self_attn", ob, q1, None)             # first layer: queries are REPLACED
            else:
                qpe = tok(hid)
                steps.append(("add", (q_cur, pe_q, qpe, Pn * T, hid, Pn * T)))
                ob = attn_block(pre + ".self_attn", qpe, qpe, q_cur, True, T, T, hid)
                q1 = tok(hid)
                o_proj(pre + ".self_attn", ob, q1, q_cur)
            q1n = tok(hid)
            steps.append(self._ln(q1, pre + ".layer_norm1", q1n, eps=1e-5))
            # (2) cross attention tokens -> image
            qpe = tok(hid)
            steps.append(("add", (q1n, pe_q, qpe, Pn * T, hid, Pn * T)))
            kpe = img(hid)
            steps.append(("add", (k_cur, self.key_pe, kpe, Pn * nk, hid, nk)))
            ob = attn_block(pre + ".cross_attn_token_to_image", qpe, kpe, k_cur, True, T, nk, hid // 2)
            q2 = tok(hid)
            o_proj(pre + ".cross_attn_token_to_image", ob, q2, q1n)
            q2n = tok(hid)
            steps.append(self._ln(q2, pre + ".layer_norm2", q2n, eps=1e-5))
            # (3) MLP (ReLU)
            hm = tok(self.sd[pre + ".mlp.proj_in.weight"].shape[0])
            w = self._lin(pre + ".mlp.proj_in")
            steps.append(("conv", ConvPlan(q2n, w[0], w[1], hm, k=1, act="relu")))
            q3 = tok(hid)
            w = self._lin(pre + ".mlp.proj_out")
            steps.append(("conv", ConvPlan(hm, w[0], w[1], q3, k=1, residual=q2n)))
            q3n = tok(hid)
            steps.append(self._ln(q3, pre + ".layer_norm3", q3n, eps=1e-5))
            # (4) cross attention image -> tokens (updates the image tokens)
            qpe2 = tok(hid)
            steps.append(("add", (q3n, pe_q, qpe2, Pn * T, hid, Pn * T)))
            ob = attn_block(pre + ".cross_attn_image_to_token", kpe, qpe2, q3n, False, nk, T, hid // 2)
            k1 = img(hid)
            o_proj(pre + ".cross_attn_image_to_token", ob, k1, k_cur)
            k1n = img(hid)
            steps.append(self._ln(k1, pre + ".layer_norm4", k1n, eps=1e-5))
            q_cur, k_cur = q3n, k1n
        # final token -> image attention
        pre = "mask_decoder.transformer"
        qpe = tok(hid)
        steps.append(("add", (q_cur, pe_q, qpe, Pn * T, hid, Pn * T)))
        kpe = img(hid)
        steps.append(("add", (k_cur, self.key_pe, kpe, Pn * nk, hid, nk)))
        ob = attn_block(pre + ".final_attn_token_to_image", qpe, kpe, k_cur, True, T, nk, hid // 2)
        qf = tok(hid)
        o_proj(pre + ".final_attn_token_to_image", ob, qf, q_cur)
        qfn = tok(hid)
        steps.append(self._ln(qf, pre + ".layer_norm_final_attn", qfn, eps=1e-5))
        # upscaling: ConvTranspose2d(2,2) == 1x1 conv to 4*C + pixel-shuffle store; high-res skips are shared by all boxes
        up1 = self._buf(Pn, 2 * S, 2 * S, hid // 4, keep)
        w = self._deconv("mask_decoder.upscale_conv1")
        steps.append(("conv", ConvPlan(k_cur, w[0], w[1], up1, k=1, pixel_shuffle=True, residual=enc["s1"], res_bcast=True)))
        up1n = self._buf(Pn, 2 * S, 2 * S, hid // 4, keep)
        steps.append(self._ln(up1, "mask_decoder.upscale_layer_norm", up1n, gelu=True, eps=1e-6))
        up2 = self._buf(Pn, 4 * S, 4 * S, hid // 8, keep)
        w = self._deconv("mask_decoder.upscale_conv2")
        steps.append(("conv", ConvPlan(up1n, w[0], w[1], up2, k=1, pixel_shuffle=True, residual=enc["s0"], res_bcast=True,
                                       act="gelu", act_after_res=True)))
        # hyper-network MLPs and IoU head, run over all token rows (the rows of interest are gathered afterwards)
        hyp_out = []
        for k in range(4):
            pre = f"mask_decoder.output_hypernetworks_mlps.{k}"
            a, b2 = tok(hid), tok(hid)
            w = self._lin(pre + ".proj_in")
            steps.append(("conv", ConvPlan(qfn, w[0], w[1], a, k=1, act="relu")))
            w = self._lin(pre + ".layers.0")
            steps.append(("conv", ConvPlan(a, w[0], w[1], b2, k=1, act="relu")))
            o = torch.zeros((1, 1, Pn * T, 32), dtype=torch.float32, device=dev)
            keep.append(o)
            w = self._lin(pre + ".proj_out")
            steps.append(("conv", ConvPlan(b2, w[0], w[1], o, k=1)))
            hyp_out.append(o)
        pre = "mask_decoder.iou_prediction_head"
        a, b2 = tok(hid), tok(hid)
        w = self._lin(pre + ".proj_in")
        steps.append(("conv", ConvPlan(qfn, w[0], w[1], a, k=1, act="relu")))
        w = self._lin(pre + ".layers.0")
        steps.append(("conv", ConvPlan(a, w[0], w[1], b2, k=1, act="relu")))
        iou_o = torch.zeros((1, 1, Pn * T, 16), dtype=torch.float32, device=dev)
        keep.append(iou_o)
        w = self._lin(pre + ".proj_out")
        steps.append(("conv", ConvPlan(b2, w[0], w[1], iou_o, k=1, act="sigmoid")))
        d.update(steps=steps, keep=keep, up2=up2, hyp_out=hyp_out, iou_o=iou_o,
                 sparse=torch.zeros((Pn, 3, hid), dtype=torch.float32, device=dev),
                 logits=torch.zeros((Pn, 4, 4 * S * 4 * S), dtype=torch.float32, device=dev),
                 sel=torch.zeros((Pn,), dtype=torch.int32, device=dev))
        return d

    # ---- execution -------------------------------------------------------------------------------------------------
    def _run(self, steps) -> None:
        l, st, pl = self.l, stream_ptr(), self.planes
        for kind, a in steps:
            if kind == "conv":
                a.run()
            elif kind == "ln":
                x, rows, c, g, b, out, gelu, eps = a
                check(l.mtb_layernorm(ptr(x), rows, c, c, 0, pl, ptr(g), ptr(b), eps, ptr(out), c, 0, pl, gelu, st),
                      "mtb_layernorm")
            elif kind == "pool":
                src, dst, h, c = a
                check(l.mtb_maxpool2x2(ptr(src), ptr(dst), 1, h, h, c, pl, st), "mtb_maxpool2x2")
            elif kind == "up":
                src, dst, h, c = a
                check(l.mtb_upsample2x(ptr(src), ptr(dst), 1, h, h, c, 0, c, 0, c, pl, st), "mtb_upsample2x")
            elif kind == "add":
                x, y, o, rows, c, brows = a
                check(l.mtb_add_planes(ptr(x), ptr(y), ptr(o), rows, c, brows, pl, st), "mtb_add_planes")
            elif kind == "attn":
                self._attn(a)

    def _attn(self, a: dict) -> None:
        d = AttnDesc()
        q, k, v, o = a["q"], a["k"], a["v"], a["out"]
        d.heads, d.hd = a["heads"], a["hd"]
        d.scale = float(a["hd"]) ** -0.5
        d.q, d.k, d.v, d.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
        d.q_ct, d.k_ct, d.v_ct, d.o_ct = q.shape[-1], k.shape[-1], v.shape[-1], o.shape[-1]
        d.q_off, d.k_off, d.v_off, d.o_off = a["q_off"], a["k_off"], a["v_off"], a["o_off"]
        d.q_ps, d.k_ps, d.v_ps, d.o_ps = (q[0].numel(), k[0].numel(), v[0].numel(), o[0].numel())
        d.planes = self.planes
        ws = a["ws"]
        if ws == -1:                       # plain batched attention (decoder)
            d.mode, d.B, d.nq, d.nk = 0, a["B"], a["nq"], a["nk"]
        elif ws == 0:                      # global attention over the whole token grid
            g = a["grid"]
            d.mode, d.B, d.nq, d.nk = 0, 1, g * g, g * g
            if "ws_buf" not in a:          # scratch for the tensor-core path (transposed V), allocated once per step
                need = self.l.mtb_attention_workspace_bytes(1, a["heads"], a["hd"], g * g)
                a["ws_buf"] = torch.empty(max(int(need), 16), dtype=torch.uint8, device=self.device)
            d.workspace, d.workspace_bytes = a["ws_buf"].data_ptr(), a["ws_buf"].numel()
        else:                              # Hiera windows (+ optional query pooling)
            g = a["grid"]
            nw = (g + ws - 1) // ws
            d.mode, d.B = 1, nw * nw
            d.grid_h = d.grid_w = g
            d.ws, d.pool = ws, a["pool"]
            d.nk = ws * ws
            d.nq = (ws // 2) * (ws // 2) if a["pool"] else ws * ws
            pq, pk, pv = a["pads"]
            d.pad_q, d.pad_k, d.pad_v = pq.data_ptr(), pk.data_ptr(), pv.data_ptr()
        check(self.l.mtb_attention(C.byref(d), stream_ptr()), "mtb_attention")

    # ---- public ----------------------------------------------------------------------------------------------------
    def encode(self, img_rgb_u8: torch.Tensor) -> dict:
        """img_rgb_u8: device uint8 HxWx3 (RGB).  Runs resize -> normalise -> patch embed -> Hiera -> neck."""
        from . import graphs
        if self._enc is None:
            self._enc = self._build_encoder()
            self._enc["in1024"] = torch.empty((self.img, self.img, 3), dtype=torch.uint8, device=self.device)
        enc = self._enc
        x = img_rgb_u8
        if x.shape[0] != self.img or x.shape[1] != self.img:
            x = resize_aa_device(x, self.img, self.img)
        enc["in1024"].copy_(x[:, :, :3])
        w = self.sd["vision_encoder.backbone.patch_embed.projection.weight"]
        b = self.sd["vision_encoder.backbone.patch_embed.projection.bias"]

        def body():
            check(self.l.mtb_sam_patch_embed(ptr(enc["in1024"]), self.img, self.img, ptr(self.norm_mean), ptr(self.norm_std),
                                             ptr(w), ptr(b), ptr(self.pos_embed), w.shape[0], 7, 4, 3, ptr(enc["x0"]),
                                             self.planes, stream_ptr()), "mtb_sam_patch_embed")
            self._run(enc["steps"])

        if graphs.ENABLED:
            if "cuda_graph" not in enc:
                enc["cuda_graph"] = graphs.CapturedGraph(body)
            enc["cuda_graph"].replay()
        else:
            body()
        return enc

    def decode(self, enc: dict, boxes_xyxy: torch.Tensor, orig_hw: Tuple[int, int], *, want_logits: bool = False):
        """boxes_xyxy: float32 [P][4] in ORIGINAL page pixels (device or host).  Returns uint8 masks [P][H][W] {0,255}
        (and, optionally, the low-res logits [P][4][256*256], selection [P] and interpolated logits [P][H][W])."""
        H, W = orig_hw
        n_real = int(boxes_xyxy.shape[0])
        dev = self.device
        if n_real == 0:
            return torch.zeros((0, H, W), dtype=torch.uint8, device=dev)
        # Prompt counts are bucketed (multiples of 4): every distinct count owns buffers, launch descriptors and a CUDA graph
        # per page size, and a detector's count changes from page to page.  The padding prompts repeat the first box (the
        # decoder treats prompts independently); their masks are dropped.
        bucket = int(os.environ.get("MTB200_SAM_PROMPT_BUCKET", "4"))
        Pn = -(-n_real // bucket) * bucket if bucket > 1 and not want_logits else n_real
        if Pn != n_real:
            boxes_xyxy = torch.cat([boxes_xyxy.to(dtype=torch.float32),
                                    boxes_xyxy[:1].to(dtype=torch.float32).expand(Pn - n_real, -1)], 0)
        from . import graphs
        if Pn not in self._dec:
            self._dec[Pn] = self._build_decoder(Pn, enc)
            self._dec[Pn]["boxes_in"] = torch.zeros((Pn, 4), dtype=torch.float32, device=dev)
            self._dec[Pn]["sizes"] = collections.OrderedDict()
            while len(self._dec) > max(1, self.max_decoders):
                self._dec.popitem(last=False)
        self._dec.move_to_end(Pn)
        d = self._dec[Pn]
        d["sizes"][(H, W)] = True
        d["sizes"].move_to_end((H, W))
        while len(d["sizes"]) > max(1, self.max_sizes_per_decoder):          # drop the oldest page size's graph + mask buffer
            old, _ = d["sizes"].popitem(last=False)
            d.pop(("graph",) + old, None)
            d.pop(("masks",) + old, None)
        d["boxes_in"].copy_(boxes_xyxy.to(dtype=torch.float32))
        boxes = d["boxes_in"]
        l = self.l
        T, S = d["T"], d["S"]
        npix = 16 * S * S
        gkey = ("graph", H, W)
        if "hyper" not in d:
            d["hyper"] = torch.zeros((Pn, 4, 32), dtype=torch.float32, device=dev)
            d["iou"] = torch.zeros((Pn, 4), dtype=torch.float32, device=dev)
        mkey = ("masks", H, W)
        if mkey not in d:
            d[mkey] = torch.empty((Pn, H, W), dtype=torch.uint8, device=dev)

        def body():
            st = stream_ptr()
            check(l.mtb_sam_prompt_boxes(ptr(boxes), float(self.img / W), float(self.img / H), Pn, ptr(self.gauss_prompt),
                                         self.hidden // 2, ptr(self.pe2), ptr(self.pe3), ptr(self.not_a_point),
                                         float(self.img), ptr(d["sparse"]), st), "mtb_sam_prompt_boxes")
            tokens = torch.cat([self.out_tokens.unsqueeze(0).expand(Pn, -1, -1), d["sparse"]], 1)      # [P][9][256]
            tp = P.split_planes(tokens.reshape(1, 1, Pn * T, self.hidden), self.planes)
            d["pe_q"].copy_(tp)
            d["queries0"].copy_(tp)
            d["keys0"].copy_(enc["emb"].expand(-1, Pn, -1, -1, -1))
            self._run(d["steps"])
            d["hyper"].copy_(torch.stack([d["hyp_out"][k][0, 0, 2 + k::T, :] for k in range(4)], 1))    # [P][4][32]
            d["iou"].copy_(d["iou_o"][0, 0, 1::T, :4])                                                   # [P][4]
            check(l.mtb_sam_hyper_masks(ptr(d["up2"]), self.planes, ptr(d["hyper"]), Pn, 4, 32, npix, ptr(d["logits"]), st),
                  "mtb_sam_hyper_masks")
            check(l.mtb_sam_select_mask(ptr(d["logits"]), ptr(d["iou"]), Pn, 4, npix, self.stab_delta, self.stab_thresh,
                                        ptr(d["sel"]), st), "mtb_sam_select_mask")
            if not want_logits:
                check(l.mtb_sam_mask_write(ptr(d["logits"]), ptr(d["sel"]), 4, 4 * S, ptr(boxes), Pn, H, W, ptr(d[mkey]), None,
                                           st), "mtb_sam_mask_write")

        if graphs.ENABLED and not want_logits:
            if gkey not in d:
                d[gkey] = graphs.CapturedGraph(body)
            d[gkey].replay()
        else:
            body()
        masks = d[mkey]
        iou = d["iou"]
        lo = None
        if want_logits:
            lo = torch.empty((Pn, H, W), dtype=torch.float32, device=dev)
            check(l.mtb_sam_mask_write(ptr(d["logits"]), ptr(d["sel"]), 4, 4 * S, ptr(boxes), Pn, H, W, ptr(masks), ptr(lo),
                                       stream_ptr()), "mtb_sam_mask_write")
        if want_logits:
            return masks, d["logits"], d["sel"], lo, iou
        return masks[:n_real]

    def decode_lowres(self, enc: dict, boxes_xyxy: torch.Tensor, orig_hw: Tuple[int, int]) -> torch.Tensor:
        """Low-res logits [P][256][256] of the selected mask (what Sam2Model returns as pred_masks)."""
        self._last = (enc, boxes_xyxy.clone(), orig_hw)
        self.decode(enc, boxes_xyxy, orig_hw)
        d = self._dec[int(boxes_xyxy.shape[0])]
        S = 4 * d["S"]
        idx = d["sel"].long().view(-1, 1, 1).expand(-1, 1, S * S)
        return torch.gather(d["logits"], 1, idx).view(-1, S, S)

    def _last_full_masks(self, h: int, w: int) -> torch.Tensor:
        """uint8 (P,H,W): bilinear upsample of the last decode's selected logits, > 0, WITHOUT the box clip."""
        enc, boxes, _ = self._last
        Pn = int(boxes.shape[0])
        d = self._dec[Pn]
        big = torch.tensor([[0.0, 0.0, float(w), float(h)]] * Pn, dtype=torch.float32, device=self.device)
        masks = torch.empty((Pn, h, w), dtype=torch.uint8, device=self.device)
        check(self.l.mtb_sam_mask_write(ptr(d["logits"]), ptr(d["sel"]), 4, 4 * d["S"], ptr(big), Pn, h, w, ptr(masks), None,
                                        stream_ptr()), "mtb_sam_mask_write")
        return masks

    def segment(self, img_rgb_u8: torch.Tensor, boxes_xyxy: torch.Tensor):
        enc = self.encode(img_rgb_u8)
        return self.decode(enc, boxes_xyxy, (img_rgb_u8.shape[0], img_rgb_u8.shape[1]))
