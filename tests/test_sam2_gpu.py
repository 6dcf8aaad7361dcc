"""GPU: B200 SAM 2.1 (tcgen05 linears, fused window attention, mask decoder, mask writer) against the real
`transformers.Sam2Model` run in fp32 on the CPU the way the reference calls it (oracle/sam2_oracle.py).

Float outputs (FPN features, low-res mask logits) must agree within 1e-3 abs (BASELINE.json north_star).  Mask BITS are
compared exactly wherever the oracle's own interpolated logit is farther than MARGIN from the `> 0` decision; closer
than that the decision is not well-posed across any two floating-point implementations (the reference's own CUDA bf16
path flips those too), so those pixels are only counted."""
import numpy as np
import pytest
import torch
from PIL import Image

import sam2_oracle as S

pytestmark = pytest.mark.gpu
TOL = 1e-3
MARGIN = 2e-3


def _large_like_config():
    """Hiera-large geometry (the checkpoint the reference loads, core/ml/model_manager.py:202-204: 144..1152 channels,
    2/4/8/16 heads = head dim 72, windows 8/4/16/8) with fewer blocks per stage so the CPU oracle stays in seconds."""
    from transformers import Sam2Config
    cfg = Sam2Config()
    bc = cfg.vision_config.backbone_config
    bc.hidden_size, bc.embed_dim_per_stage = 144, [144, 288, 576, 1152]
    bc.blocks_per_stage, bc.num_attention_heads_per_stage = [1, 2, 4, 2], [2, 4, 8, 16]
    bc.global_attention_blocks, bc.window_size_per_stage = [4, 6], [8, 4, 16, 8]
    bc.window_positional_embedding_background_size = [7, 7]
    cfg.vision_config.backbone_channel_list = [1152, 576, 288, 144]
    return cfg


def _case(seed, h, w, n_boxes, spread, config=None):
    from mangatranslator_b200 import synth
    from mangatranslator_b200.sam2 import Sam2B200
    pg = synth.make_page(seed, h, w, n_bubbles=max(n_boxes, 3))
    boxes = pg.boxes_xyxy[:n_boxes].astype(np.float32)
    m = S.make_model(seed, config=config, spread=spread)
    ref = S.segment(m, S.make_processor(), Image.fromarray(pg.image_rgb), boxes)
    dev = torch.device("cuda:0")
    net = Sam2B200(m.state_dict(), m.config, dev)
    img = torch.from_numpy(pg.image_rgb).to(dev)
    enc = net.encode(img)
    masks, logits, sel, full, iou = net.decode(enc, torch.from_numpy(boxes), (h, w), want_logits=True)
    torch.cuda.synchronize()
    return m, net, enc, ref, masks, logits, sel, full, iou


@pytest.mark.parametrize("hw,spread,large", [((480, 400), 1.0, False), ((600, 448), 100.0, False), ((480, 400), 100.0, True)],
                         ids=["default_init", "spread100", "hiera_large_geometry"])
def test_sam2_matches_transformers_oracle(hw, spread, large):
    from mangatranslator_b200 import planes as P
    m, net, enc, ref, masks, logits, sel, full, iou = _case(3, hw[0], hw[1], 3, spread, _large_like_config() if large else None)
    # encoder / neck features
    extra = (m.prompt_encoder.no_mask_embed.weight.reshape(-1)).view(1, -1, 1, 1)
    emb = P.planes_to_nchw(enc["emb"], 256).cpu() - extra
    assert (emb - ref["image_embeddings"][2]).abs().max().item() < TOL
    assert (P.planes_to_nchw(enc["s1"], 64).cpu() - ref["image_embeddings"][1]).abs().max().item() < TOL
    assert (P.planes_to_nchw(enc["s0"], 32).cpu() - ref["image_embeddings"][0]).abs().max().item() < TOL
    # low-res logits of the selected mask
    got = torch.stack([logits[p, int(sel[p])].view(256, 256) for p in range(logits.shape[0])]).cpu()
    scale = max(1.0, float(ref["pred_masks"].abs().max()))
    assert (got - ref["pred_masks"]).abs().max().item() < TOL * scale
    # IoU head of the selected mask
    got_iou = torch.stack([iou[p, int(sel[p])] for p in range(iou.shape[0])]).cpu()
    assert (got_iou - ref["iou"][:, 0]).abs().max().item() < TOL
    # final masks: bit-exact outside the knife-edge band of the oracle's own logits
    decided = ref["full_logits"].abs().numpy() > MARGIN * scale
    mine, exp = masks.cpu().numpy(), ref["masks"]
    assert set(np.unique(mine)) <= {0, 255}
    assert np.array_equal(mine[decided], exp[decided])
    assert (mine != exp).mean() < 2e-3


def test_sam2_single_box_and_many_boxes_shapes():
    m, net, enc, ref, masks, *_ = _case(5, 384, 512, 1, 100.0)
    assert masks.shape == (1, 384, 512)
    boxes = torch.tensor([[10., 20., 200., 300.]] * 12)
    out = net.decode(enc, boxes, (384, 512))
    assert out.shape == (12, 384, 512) and torch.equal(out[0], out[11])


@pytest.mark.parametrize("shape", [(3, 8, 16, 9, 1024), (2, 2, 32, 100, 77), (1, 4, 96, 512, 512), (12, 8, 16, 9, 4096),
                                   (1, 4, 96, 4096, 4096), (2, 2, 64, 200, 300), (1, 2, 128, 256, 320),
                                   (12, 8, 16, 4096, 9), (2, 4, 32, 300, 17), (1, 8, 16, 9, 1003), (2, 2, 32, 16, 600),
                                   (3, 2, 96, 70, 131), (2, 4, 72, 49, 49), (5, 2, 64, 4, 16), (2, 1, 128, 40, 90)],
                         ids=["few_queries", "general", "tc_hd96", "decoder_t2i", "tc_sam_global", "tc_ragged_hd64",
                              "tc_hd128", "decoder_i2t_few_keys", "few_keys_hd32", "cluster_ragged_keys", "cluster_hd32",
                              "lpq_hd96_two_key_tiles", "lpq_hd72", "lpq_tiny_window", "general_hd128"])
def test_attention_kernels_match_torch(shape):
    """mtb_attention (mode 0) against torch softmax attention on the plane-rounded inputs: every dispatch target
    (general register-blocked kernel, few-queries/many-keys kernel with its keys split over a thread-block cluster,
    many-queries/few-keys kernel, tensor-core kernel) must agree to fp32 accuracy."""
    import ctypes as C
    from mangatranslator_b200 import planes as P
    from mangatranslator_b200._lib import check, lib, stream_ptr
    from mangatranslator_b200.sam2 import AttnDesc, _declare
    B, heads, hd, nq, nk = shape
    dev = torch.device("cuda:0")
    l = lib()
    _declare(l)
    g = torch.Generator(device="cpu").manual_seed(11)
    ct = heads * hd

    def mk(n):
        x = torch.randn(B * n, ct, generator=g).to(dev)
        xp = torch.stack([x.to(torch.bfloat16), (x - x.to(torch.bfloat16).float()).to(torch.bfloat16)])   # hi, lo planes
        return xp.contiguous(), (xp[0].float() + xp[1].float()).double()

    qp, q = mk(nq)
    kp, k = mk(nk)
    vp, v = mk(nk)
    out = torch.zeros(2, B * nq, ct, dtype=torch.bfloat16, device=dev)
    d = AttnDesc()
    d.B, d.heads, d.hd, d.nq, d.nk = B, heads, hd, nq, nk
    d.scale = float(hd) ** -0.5
    d.q, d.k, d.v, d.out = qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), out.data_ptr()
    d.q_ct = d.k_ct = d.v_ct = d.o_ct = ct
    d.q_ps, d.k_ps, d.v_ps, d.o_ps = qp[0].numel(), kp[0].numel(), vp[0].numel(), out[0].numel()
    d.planes, d.mode = 2, 0
    expect_tc = hd in (64, 96, 128) and nq >= 128 and nk >= 256
    ws = torch.empty(max(int(l.mtb_attention_workspace_bytes(B, heads, hd, nk)), 16), dtype=torch.uint8, device=dev)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    n0 = l.mtb_launch_count()
    check(l.mtb_attention(C.byref(d), stream_ptr()), "mtb_attention")
    torch.cuda.synchronize()
    assert (l.mtb_launch_count() - n0 == 2) == expect_tc        # tensor-core path = V transpose + attention
    qh = q.view(B, nq, heads, hd).permute(0, 2, 1, 3)
    kh = k.view(B, nk, heads, hd).permute(0, 2, 1, 3)
    vh = v.view(B, nk, heads, hd).permute(0, 2, 1, 3)
    ref = torch.softmax(qh @ kh.transpose(-1, -2) * d.scale, -1) @ vh
    ref = ref.permute(0, 2, 1, 3).reshape(B * nq, ct)
    got = out[0].double() + out[1].double()
    assert (got - ref).abs().max().item() < 2e-5
