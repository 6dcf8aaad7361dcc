"""GPU: device time of one SAM 2.1 encode + 12-box decode (CUDA graphs on), for A/B runs of the attention kernels
(MTB200_ATTN_LPQ / MTB200_ATTN_FEWK / MTB200_ATTN_CLUSTER = 0 select the round-1 kernels)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MTB200_SYNTHETIC_WEIGHTS", "1")
import numpy as np, torch
from mangatranslator_b200 import synth, weights as W
from mangatranslator_b200.sam2 import Sam2B200
dev = torch.device("cuda:0")
cfg, sd = W.sam2_model_and_state(0, "tiny")
net = Sam2B200(sd, cfg, dev)
pg = synth.make_page(1, 1536, 1024)
img = torch.from_numpy(pg.image_rgb.copy()).to(dev)
boxes = torch.tensor(pg.boxes_xyxy[:12], dtype=torch.float32, device=dev)
def run():
    enc = net.encode(img)
    return net.decode(enc, boxes, (1536, 1024))
for _ in range(3): out = run()
torch.cuda.synchronize()
res = {}
for name, fn in (("encode", lambda: net.encode(img)), ("encode+decode", run)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(20):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name] = round(float(np.median(ts)), 3)
res["mask_pixels"] = int((out > 0).sum())
res["env"] = {k: os.environ.get(k) for k in ("MTB200_ATTN_LPQ", "MTB200_ATTN_FEWK", "MTB200_ATTN_CLUSTER")}
print(json.dumps(res))

# ---- the Hiera window-attention launches on their own (mode 1), median of 20 -------------------------------------------
import ctypes as C
from mangatranslator_b200._lib import check, lib, stream_ptr
from mangatranslator_b200.sam2 import AttnDesc, _declare
l = lib(); _declare(l)
def window_case(grid, ws, heads, hd, pool):
    Cq = heads * hd
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn((grid * grid, 3 * Cq), generator=g).to(dev)
    qp = torch.stack([qkv.to(torch.bfloat16), (qkv - qkv.to(torch.bfloat16).float()).to(torch.bfloat16)]).contiguous()
    go = grid // 2 if pool else grid
    out = torch.zeros((2, go * go, Cq), dtype=torch.bfloat16, device=dev)
    pads = torch.zeros((3, Cq), dtype=torch.float32, device=dev)
    d = AttnDesc()
    nw = (grid + ws - 1) // ws
    d.heads, d.hd, d.scale = heads, hd, hd ** -0.5
    d.q = d.k = d.v = qp.data_ptr(); d.out = out.data_ptr()
    d.q_ct = d.k_ct = d.v_ct = 3 * Cq; d.o_ct = Cq
    d.q_off, d.k_off, d.v_off, d.o_off = 0, Cq, 2 * Cq, 0
    d.q_ps = d.k_ps = d.v_ps = qp[0].numel(); d.o_ps = out[0].numel()
    d.planes, d.mode, d.B = 2, 1, nw * nw
    d.grid_h = d.grid_w = grid; d.ws, d.pool = ws, int(pool)
    d.nk = ws * ws; d.nq = (ws // 2) ** 2 if pool else ws * ws
    d.pad_q, d.pad_k, d.pad_v = pads[0].data_ptr(), pads[1].data_ptr(), pads[2].data_ptr()
    for _ in range(3): check(l.mtb_attention(C.byref(d), stream_ptr()), "attn")
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); check(l.mtb_attention(C.byref(d), stream_ptr()), "attn"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(float(np.median(ts)), 1), float(out.float().abs().sum())
cases = {"s1 256^2 ws8 h1": (256, 8, 1, 96, False), "s2a 256^2 ws8 h2 pool": (256, 8, 2, 96, True), "s2b 128^2 ws4 h2": (128, 4, 2, 96, False),
         "s3a 128^2 ws4 h4 pool": (128, 4, 4, 96, True), "s3 64^2 ws14 h4": (64, 14, 4, 96, False), "s4a 64^2 ws14 h8 pool": (64, 14, 8, 96, True),
         "s4 32^2 ws7 h8": (32, 7, 8, 96, False)}
print(json.dumps({k: window_case(*v) for k, v in cases.items()}))
