"""GPU dev check #1: UMMA shifted-descriptor probe + conv kernel vs torch fp32 conv. Run under gpurun."""
import sys, json, time
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from mangatranslator_b200 import _lib, planes as P
from mangatranslator_b200.ops import ConvPlan

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
L = _lib.lib()
LX = _lib.exp_lib()
out = {}


def probe():
    res = []
    rows = torch.arange(512, dtype=torch.float32, device=dev)
    A_row = (rows[:, None] % 256).expand(512, 64).contiguous()          # value = row % 256 (bf16 exact)
    A_row_hi = (rows[:, None] // 256).expand(512, 64).contiguous()      # 0/1 : which half
    A_col = torch.arange(64, dtype=torch.float32, device=dev)[None, :].expand(512, 64).contiguous()
    B = torch.eye(64, device=dev)
    D = torch.zeros(128, 64, device=dev)
    Bb = B.to(torch.bfloat16).contiguous()
    for shift in [0, 1, 2, 3, 7, 8, 9, 10, 11, 20]:
        for sbo in [1024, 1280, 2048]:
            for bo in sorted({0, shift & 7}):
                maps = []
                for A in (A_row, A_row_hi, A_col):
                    Ab = A.to(torch.bfloat16).contiguous()
                    _lib.check(LX.mtb_exp_shifted_desc(Ab.data_ptr(), Bb.data_ptr(), D.data_ptr(), shift, sbo, bo, None))
                    torch.cuda.synchronize()
                    maps.append(D.clone())
                src_row = (maps[0] + 256 * maps[1]).long()   # [128][64] source row of each element
                src_col = maps[2].long()
                r = torch.arange(128, device=dev)
                exp_row = shift + (r // 8) * (sbo // 128) + (r % 8)
                ok_row = bool((src_row == exp_row[:, None]).all())
                ok_col = bool((src_col == torch.arange(64, device=dev)[None, :]).all())
                res.append(dict(shift=shift, sbo=sbo, bo=bo, ok=ok_row and ok_col, ok_row=ok_row, ok_col=ok_col,
                                row0=src_row[:10, 0].tolist(), col_r1=src_col[1, ::8].tolist()))
                print(res[-1], flush=True)
    return res


def conv_case(name, n, cin, cout, h, w, k, stride, pad, planes_in, planes_out, act, use_res, use_sums, bias=True):
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev) if bias else None
    xp = P.nchw_to_planes(x, planes_in)
    wp = P.conv_weight_to_planes(wt, planes_in)
    bp = P.pad_bias(b, cout)
    coutp = wp.shape[2]
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    if planes_out == 4:
        o = torch.zeros(n, ho, wo, coutp, device=dev)
    else:
        o = torch.zeros(planes_out, n, ho, wo, coutp, dtype=torch.bfloat16, device=dev)
    res = None
    resf = None
    if use_res:
        resf = torch.randn(n, cout, ho, wo, device=dev)
        res = P.nchw_to_planes(resf, 2, cpad=16)
        assert res.shape[-1] == coutp, (res.shape, coutp)
    plan0 = ConvPlan(xp, wp, bp, o, k=k, stride=stride, pad=pad, act=act, residual=res)
    sums = torch.zeros(plan0.num_sum_rows, coutp, device=dev) if use_sums else None
    plan = ConvPlan(xp, wp, bp, o, k=k, stride=stride, pad=pad, act=act, residual=res, tile_sums=sums)
    plan.run()
    torch.cuda.synchronize()
    # reference: fp32 conv on the exact values the kernel saw
    xr = P.planes_to_nchw(xp, cin)
    wr = P.merge_planes(wp)[:, :cout, :cin].reshape(k, k, cout, cin).permute(2, 3, 0, 1).contiguous()
    ref = F.conv2d(xr.double(), wr.double(), b.double() if b is not None else None, stride=stride, padding=pad)
    if act == "relu": ref = F.relu(ref)
    elif act == "silu": ref = F.silu(ref)
    elif act == "gelu": ref = F.gelu(ref)
    if use_res: ref = ref + P.planes_to_nchw(res, cout).double()
    got = (o[..., :cout].permute(0, 3, 1, 2) if planes_out == 4 else P.planes_to_nchw(o, cout)).double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    r = dict(name=name, max_abs_err=err, ref_absmax=scale)
    if use_sums:
        s = sums.sum(0)[:cout].double()
        rs = ref.sum((0, 2, 3))
        r["sums_err"] = (s - rs).abs().max().item()
        r["sums_scale"] = rs.abs().max().item()
    # timing
    for _ in range(3): plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 10
    for _ in range(iters): plan.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * ho * wo * cout * cin * k * k
    r["ms"] = ms
    r["tflops_alg"] = flops / ms / 1e9
    print(r, flush=True)
    return r


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "probe"):
        try:
            out["probe"] = probe()
        except Exception as e:
            print("PROBE FAILED", repr(e)); out["probe_error"] = repr(e)
    if which in ("all", "conv"):
        cases = [
            ("lin_bf16_small", 1, 64, 64, 1, 256, 1, 1, 0, 1, 1, None, False, False),
            ("lin_x3_small", 1, 128, 96, 1, 300, 1, 1, 0, 2, 2, "gelu", False, False),
            ("conv3_bf16", 1, 64, 64, 32, 48, 3, 1, 1, 1, 1, "relu", False, False),
            ("conv3_x3_res_sums", 2, 64, 64, 40, 56, 3, 1, 1, 2, 2, None, True, True),
            ("conv3_x3_f32out", 1, 128, 48, 33, 37, 3, 1, 1, 2, 4, "silu", False, False),
            ("conv3_s2_x3", 1, 64, 128, 64, 80, 3, 2, 1, 2, 2, "silu", False, False),
            ("conv1_x3_big_n", 1, 192, 576, 20, 24, 1, 1, 0, 2, 2, "silu", False, False),
            ("rcan_layer_x3", 1, 64, 64, 1536, 1024, 3, 1, 1, 2, 2, "relu", False, True),
            ("rcan_layer_bf16", 1, 64, 64, 1536, 1024, 3, 1, 1, 1, 1, "relu", False, False),
            ("rcan_up_x3", 1, 64, 256, 512, 512, 3, 1, 1, 2, 2, None, False, False),
        ]
        out["conv"] = []
        for c in cases:
            try:
                out["conv"].append(conv_case(*c))
            except Exception as e:
                print("CASE FAILED", c[0], repr(e), flush=True)
                out["conv"].append(dict(name=c[0], error=repr(e)))
                break
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/dev_conv.json", "w"), indent=1)
