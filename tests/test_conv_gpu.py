"""GPU: tcgen05 implicit-GEMM conv / linear kernel vs a float64 torch convolution of the same (plane-rounded) inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # name, n, cin, cout, h, w, k, stride, pad, planes_in, planes_out, act, residual, sums, tol
    ("lin_bf16", 1, 64, 64, 1, 256, 1, 1, 0, 1, 1, None, False, False, 5e-2),
    ("lin_x3_gelu", 1, 128, 96, 1, 300, 1, 1, 0, 2, 2, "gelu", False, False, 2e-4),
    ("conv3_bf16_relu", 1, 64, 64, 32, 48, 3, 1, 1, 1, 1, "relu", False, False, 5e-2),
    ("conv3_x3_res_sums", 2, 64, 64, 40, 56, 3, 1, 1, 2, 2, None, True, True, 2e-4),
    ("conv3_x3_f32out_silu", 1, 128, 48, 33, 37, 3, 1, 1, 2, 4, "silu", False, False, 2e-4),
    ("conv3_s2_x3", 1, 64, 128, 64, 80, 3, 2, 1, 2, 2, "silu", False, False, 2e-4),
    ("conv1_x3_n576", 1, 192, 576, 20, 24, 1, 1, 0, 2, 2, "silu", False, False, 2e-4),
    ("conv3_x3_odd_size", 1, 64, 64, 19, 23, 3, 1, 1, 2, 2, "relu", False, True, 2e-4),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_fp64_reference(case):
    from mangatranslator_b200 import planes as P
    from mangatranslator_b200.ops import ConvPlan
    name, n, cin, cout, h, w, k, stride, pad, pin, pout, act, use_res, use_sums, tol = case
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    xp, wp, bp = P.nchw_to_planes(x, pin), P.conv_weight_to_planes(wt, pin), P.pad_bias(b, cout)
    coutp = wp.shape[2]
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    o = (torch.zeros(n, ho, wo, coutp, device=dev) if pout == 4 else
         torch.zeros(pout, n, ho, wo, coutp, dtype=torch.bfloat16, device=dev))
    res = P.nchw_to_planes(torch.randn(n, cout, ho, wo, device=dev), 2, cpad=16) if use_res else None
    probe = ConvPlan(xp, wp, bp, o, k=k, stride=stride, pad=pad, act=act, residual=res)
    sums = torch.zeros(probe.num_sum_rows, coutp, device=dev) if use_sums else None
    plan = ConvPlan(xp, wp, bp, o, k=k, stride=stride, pad=pad, act=act, residual=res, tile_sums=sums)
    plan.run()
    torch.cuda.synchronize()
    xr = P.planes_to_nchw(xp, cin).double()
    wr = P.merge_planes(wp)[:, :cout, :cin].reshape(k, k, cout, cin).permute(2, 3, 0, 1).contiguous().double()
    ref = F.conv2d(xr, wr, b.double(), stride=stride, padding=pad)
    ref = {"relu": F.relu, "silu": F.silu, "gelu": F.gelu, None: lambda t: t}[act](ref)
    if use_res:
        ref = ref + P.planes_to_nchw(res, cout).double()
    got = (o[..., :cout].permute(0, 3, 1, 2) if pout == 4 else P.planes_to_nchw(o, cout)).double()
    assert (got - ref).abs().max().item() < tol
    if use_sums:
        s = sums.sum(0)[:cout].double()
        rs = ref.sum((0, 2, 3))
        assert (s - rs).abs().max().item() < 1e-3 * max(1.0, rs.abs().max().item())


def test_pixel_major_halo_kernel_still_matches():
    """mode 3 pins the pixel-major bf16x3 halo kernel (conv_halo.cu), kept as the A/B partner of the channel-major one."""
    from mangatranslator_b200 import planes as P
    from mangatranslator_b200.ops import ConvPlan
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    x = torch.randn(1, 64, 70, 45, device=dev)
    wt = torch.randn(64, 64, 3, 3, device=dev) / 24
    xp, wp = P.nchw_to_planes(x, 2), P.conv_weight_to_planes(wt, 2)
    outs = []
    for mode in (3, 2):
        o = torch.zeros(2, 1, 70, 45, 64, dtype=torch.bfloat16, device=dev)
        ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", mode=mode).run()
        torch.cuda.synchronize()
        outs.append(P.planes_to_nchw(o, 64).double())
    ref = F.relu(F.conv2d(P.planes_to_nchw(xp, 64).double(),
                          P.merge_planes(wp)[:, :64, :64].reshape(3, 3, 64, 64).permute(2, 3, 0, 1).contiguous().double(),
                          padding=1))
    for got in outs:
        assert (got - ref).abs().max().item() < 2e-4


def test_border_line_sums_from_the_conv_epilogue():
    """The channel-major halo kernel can emit per-channel sums of its output over row 0, row H-1, column 0, column W-1
    next to the totals (the RCAB gate needs them); other kernels must refuse."""
    from mangatranslator_b200 import planes as P
    from mangatranslator_b200.ops import ConvPlan
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    h, w = 67, 41
    x = torch.randn(1, 64, h, w, device=dev)
    wt = torch.randn(64, 64, 3, 3, device=dev) / 24
    xp, wp = P.nchw_to_planes(x, 2), P.conv_weight_to_planes(wt, 2)
    o = torch.zeros(2, 1, h, w, 64, dtype=torch.bfloat16, device=dev)
    probe = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu")
    rows = probe.num_sum_rows
    sums = torch.zeros(rows, 64, device=dev)
    border = torch.zeros(rows, 4, 64, device=dev)
    plan = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", tile_sums=sums)
    assert plan.set_border_sums(border)
    plan.run()
    torch.cuda.synchronize()
    y = P.planes_to_nchw(o, 64).double()[0]
    ref = torch.stack([y[:, 0, :].sum(1), y[:, h - 1, :].sum(1), y[:, :, 0].sum(1), y[:, :, w - 1].sum(1)])
    got = border.sum(0).double()
    assert (got - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())
    assert (sums.sum(0).double() - y.sum((1, 2))).abs().max().item() < 1e-3 * max(1.0, y.sum((1, 2)).abs().max().item())
    per_tap = ConvPlan(xp, wp, None, o, k=3, pad=1, act="relu", tile_sums=torch.zeros(probe.num_mtiles * 4, 64, device=dev), mode=1)
    assert not per_tap.set_border_sums(border)
