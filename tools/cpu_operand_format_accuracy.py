"""CPU experiment (no GPU): how accurate is the seeded 10x20 RCAN when the 64->64 3x3 body convs see their operands in
cheaper formats than the bf16x3 parity path?  Everything is evaluated in float64 with only the operand roundings applied,
so the numbers isolate the format.  Prints max / mean |dy| on the network output against the exact float64 result.

  bf16x3          X = bf16 hi + bf16 lo, W = bf16 hi + bf16 lo, products hi*hi + hi*lo + lo*hi      (today's kernel)
  tf32            both operands rounded to 11 significant bits
  fp16            X rounded to fp16, W exact (fp16 hi + lo)                                          (one MMA per tap)
  fp16+e5m2       ... plus the correction product e5m2(X - fp16(X)) * e5m2(W)                        (kind::f8f6f4, K = 32)
  fp16+e4m3       same with e4m3 (underflows: X - fp16(X) is ~2^-12 |X|)
  fp16+mxfp4      same with e2m1 values and one power-of-two scale per 32 input channels (kind::mxf4, K = 64)
  +trunk3         additionally every RCAB / group output (the residual trunk) is STORED as fp16 hi + e5m2 lo (3 bytes)

    python tools/cpu_operand_format_accuracy.py [size] [seeds] [init]
init = "bench" (mangatranslator_b200.weights.rcan_state_dict, the bench's weights; smoothed random input) or "test"
(torch default init via oracle make_model + random uint8 input: what tests/test_rcan_gpu.py's full-depth case runs,
where the 1e-3 bound is asserted).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.rcan_oracle import RCAB, RCAN, ResidualGroup  # noqa: E402  (the architecture only; nothing here is product code)
from mangatranslator_b200.weights import rcan_state_dict  # noqa: E402

F8 = {"e5m2": torch.float8_e5m2, "e4m3": torch.float8_e4m3fn}


def q(x, dt):
    return x.to(dt).to(torch.float64)


def tf32(x):
    xi = x.float().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32).double()


class Conv(torch.nn.Module):
    def __init__(self, conv, mode):
        super().__init__()
        self.c, self.mode = conv, mode

    def forward(self, x):
        w, b, m = self.c.weight, self.c.bias, self.mode
        if m == "bf16x3":
            xh, wh = q(x, torch.bfloat16), q(w, torch.bfloat16)
            xl, wl = q(x - xh, torch.bfloat16), q(w - wh, torch.bfloat16)
            return F.conv2d(xh, wh + wl, b, padding=1) + F.conv2d(xl, wh, None, padding=1)
        if m == "tf32":
            return F.conv2d(tf32(x), tf32(w), b, padding=1)
        xh, wh = q(x, torch.float16), q(w, torch.float16)
        wl = q(w - wh, torch.float16)
        if m.startswith("fp16w1"):               # weights as ONE fp16 term (18 or 27 slots instead of 36 / 54)
            wl = wl * 0
            m = m.replace("fp16w1", "fp16")
        y = F.conv2d(xh, wh + wl, b, padding=1)
        if m.endswith("+mxfp4"):
            y = y + F.conv2d(mxfp4(x - xh, 1), mxfp4(w, 1), None, padding=1)
        elif "+" in m:
            name = m.split("+")[1]
            sc = 1.0
            if "s" in name:                      # "e5m2s4": residual scaled by 2^4 before the cast, weights by 2^-4
                name, sh = name.split("s")
                sc = 2.0 ** int(sh)
            dt = F8[name]
            y = y + F.conv2d(q(((x - xh) * sc).float(), dt), q((w / sc).float(), dt), None, padding=1)
        return y


E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def mxfp4(t, dim):
    """e2m1 with a shared power-of-two scale per block of 32 along `dim` (the contraction dimension)."""
    t = t.movedim(dim, -1)
    shape = t.shape
    b = t.reshape(*shape[:-1], shape[-1] // 32, 32)
    amax = b.abs().amax(-1, keepdim=True).clamp_min(1e-300)
    scale = torch.exp2(torch.ceil(torch.log2(amax / 6.0)))
    v = (b / scale).abs().clamp(max=6.0)
    idx = (v.unsqueeze(-1) - E2M1).abs().argmin(-1)
    qv = E2M1[idx] * torch.sign(b) * scale
    return qv.reshape(shape).movedim(-1, dim)


def store3(x):
    xh = q(x, torch.float16)
    return xh + q((x - xh).float(), torch.float8_e5m2)


INIT = "bench"


def build(seed, mode):
    if INIT == "test":
        from oracle.rcan_oracle import make_model
        m = make_model(seed).double().eval()
    else:
        m = RCAN().double().eval()
        m.load_state_dict({k: v.double() for k, v in rcan_state_dict(seed).items()})
    conv_mode = mode.replace("+trunk3", "") if mode else None

    def rec(mod):
        for n, ch in list(mod.named_children()):
            if isinstance(ch, torch.nn.Conv2d) and ch.kernel_size == (3, 3) and ch.in_channels == ch.out_channels == 64:
                setattr(mod, n, Conv(ch, conv_mode))
            else:
                rec(ch)
    if mode:
        rec(m)
        if mode.endswith("+trunk3"):
            for mod in m.modules():
                if isinstance(mod, (RCAB, ResidualGroup)):
                    mod.register_forward_hook(lambda _m, _i, out: store3(out))
    return m


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    global INIT
    INIT = sys.argv[3] if len(sys.argv) > 3 else "bench"
    torch.set_num_threads(os.cpu_count() or 1)
    for seed in range(seeds):
        g = torch.Generator().manual_seed(seed + 100)
        if INIT == "test":
            x = torch.randint(0, 256, (1, 3, size, size), generator=g).double() / 255.0
        else:
            x = torch.rand((1, 3, size, size), generator=g, dtype=torch.float64)
            x = F.avg_pool2d(F.pad(x, (2, 2, 2, 2), mode="reflect"), 5, 1)
        with torch.no_grad():
            ref = build(seed, None)(x)
            modes = ("bf16x3", "tf32", "fp16", "fp16+e5m2", "fp16+e4m3", "fp16+mxfp4", "fp16+e5m2+trunk3")
            if len(sys.argv) > 4:
                modes = tuple(sys.argv[4].split(","))
            for mode in modes:
                d = (build(seed, mode)(x) - ref).abs()
                print(f"seed {seed} {mode:18s} max {float(d.max()):.2e} mean {float(d.mean()):.2e}", flush=True)


if __name__ == "__main__":
    main()
