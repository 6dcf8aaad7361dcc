"""TEST INFRASTRUCTURE ONLY — NumPy walk through exactly what csrc/conjoined_kernels.cu computes from a split plan
(mangatranslator_b200.conjoined.plan_split): seeds, per-pair linear classifiers in float64, and the nearest-seed rule
through the closed-form 16.16 chamfer norm with per-row nearest-seed tables.  Lets the CPU suite check the plan and the
kernel's algorithm against the oracle without a GPU."""
import numpy as np

A, B, C = 65536, 91750, 143976


def chamfer_norm(dx, dy):
    M, m = np.maximum(dx, dy), np.minimum(dx, dy)
    return np.where(M >= 2 * m, m * C + (M - 2 * m) * A, (M - m) * C + (2 * m - M) * B)


def row_nearest(seed):
    """|x - nearest seed x| per row (int64, -1 where the row has no seed pixel)."""
    h, w = seed.shape
    big = 10 ** 9
    xs = np.arange(w)[None, :]
    left = np.maximum.accumulate(np.where(seed, xs, -big), axis=1)
    right = np.minimum.accumulate(np.where(seed, xs, big)[:, ::-1], axis=1)[:, ::-1]
    d = np.minimum(xs - left, right - xs)
    return np.where(d > 10 ** 8, -1, d)


def apply_plan(parent_u8, plan, include_child_rects=True):
    base = np.asarray(parent_u8) > 0
    h, w = base.shape
    k = len(plan.rects)
    rects = []
    for (x0, y0, x1, y1) in plan.rects:
        r = np.zeros((h, w), bool)
        r[y0:y1, x0:x1] = True
        rects.append(r)
        if include_child_rects:
            base = base | r
    owned = [base & r for r in rects]
    yy, xx = np.mgrid[0:h, 0:w]
    for kk in range(k):
        if not owned[kk].any() and base.any():
            cx, cy = plan.centers[kk]
            d = (xx - cx) * (xx - cx) + (yy - cy) * (yy - cy)
            d = np.where(base, d, np.inf)
            n = int(np.argmin(d))                        # first minimum in row-major order
            owned[kk].flat[n] = True
    for (i, j, mode, cx, cy, ax, ay, off) in plan.pairs:
        zone = base & rects[i] & rects[j]
        owned[i] &= ~zone
        owned[j] &= ~zone
        if mode:
            v = ((xx - cx) * ax + (yy - cy) * ay) - off
            owned[i] |= zone & ((v <= 0) if mode == 1 else (v >= 0))
            owned[j] |= zone & ((v > 0) if mode == 1 else (v < 0))
    taken = np.zeros_like(base)
    for m in owned:
        taken |= m
    rest = base & ~taken
    out = [m.copy() for m in owned]
    ys, xs = np.nonzero(rest)
    if len(ys):
        dist = np.full((k, len(ys)), np.inf, np.float32)
        rows = np.arange(h)
        for kk in range(k):
            if not owned[kk].any():
                continue
            rn = row_nearest(owned[kk])                  # [h][w]
            col = rn[:, xs]                              # [h][n]: nearest seed distance of every row, at each pixel's column
            dy = np.abs(rows[:, None] - ys[None, :])
            val = np.where(col >= 0, chamfer_norm(np.maximum(col, 0), dy), np.iinfo(np.int64).max)
            dist[kk] = val.min(axis=0).astype(np.uint32).astype(np.float32)
        win = np.argmin(dist, axis=0)
        for kk in range(k):
            sel = win == kk
            out[kk][ys[sel], xs[sel]] = True
    return [m.astype(np.uint8) * 255 for m in out]
