// Safe text box of a cleaned bubble mask: bit-exact re-implementation of the reference's
// core/image/image_utils.py:173-348 `calculate_centroid_expansion_box` (what its renderer runs per bubble on the masks
// the cleaning stage produced), redesigned for the GPU.  One CTA per bubble, all bubbles of a page (or batch) in one
// launch; everything happens inside the bubble's WINDOW = tight bounding box of the mask + a one-pixel ring, because a
// pixel's distance to the nearest zero never looks further than that ring.
//
//   reference (full frame, per bubble)                              here (window, integer)
//   ---------------------------------------------------------------------------------------------------------------
//   pad with a ring of zeros, cv2.distanceTransform(DIST_L2,       exact squared Euclidean distance d2:  column sweeps
//   DIST_MASK_PRECISE) = float32 sqrt of the exact squared          give g = distance to the nearest zero in the column,
//   distance (:210-217)                                             then d2(x) = min_x' (x-x')^2 + g(x')^2 searched
//                                                                   outwards from x until (x-x')^2 >= best
//   dist >= padding (:218)                                          d2 >= t2, t2 = smallest n with sqrtf(n) >= (float)padding
//   cv2.moments -> m10/m00, m01/m00 (:228-234)                      64-bit integer sums, (255*Sx)/(255*S) in double
//   cv2.minMaxLoc (:237)                                            max of (d2, first raster index) as one 64-bit key
//   dist[centroid] < max*0.70 -> pole of inaccessibility (:239-254) float32 sqrt + the double product cast to float32
//   nearest safe pixel by float64 sqrt distance, np.argmin (:262-281) IEEE dsub/dmul/dadd/dsqrt, (value, raster index) minimum
//   four np.where ray casts (:283-293)                              one parallel sweep of the anchor's row and column
//   -1 rule, Python round() of the corner, bounds (:295-327)        same integer / double arithmetic (rint = half to even)
//
// The same source compiles for the device and, for the unit tests only, as a sequential host emulation (hd_emul.cuh).
#pragma once
#include <stdint.h>

#include "hd_emul.cuh"

namespace mtbsafe {

enum Status : int {
  ST_OK = 0,
  ST_EMPTY_MASK = 1,     // reference: ImageProcessingError("Invalid or empty mask provided") (image_utils.py:204-205)
  // the next three all end in ImageProcessingError("Safe area calculation failed") (:348)
  ST_NO_SAFE_AREA = 2,   // padding larger than the bubble (:220-226)
  ST_BAD_DIMS = 3,       // ray casts leave no width or height (:306-312)
  ST_OUT_OF_BOUNDS = 4,  // box leaves the image (:329-335)
  ST_WORKSPACE = 5,      // window larger than the caller's workspace (no reference counterpart)
};

enum Moved : int { MOVED_POLE = 1, MOVED_NEAREST = 2 };

struct Job {
  const uint8_t* mask;  // H x W uint8, nonzero = bubble interior (the reference passes 0/255)
  long long pitch;      // bytes between rows
  int H, W;
  uint32_t t2;          // safe <=> d2 >= t2
  int cap;              // capacity of g / safe in pixels; the window (bbox + ring) must fit
  uint16_t* g;          // workspace [cap]
  uint8_t* safe;        // workspace [cap]
};

struct Result {
  int status;
  int box[4];        // x, y, width, height (:320)
  int moved;         // Moved flags: which anchor rule fired
  int max_d2;        // squared distance at the pole of inaccessibility
  int anchor[2];     // integer anchor pixel the rays were cast from
  int mask_bbox[4];  // in: {-x0, -y0, x1, y1} of the nonzero pixels (running maxima of the bounds pass); out: x0, y0, x1, y1
  int reserved;
  double cx, cy;     // centroid returned to the caller (:255)
};

struct Shared {
  unsigned long long cnt, sx, sy, maxkey, minkey;
  int minidx, flag, px, py, left, right, up, down;
  double cx, cy;
};

MTB_HD uint32_t d2_at(const uint16_t* grow, int ww, int x) {
  const uint32_t gv = grow[x];
  uint32_t best = gv * gv;
  for (int dx = 1; static_cast<uint32_t>(dx) * dx < best; ++dx) {
    const uint32_t dd = static_cast<uint32_t>(dx) * dx;
    if (x - dx >= 0) {
      const uint32_t a = grow[x - dx];
      const uint32_t c = dd + a * a;
      if (c < best) best = c;
    }
    if (x + dx < ww) {
      const uint32_t a = grow[x + dx];
      const uint32_t c = dd + a * a;
      if (c < best) best = c;
    }
  }
  return best;
}

MTB_HD void safe_job(const Job& J, Result& R, Shared* sh) {
  const int tid = MTB_TID, nthr = MTB_NTHR;
  const int H = J.H, W = J.W;
  int bx0 = -R.mask_bbox[0], by0 = -R.mask_bbox[1], bx1 = R.mask_bbox[2], by1 = R.mask_bbox[3];
  MTB_SYNC();   // everyone has read the bounds before thread 0 overwrites them
  if (tid == 0) {
    R.box[0] = R.box[1] = R.box[2] = R.box[3] = 0;
    R.moved = R.max_d2 = R.anchor[0] = R.anchor[1] = R.reserved = 0;
    R.cx = R.cy = 0.0;
  }
  if (bx1 < 0 || by1 < 0) {  // np.any(mask) is False
    if (tid == 0) {
      R.status = ST_EMPTY_MASK;
      R.mask_bbox[0] = R.mask_bbox[1] = 0;
      R.mask_bbox[2] = R.mask_bbox[3] = -1;
    }
    return;
  }
  const int tx0 = bx0, ty0 = by0, tx1 = bx1, ty1 = by1;   // tight bounds, reported back
  if (J.t2 == 0) {  // padding <= 0: every pixel of the image is "safe" (0 >= padding), the window is the whole image
    bx0 = by0 = 0;
    bx1 = W - 1;
    by1 = H - 1;
  }
  const int wx0 = bx0 - 1, wy0 = by0 - 1, ww = bx1 - bx0 + 3, wh = by1 - by0 + 3;
  if (static_cast<long long>(ww) * wh > J.cap) {
    if (tid == 0) {
      R.status = ST_WORKSPACE;
      R.mask_bbox[0] = tx0; R.mask_bbox[1] = ty0; R.mask_bbox[2] = tx1; R.mask_bbox[3] = ty1;
    }
    return;
  }
  uint16_t* g = J.g;
  uint8_t* safe = J.safe;
  if (tid == 0) {
    sh->cnt = sh->sx = sh->sy = sh->maxkey = 0;
    sh->minkey = ~0ull;
    sh->minidx = 0x7fffffff;
    sh->flag = 0;
  }

  // 1. g(x, y) = distance to the nearest zero of column x (window rows 0 and wh-1 are zero: ring or outside the image)
  for (int x = tid; x < ww; x += nthr) {
    const int X = wx0 + x;
    const bool col_in = X >= 0 && X < W;
    uint32_t run = 0;
    for (int y = 0; y < wh; ++y) {
      const int Y = wy0 + y;
      const bool on = col_in && Y >= 0 && Y < H && J.mask[static_cast<long long>(Y) * J.pitch + X] != 0;
      run = on ? run + 1 : 0;
      g[static_cast<size_t>(y) * ww + x] = static_cast<uint16_t>(run < 65535u ? run : 65535u);
    }
    run = 0;
    for (int y = wh - 1; y >= 0; --y) {
      const uint16_t v = g[static_cast<size_t>(y) * ww + x];
      run = v ? run + 1 : 0;
      if (run < v) g[static_cast<size_t>(y) * ww + x] = static_cast<uint16_t>(run);
    }
  }
  MTB_SYNC();

  // 2. exact squared distance per pixel, safe map, moments, maximum
  {
    unsigned long long cnt = 0, sx = 0, sy = 0, best = 0;
    const int n = ww * wh;
    for (int i = tid; i < n; i += nthr) {
      const int y = i / ww, x = i - y * ww;
      const int X = wx0 + x, Y = wy0 + y;
      const bool in_img = X >= 0 && X < W && Y >= 0 && Y < H;
      const uint32_t d2 = g[i] ? d2_at(g + static_cast<size_t>(y) * ww, ww, x) : 0u;
      const bool s = in_img && d2 >= J.t2;
      safe[i] = s ? 255 : 0;
      if (s) {
        ++cnt;
        sx += static_cast<unsigned>(X);
        sy += static_cast<unsigned>(Y);
      }
      if (in_img) {
        const unsigned long long key = (static_cast<unsigned long long>(d2) << 32) | (0xffffffffu - static_cast<uint32_t>(i));
        if (key > best) best = key;
      }
    }
    if (cnt) {
      mtb_atomic_add(&sh->cnt, cnt);
      mtb_atomic_add(&sh->sx, sx);
      mtb_atomic_add(&sh->sy, sy);
    }
    mtb_atomic_max64(&sh->maxkey, best);
  }
  MTB_SYNC();

  // 3. anchor: centroid of the safe area, or the pole of inaccessibility when the centroid sits in a constriction
  if (tid == 0) {
    R.mask_bbox[0] = tx0; R.mask_bbox[1] = ty0; R.mask_bbox[2] = tx1; R.mask_bbox[3] = ty1;
    const uint32_t maxd2 = static_cast<uint32_t>(sh->maxkey >> 32);
    R.max_d2 = static_cast<int>(maxd2);
    if (sh->cnt == 0) {
      R.status = ST_NO_SAFE_AREA;
      sh->flag = -1;
    } else {
      double cx = mtb_ddiv(static_cast<double>(255ull * sh->sx), static_cast<double>(255ull * sh->cnt));
      double cy = mtb_ddiv(static_cast<double>(255ull * sh->sy), static_cast<double>(255ull * sh->cnt));
      const uint32_t midx = 0xffffffffu - static_cast<uint32_t>(sh->maxkey & 0xffffffffu);
      const int my = static_cast<int>(midx / ww), mx = static_cast<int>(midx - static_cast<uint32_t>(my) * ww);
      int qx = static_cast<int>(mtb_rint(cx)), qy = static_cast<int>(mtb_rint(cy));
      qx = qx < 0 ? 0 : (qx > W - 1 ? W - 1 : qx);
      qy = qy < 0 ? 0 : (qy > H - 1 ? H - 1 : qy);
      uint32_t d2c = 0;
      {
        const int x = qx - wx0, y = qy - wy0;
        if (x >= 0 && x < ww && y >= 0 && y < wh && g[static_cast<size_t>(y) * ww + x])
          d2c = d2_at(g + static_cast<size_t>(y) * ww, ww, x);
      }
      const float max_val = mtb_fsqrt(static_cast<float>(maxd2));
      const float limit = static_cast<float>(mtb_dmul(static_cast<double>(max_val), 0.70));
      if (mtb_fsqrt(static_cast<float>(d2c)) < limit) {
        cx = static_cast<double>(wx0 + mx);
        cy = static_cast<double>(wy0 + my);
        R.moved |= MOVED_POLE;
      }
      const int px = static_cast<int>(mtb_rint(cx)), py = static_cast<int>(mtb_rint(cy));
      bool ok = px >= 0 && px < W && py >= 0 && py < H;
      if (ok) {
        const int x = px - wx0, y = py - wy0;
        ok = x >= 0 && x < ww && y >= 0 && y < wh && safe[static_cast<size_t>(y) * ww + x] != 0;
      }
      sh->cx = cx;
      sh->cy = cy;
      sh->px = px;
      sh->py = py;
      sh->flag = ok ? 0 : 1;
    }
  }
  MTB_SYNC();
  if (sh->flag < 0) return;

  // 3b. anchor outside the safe area: nearest safe pixel, float64 sqrt distance, first in raster order on ties
  if (sh->flag == 1) {
    const double cx = sh->cx, cy = sh->cy;
    const int n = ww * wh;
    unsigned long long best = ~0ull;
    for (int i = tid; i < n; i += nthr) {
      if (!safe[i]) continue;
      const int y = i / ww, x = i - y * ww;
      const double dy = mtb_dsub(static_cast<double>(wy0 + y), cy), dx = mtb_dsub(static_cast<double>(wx0 + x), cx);
      const unsigned long long k = mtb_double_bits(mtb_dsqrt(mtb_dadd(mtb_dmul(dy, dy), mtb_dmul(dx, dx))));
      if (k < best) best = k;
    }
    mtb_atomic_min64(&sh->minkey, best);
    MTB_SYNC();
    const unsigned long long want = sh->minkey;
    for (int i = tid; i < n; i += nthr) {
      if (!safe[i]) continue;
      const int y = i / ww, x = i - y * ww;
      const double dy = mtb_dsub(static_cast<double>(wy0 + y), cy), dx = mtb_dsub(static_cast<double>(wx0 + x), cx);
      if (mtb_double_bits(mtb_dsqrt(mtb_dadd(mtb_dmul(dy, dy), mtb_dmul(dx, dx)))) == want) {
        mtb_atomic_min(&sh->minidx, i);
        break;   // this thread's later pixels have larger raster indices
      }
    }
    MTB_SYNC();
    if (tid == 0) {
      const int y = sh->minidx / ww, x = sh->minidx - y * ww;
      sh->px = wx0 + x;
      sh->py = wy0 + y;
      sh->cx = static_cast<double>(sh->px);
      sh->cy = static_cast<double>(sh->py);
      R.moved |= MOVED_NEAREST;
    }
    MTB_SYNC();
  }

  // 4. ray casts along the anchor's row and column: distance to the nearest unsafe pixel, or to the image edge
  if (tid == 0) {
    sh->left = sh->px;
    sh->right = W - sh->px;
    sh->up = sh->py;
    sh->down = H - sh->py;
  }
  MTB_SYNC();
  {
    const int px = sh->px, py = sh->py;
    const int ry = py - wy0, rx = px - wx0;
    for (int x = tid; x < ww; x += nthr) {
      const int X = wx0 + x;
      if (X < 0 || X >= W || safe[static_cast<size_t>(ry) * ww + x]) continue;
      if (X < px) mtb_atomic_min(&sh->left, px - X);
      else mtb_atomic_min(&sh->right, X - px);
    }
    for (int y = tid; y < wh; y += nthr) {
      const int Y = wy0 + y;
      if (Y < 0 || Y >= H || safe[static_cast<size_t>(y) * ww + rx]) continue;
      if (Y < py) mtb_atomic_min(&sh->up, py - Y);
      else mtb_atomic_min(&sh->down, Y - py);
    }
  }
  MTB_SYNC();
  if (tid == 0) {
    int hw = sh->left < sh->right ? sh->left : sh->right;
    int hh = sh->up < sh->down ? sh->up : sh->down;
    hw = hw > 1 ? hw - 1 : hw;
    hh = hh > 1 ? hh - 1 : hh;
    const int bw = 2 * (hw > 0 ? hw : 0), bh = 2 * (hh > 0 ? hh : 0);
    R.anchor[0] = sh->px;
    R.anchor[1] = sh->py;
    R.cx = sh->cx;
    R.cy = sh->cy;
    if (bw <= 0 || bh <= 0) {
      R.status = ST_BAD_DIMS;
    } else {
      const int bx = static_cast<int>(mtb_rint(mtb_dsub(sh->cx, mtb_ddiv(static_cast<double>(bw), 2.0))));
      const int by = static_cast<int>(mtb_rint(mtb_dsub(sh->cy, mtb_ddiv(static_cast<double>(bh), 2.0))));
      R.box[0] = bx; R.box[1] = by; R.box[2] = bw; R.box[3] = bh;
      R.status = (bx >= 0 && by >= 0 && bx + bw <= W && by + bh <= H) ? ST_OK : ST_OUT_OF_BOUNDS;
    }
  }
}

}  // namespace mtbsafe
