// Bubble cleaning: bit-exact re-implementation of the reference's per-bubble text-mask extraction
// (reference: core/image/cleaning.py:210-521 `process_single_bubble`, :524-1048 `clean_speech_bubbles`),
// redesigned for the GPU:
//   * one CTA per bubble ("job"), all bubbles of all pages of a batch in one launch;
//   * every intermediate mask is a BIT-PLANE of the bubble's crop window (32 pixels per word), so the ~15 full-frame
//     OpenCV passes per bubble of the reference become word-parallel AND/OR/shift sweeps over a few KB that stay in L1/L2;
//   * cv2.dilate/erode(MORPH_ELLIPSE)        -> row-run erosion/dilation of bit rows (an ellipse row is a run);
//   * cv2.distanceTransform(DIST_L2,5) >= t  -> erosion by the chamfer ball {N(dx,dy) < t} (a=1,b=1.4,c=2.1969 in 16.16
//                                               fixed point; the image border is not a source);
//   * findContours(RETR_EXTERNAL)+drawContours(FILLED) -> outer-background flood fill (4-conn) + run-based union-find
//                                               (8-conn) + Suzuki outer-border tracing for the polygon area/moments
//                                               (integer Green sums, exactly cv2.contourArea / cv2.moments);
//   * Otsu, medians, HSV saturation            -> 256-bin histograms + the reference's exact arithmetic.
//
// The same source compiles for the device (nvcc) and, for the unit tests only, as a sequential host emulation
// (-DMTB_HOST_EMUL, tests/host_emul) so the logic can be checked against cv2 in a container without a GPU.
#pragma once
#include <stdint.h>

#include "hd_emul.cuh"

namespace mtbclean {

enum Status : int {
  ST_OK = 0,
  ST_EMPTY_MASK = 1,     // reference: CleaningError("Empty mask ...") (cleaning.py:269-274)
  ST_NO_CONTOUR = 2,     // reference: CleaningError("Failed to process bubble mask") (cleaning.py:514)
  ST_WINDOW_TOO_SMALL = 3,
  ST_WORKSPACE_OVERFLOW = 4,
};

constexpr int kMaxSE = 63;        // reference scale_kernel clamp (core/scaling.py:64-96)
constexpr int kMaxBall = 65;      // roi_shrink clamp is 64 px (cleaning.py:629-636)
constexpr int kMaxNeighbors = 16;

struct Params {
  int thr_value;        // fixed threshold (cleaning.py:312-314)
  int use_otsu;         // first attempt uses Otsu (cleaning.py:299-310)
  int retry_otsu;       // on failure retry once with Otsu (cleaning.py:690-734)
  int kd, ke;           // dilation / erosion ellipse sizes (odd)
  int sed_hw[kMaxSE];   // half-width of each ellipse row, -1 = empty row
  int see_hw[kMaxSE];
  int ball_r;                       // chamfer ball row radius for the uniform shrink
  int ball_hw[2 * kMaxBall + 1];    // half-width per dy = -ball_r..ball_r (-1 = empty); all -1 => no shrink
  int jball_r;                      // chamfer ball for the junction-zone minimal shrink (cleaning.py:155-207)
  int jball_hw[2 * kMaxBall + 1];
  int junction_margin;
  double min_area;      // contour kept iff contourArea > min_area (cleaning.py:345-347)
  int margin;           // window margin the host applied around the detection bbox
};

struct Job {
  const uint8_t* img;   // page, interleaved BGR(A)
  long long img_pitch;
  int img_h, img_w, img_c;
  const uint8_t* mask;  // mask bytes (>0 = set), rectangle placed at (mask_x0, mask_y0) in page coordinates
  long long mask_pitch;
  int mask_x0, mask_y0, mask_w, mask_h;
  int wx0, wy0, cw, ch;  // crop window (inside the page)
  int bbox[4];           // detection bbox (x0,y0,x1,y1), used by the junction logic
  int n_neighbors;
  int neighbors[kMaxNeighbors][4];
  uint32_t* work;        // workspace: planes + run tables
  int max_runs;
  int page_index;
};

struct Result {
  int status;
  int used_otsu, otsu_thr;
  int is_black;
  int fill_bgr[3];
  int text_bbox[4];
  int has_text_color;
  int text_color[4];
  int n_components, n_valid;
  int final_start;       // raster index (page coords) of the chosen component's first pixel
  long long final_pixels;
  double final_area;
  unsigned long long gray_sum;
  unsigned int gray_cnt;
};

// plane indices inside the workspace
enum Plane : int { PL_M = 0, PL_ROI, PL_E, PL_T, PL_S, PL_F, PL_O, PL_G, PL_V, PL_FINAL, PL_TXT, PL_TXE, PL_COUNT };

struct Shared {
  unsigned long long gray_sum;
  unsigned int gray_cnt;
  int changed;
  int status;
  int n_runs;
  int n_comp;
  int n_valid;
  int best;           // component index of the largest valid contour
  int thr;
  int is_black;
  int bb[4];
  unsigned int hist[4][256];
  int mask_out_of_window;
};

struct Ctx {
  const Params* P;
  const Job* J;
  Shared* sh;
  int cwords;
  int plane_words;
  uint32_t* planes;
  // run tables
  int* row_off;     // [ch + 1]
  int* run_xs;      // [max_runs]
  int* run_xe;
  int* run_parent;
  int* comp_root;   // [max_runs] component -> root run
  // per-component results
  long long* comp_a00;
  double* comp_area;   // area of trace #1 (on F), then area of trace #2 (on V)
  int* comp_flag;      // bit0: valid
};

MTB_HD uint32_t* plane(const Ctx& c, int p) { return c.planes + static_cast<size_t>(p) * c.plane_words; }

#ifndef MTB_HOST_EMUL
__host__
#endif
MTB_HD size_t workspace_words(int cw, int ch, int max_runs) {
  const size_t cwords = (cw + 31) / 32;
  size_t w = static_cast<size_t>(PL_COUNT) * cwords * ch;
  w += (ch + 1);                 // row_off
  w += 4 * static_cast<size_t>(max_runs);  // xs, xe, parent, comp_root
  w += 2 * static_cast<size_t>(max_runs);  // comp_a00 (int64)
  w += 2 * static_cast<size_t>(max_runs);  // comp_area (double)
  w += static_cast<size_t>(max_runs);      // comp_flag
  return (w + 3) & ~static_cast<size_t>(3);
}

// ---- bit-plane access with the border policy -------------------------------------------------------------
// Outside the crop window a plane reads as `fill` if that side of the window coincides with the page border
// (cv2 treats out-of-image pixels as "ignored": erode pads with max, the chamfer transform has no border source),
// otherwise 0 (the host margin guarantees the true value there is 0).
MTB_HD uint32_t get_word(const Ctx& c, const uint32_t* pl, int y, int wx, bool fill) {
  const Job& J = *c.J;
  if (y < 0) return (fill && J.wy0 == 0) ? 0xFFFFFFFFu : 0u;
  if (y >= J.ch) return (fill && J.wy0 + J.ch == J.img_h) ? 0xFFFFFFFFu : 0u;
  if (wx < 0) return (fill && J.wx0 == 0) ? 0xFFFFFFFFu : 0u;
  const bool right_is_border = (J.wx0 + J.cw == J.img_w);
  if (wx >= c.cwords) return (fill && right_is_border) ? 0xFFFFFFFFu : 0u;
  uint32_t v = pl[static_cast<size_t>(y) * c.cwords + wx];
  if (wx == c.cwords - 1 && (J.cw & 31)) {
    const uint32_t valid = (1u << (J.cw & 31)) - 1u;
    v &= valid;
    if (fill && right_is_border) v |= ~valid;
  }
  return v;
}

// word whose bit i is the row bit at x = wx*32 + i + s
MTB_HD uint32_t row_shift(const Ctx& c, const uint32_t* pl, int y, int wx, int s, bool fill) {
  int q = s >> 5;  // floor division
  const int r = s & 31;
  const uint32_t lo = get_word(c, pl, y, wx + q, fill);
  if (r == 0) return lo;
  const uint32_t hi = get_word(c, pl, y, wx + q + 1, fill);
  return (lo >> r) | (hi << (32 - r));
}

MTB_HD uint32_t valid_mask(const Ctx& c, int wx) {
  if (wx == c.cwords - 1 && (c.J->cw & 31)) return (1u << (c.J->cw & 31)) - 1u;
  return 0xFFFFFFFFu;
}

MTB_HD int get_bit(const Ctx& c, const uint32_t* pl, int x, int y) {
  if (x < 0 || y < 0 || x >= c.J->cw || y >= c.J->ch) return 0;
  return (pl[static_cast<size_t>(y) * c.cwords + (x >> 5)] >> (x & 31)) & 1u;
}

// erosion (is_erode) / dilation of `src` by a structuring element given as per-row half-widths
MTB_HD void morph_rows(const Ctx& c, const uint32_t* src, uint32_t* dst, const int* hw, int radius, bool is_erode) {
  const int n = c.plane_words;
  for (int i = MTB_TID; i < n; i += MTB_NTHR) {
    const int y = i / c.cwords;
    const int wx = i - y * c.cwords;
    uint32_t acc = is_erode ? 0xFFFFFFFFu : 0u;
    for (int dy = -radius; dy <= radius; ++dy) {
      const int w = hw[dy + radius];
      if (w < 0) continue;
      for (int s = -w; s <= w; ++s) {
        const uint32_t v = row_shift(c, src, y + dy, wx, s, is_erode);
        acc = is_erode ? (acc & v) : (acc | v);
      }
      if (is_erode && acc == 0u) break;
    }
    dst[i] = acc & valid_mask(c, wx);
  }
}

MTB_HD int gray_at(const Job& J, int X, int Y) {
  const uint8_t* p = J.img + static_cast<long long>(Y) * J.img_pitch + static_cast<long long>(X) * J.img_c;
  // cv2.cvtColor(BGR2GRAY) 8-bit: (B*3735 + G*19235 + R*9798 + 16384) >> 15   
  return (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
}

// cv2 getThreshVal_Otsu_8u (imgproc/src/thresh.cpp) restated; hist over the ROI pixels (cleaning.py:300-303)
MTB_HD int otsu_threshold(const unsigned int* h, unsigned int total) {
  double mu = 0.0;
  const double scale = mtb_ddiv(1.0, static_cast<double>(total));
  for (int i = 0; i < 256; ++i) mu = mtb_dadd(mu, mtb_dmul(static_cast<double>(i), static_cast<double>(h[i])));
  mu = mtb_dmul(mu, scale);
  double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
  int max_val = 0;
  const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
  for (int i = 0; i < 256; ++i) {
    const double p_i = mtb_dmul(static_cast<double>(h[i]), scale);
    mu1 = mtb_dmul(mu1, q1);
    q1 = mtb_dadd(q1, p_i);
    const double q2 = mtb_dsub(1.0, q1);
    const double mn = q1 < q2 ? q1 : q2, mx = q1 < q2 ? q2 : q1;
    if (mn < eps || mx > mtb_dsub(1.0, eps)) continue;
    mu1 = mtb_ddiv(mtb_dadd(mu1, mtb_dmul(static_cast<double>(i), p_i)), q1);
    const double mu2 = mtb_ddiv(mtb_dsub(mu, mtb_dmul(q1, mu1)), q2);
    const double d = mtb_dsub(mu1, mu2);
    const double sigma = mtb_dmul(mtb_dmul(mtb_dmul(q1, q2), d), d);
    if (sigma > max_sigma) {
      max_sigma = sigma;
      max_val = i;
    }
  }
  return max_val;
}

// cv2 8-bit BGR2HSV saturation channel: s = (diff * sdiv_table[v] + 2048) >> 12, sdiv_table[v] = round(255*4096/v)
MTB_HD int hsv_saturation(int b, int g, int r) {
  int v = b > g ? b : g;
  v = v > r ? v : r;
  int mn = b < g ? b : g;
  mn = mn < r ? mn : r;
  const int diff = v - mn;
  if (v == 0) return 0;
  const double q = mtb_ddiv(static_cast<double>(255 << 12), static_cast<double>(v));
  // cvRound: round half to even
  double fl = floor(q);
  double fr = q - fl;
  int sd = static_cast<int>(fl);
  if (fr > 0.5 || (fr == 0.5 && (sd & 1))) sd += 1;
  return (diff * sd + (1 << 11)) >> 12;
}

struct TraceOut {
  long long a00, a10, a01;  // Green sums over the closed pixel chain (page coordinates)
  int minx, miny, maxx, maxy;
  int steps;
};

// Suzuki outer-border following exactly as cv2.findContours(RETR_EXTERNAL) walks it (contours.cpp icvFetchContour),
// accumulating the polygon sums of cv2.contourArea / cv2.moments instead of storing the points.
MTB_HD void trace_outer(const Ctx& c, const uint32_t* pl, int sx, int sy, TraceOut& o) {
  const int dxs[8] = {1, 1, 0, -1, -1, -1, 0, 1};
  const int dys[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  const long long ox = c.J->wx0, oy = c.J->wy0;
  o.a00 = o.a10 = o.a01 = 0;
  o.minx = o.maxx = sx;
  o.miny = o.maxy = sy;
  o.steps = 0;
  int s = 4;
  int i1x = sx, i1y = sy;
  bool single = true;
  do {
    s = (s - 1) & 7;
    i1x = sx + dxs[s];
    i1y = sy + dys[s];
    if (get_bit(c, pl, i1x, i1y)) {
      single = false;
      break;
    }
  } while (s != 4);
  if (single) return;
  int cx = sx, cy = sy;
  const int max_steps = 4 * (c.J->cw + 2) * (c.J->ch + 2);
  for (;;) {
    int nx = cx, ny = cy, sn = s;
    for (int k = 1; k <= 8; ++k) {
      sn = (s + k) & 7;
      nx = cx + dxs[sn];
      ny = cy + dys[sn];
      if (get_bit(c, pl, nx, ny)) break;
    }
    s = sn;
    // edge (cx,cy) -> (nx,ny) in page coordinates
    const long long x0 = cx + ox, y0 = cy + oy, x1 = nx + ox, y1 = ny + oy;
    const long long d = x0 * y1 - x1 * y0;
    o.a00 += d;
    o.a10 += d * (x0 + x1);
    o.a01 += d * (y0 + y1);
    if (nx < o.minx) o.minx = nx;
    if (nx > o.maxx) o.maxx = nx;
    if (ny < o.miny) o.miny = ny;
    if (ny > o.maxy) o.maxy = ny;
    ++o.steps;
    if ((nx == sx && ny == sy && cx == i1x && cy == i1y) || o.steps > max_steps) break;
    cx = nx;
    cy = ny;
    s = (s + 4) & 7;
  }
}

MTB_HD int find_root(int* parent, int i) {
  int r = i;
  while (parent[r] != r) r = parent[r];
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// The per-bubble pipeline.  Every thread of the CTA calls this; `sh` points to CTA-shared scratch.
// ---------------------------------------------------------------------------------------------------------
MTB_HD void clean_job(const Params& P, const Job& J, Result& R, Shared* sh) {
  Ctx c;
  c.P = &P;
  c.J = &J;
  c.sh = sh;
  c.cwords = (J.cw + 31) / 32;
  c.plane_words = c.cwords * J.ch;
  c.planes = J.work;
  {
    uint32_t* p = J.work + static_cast<size_t>(PL_COUNT) * c.plane_words;
    c.row_off = reinterpret_cast<int*>(p);
    p += J.ch + 1;
    c.run_xs = reinterpret_cast<int*>(p);
    p += J.max_runs;
    c.run_xe = reinterpret_cast<int*>(p);
    p += J.max_runs;
    c.run_parent = reinterpret_cast<int*>(p);
    p += J.max_runs;
    c.comp_root = reinterpret_cast<int*>(p);
    p += J.max_runs;
    // 8-byte alignment for the 64-bit tables
    size_t off = static_cast<size_t>(p - J.work);
    if (off & 1) ++p;
    c.comp_a00 = reinterpret_cast<long long*>(p);
    p += 2 * static_cast<size_t>(J.max_runs);
    c.comp_area = reinterpret_cast<double*>(p);
    p += 2 * static_cast<size_t>(J.max_runs);
    c.comp_flag = reinterpret_cast<int*>(p);
  }
  const int tid = MTB_TID, nthr = MTB_NTHR;
  uint32_t* M = plane(c, PL_M);
  uint32_t* ROI = plane(c, PL_ROI);
  uint32_t* E = plane(c, PL_E);
  uint32_t* T = plane(c, PL_T);
  uint32_t* S = plane(c, PL_S);
  uint32_t* F = plane(c, PL_F);
  uint32_t* O = plane(c, PL_O);
  uint32_t* G = plane(c, PL_G);
  uint32_t* V = plane(c, PL_V);
  uint32_t* FIN = plane(c, PL_FINAL);
  uint32_t* TXT = plane(c, PL_TXT);
  uint32_t* TXE = plane(c, PL_TXE);

  if (tid == 0) {
    sh->gray_sum = 0;
    sh->gray_cnt = 0;
    sh->status = ST_OK;
    sh->mask_out_of_window = 0;
  }
  MTB_SYNC();

  // ---- P0: M = (mask > 0) inside the window; gray statistics under the mask (cleaning.py:258,268-278) ----
  for (int i = tid; i < c.plane_words; i += nthr) {
    const int y = i / c.cwords, wx = i - y * c.cwords;
    const int Y = J.wy0 + y;
    const int my = Y - J.mask_y0;
    uint32_t bits = 0;
    unsigned int gs = 0, gc = 0;
    if (my >= 0 && my < J.mask_h) {
      const uint8_t* mrow = J.mask + static_cast<long long>(my) * J.mask_pitch;
      for (int b = 0; b < 32; ++b) {
        const int x = wx * 32 + b;
        if (x >= J.cw) break;
        const int X = J.wx0 + x;
        const int mx = X - J.mask_x0;
        if (mx >= 0 && mx < J.mask_w && mrow[mx] > 0) {
          bits |= 1u << b;
          gs += static_cast<unsigned int>(gray_at(J, X, Y));
          ++gc;
        }
      }
    }
    M[i] = bits;
    if (gc) {
      mtb_atomic_add(&sh->gray_sum, static_cast<unsigned long long>(gs));
      mtb_atomic_add(&sh->gray_cnt, gc);
    }
  }
  MTB_SYNC();
  if (sh->gray_cnt == 0) {
    if (tid == 0) {
      R.status = ST_EMPTY_MASK;
      R.gray_sum = 0;
      R.gray_cnt = 0;
    }
    return;
  }
  // the mask must keep `margin` pixels from every window side that is not the page border
  for (int i = tid; i < c.plane_words; i += nthr) {
    const uint32_t v = M[i];
    if (!v) continue;
    const int y = i / c.cwords, wx = i - y * c.cwords;
    const int m = P.margin;
    bool bad = false;
    if (J.wy0 > 0 && y < m) bad = true;
    if (J.wy0 + J.ch < J.img_h && y >= J.ch - m) bad = true;
    const int xlo = wx * 32 + mtb_ffs(v) - 1;
    int xhi = wx * 32 + 31;
    while (!((v >> (xhi & 31)) & 1u)) --xhi;
    if (J.wx0 > 0 && xlo < m) bad = true;
    if (J.wx0 + J.cw < J.img_w && xhi >= J.cw - m) bad = true;
    if (bad) sh->mask_out_of_window = 1;
  }
  MTB_SYNC();
  if (sh->mask_out_of_window) {
    if (tid == 0) R.status = ST_WINDOW_TOO_SMALL;
    return;
  }
  const int is_black = (sh->gray_sum < 128ull * sh->gray_cnt) ? 1 : 0;  // mean < GRAYSCALE_MIDPOINT (cleaning.py:276-277)

  // ---- P1/P2: ROI = dilate(M, ellipse kd) (cleaning.py:288); E = erode(M, ellipse ke) (cleaning.py:336-338) ----
  morph_rows(c, M, ROI, P.sed_hw, P.kd / 2, false);
  morph_rows(c, M, E, P.see_hw, P.ke / 2, true);
  MTB_SYNC();

  // ---- P4: S = (chamfer distance transform of ROI >= shrink) (cleaning.py:318-333, 155-207) ----
  {
    bool any = false;
    for (int k = 0; k <= 2 * P.ball_r; ++k) any = any || (P.ball_hw[k] >= 0);
    if (any) {
      morph_rows(c, ROI, S, P.ball_hw, P.ball_r, true);
    } else {
      // roi_shrink == 0: `dist >= 0` holds for EVERY pixel of the page (cleaning.py:330); the host gives such jobs
      // the whole page as window so that "everything" is representable
      for (int i = tid; i < c.plane_words; i += nthr) S[i] = valid_mask(c, i % c.cwords);
    }
    MTB_SYNC();
    if (J.n_neighbors > 0 && any) {
      // junction zones: restore pixels with dist >= junction_min_shrink inside the margin box around the
      // intersection with each conjoined neighbour (O plane used as scratch)
      bool anyj = false;
      for (int k = 0; k <= 2 * P.jball_r; ++k) anyj = anyj || (P.jball_hw[k] >= 0);
      if (anyj) {
        morph_rows(c, ROI, O, P.jball_hw, P.jball_r, true);
      } else {
        for (int i = tid; i < c.plane_words; i += nthr) O[i] = ROI[i];
      }
      MTB_SYNC();
      const int am = P.junction_margin;
      const int x1 = J.bbox[0], y1 = J.bbox[1], x2 = J.bbox[2], y2 = J.bbox[3];
      for (int nb = 0; nb < J.n_neighbors; ++nb) {
        const int ox1 = J.neighbors[nb][0], oy1 = J.neighbors[nb][1], ox2 = J.neighbors[nb][2], oy2 = J.neighbors[nb][3];
        if (x1 - am > ox2 || ox1 - am > x2 || y1 - am > oy2 || oy1 - am > y2) continue;
        int zx1 = (x1 > ox1 ? x1 : ox1) - am;
        int zy1 = (y1 > oy1 ? y1 : oy1) - am;
        int zx2 = (x2 < ox2 ? x2 : ox2) + am;
        int zy2 = (y2 < oy2 ? y2 : oy2) + am;
        if (zx1 < 0) zx1 = 0;
        if (zy1 < 0) zy1 = 0;
        if (zx2 > J.img_w) zx2 = J.img_w;
        if (zy2 > J.img_h) zy2 = J.img_h;
        if (zx2 <= zx1 || zy2 <= zy1) continue;
        for (int i = tid; i < c.plane_words; i += nthr) {
          const int y = i / c.cwords, wx = i - y * c.cwords;
          const int Y = J.wy0 + y;
          if (Y < zy1 || Y >= zy2) continue;
          uint32_t zone = 0;
          for (int b = 0; b < 32; ++b) {
            const int X = J.wx0 + wx * 32 + b;
            if (X >= zx1 && X < zx2) zone |= 1u << b;
          }
          S[i] |= O[i] & zone;
        }
        MTB_SYNC();
      }
    }
  }

  int attempt_otsu = P.use_otsu;
  for (int attempt = 0; attempt < 2; ++attempt) {
    // ---- P3: threshold inside the ROI (cleaning.py:289-316) ----
    int thr = P.thr_value;
    if (attempt_otsu) {
      for (int i = tid; i < 256; i += nthr) sh->hist[0][i] = 0;
      MTB_SYNC();
      for (int i = tid; i < c.plane_words; i += nthr) {
        uint32_t v = ROI[i];
        const int y = i / c.cwords, wx = i - y * c.cwords;
        while (v) {
          const int b = mtb_ffs(v) - 1;
          v &= v - 1;
          int g = gray_at(J, J.wx0 + wx * 32 + b, J.wy0 + y);
          if (is_black) g = 255 - g;
          mtb_atomic_add(&sh->hist[0][g], 1u);
        }
      }
      MTB_SYNC();
      if (tid == 0) {
        unsigned int total = 0;
        for (int i = 0; i < 256; ++i) total += sh->hist[0][i];
        sh->thr = otsu_threshold(sh->hist[0], total);
      }
      MTB_SYNC();
      thr = sh->thr;
    }
    for (int i = tid; i < c.plane_words; i += nthr) {
      uint32_t v = ROI[i];
      const int y = i / c.cwords, wx = i - y * c.cwords;
      uint32_t t = 0;
      while (v) {
        const int b = mtb_ffs(v) - 1;
        v &= v - 1;
        int g = gray_at(J, J.wx0 + wx * 32 + b, J.wy0 + y);
        if (is_black) g = 255 - g;
        if (g > thr) t |= 1u << b;
      }
      T[i] = t;
      F[i] = t & S[i];  // thresholded_roi after both ANDs (cleaning.py:316,333)
    }
    MTB_SYNC();

    // ---- P6: outer background O = 4-connected flood of ~F from the window frame; G = ~O = top-level components
    //          with their holes filled (== findContours(RETR_EXTERNAL) + drawContours(FILLED)) ----
    for (int i = tid; i < c.plane_words; i += nthr) {
      const int y = i / c.cwords, wx = i - y * c.cwords;
      const uint32_t pass = ~F[i] & valid_mask(c, wx);
      uint32_t seed = 0;
      if (y == 0 || y == J.ch - 1) seed = pass;
      if (wx == 0) seed |= pass & 1u;
      if (wx == c.cwords - 1) seed |= pass & (1u << ((J.cw - 1) & 31));
      O[i] = seed;
    }
    MTB_SYNC();
    for (int iter = 0; iter < J.cw * J.ch + 2; ++iter) {
      if (tid == 0) sh->changed = 0;
      MTB_SYNC();
      int local_changed = 0;
      for (int i = tid; i < c.plane_words; i += nthr) {
        const int y = i / c.cwords, wx = i - y * c.cwords;
        const uint32_t pass = ~F[i] & valid_mask(c, wx);
        const uint32_t cur = O[i];
        uint32_t g = cur;
        if (y > 0) g |= O[i - c.cwords];
        if (y < J.ch - 1) g |= O[i + c.cwords];
        if (wx > 0) g |= O[i - 1] >> 31;
        if (wx < c.cwords - 1) g |= O[i + 1] << 31;
        g &= pass;
        // in-word occluded fill, both directions
        uint32_t pro = pass, gen = g;
        gen |= pro & (gen << 1); pro &= pro << 1;
        gen |= pro & (gen << 2); pro &= pro << 2;
        gen |= pro & (gen << 4); pro &= pro << 4;
        gen |= pro & (gen << 8); pro &= pro << 8;
        gen |= pro & (gen << 16);
        pro = pass;
        gen |= pro & (gen >> 1); pro &= pro >> 1;
        gen |= pro & (gen >> 2); pro &= pro >> 2;
        gen |= pro & (gen >> 4); pro &= pro >> 4;
        gen |= pro & (gen >> 8); pro &= pro >> 8;
        gen |= pro & (gen >> 16);
        if (gen != cur) {
          O[i] = gen;
          local_changed = 1;
        }
      }
      if (local_changed) sh->changed = 1;
      MTB_SYNC();
      const int ch_flag = sh->changed;
      MTB_SYNC();
      if (!ch_flag) break;
    }
    for (int i = tid; i < c.plane_words; i += nthr) {
      const int wx = i % c.cwords;
      G[i] = ~O[i] & valid_mask(c, wx);
    }
    MTB_SYNC();

    // ---- P8: runs of G + union-find (8-connectivity) ----
    for (int y = tid; y < J.ch; y += nthr) {
      int cnt = 0;
      uint32_t prev_bit = 0;
      for (int wx = 0; wx < c.cwords; ++wx) {
        const uint32_t v = G[static_cast<size_t>(y) * c.cwords + wx];
        const uint32_t starts = v & ~((v << 1) | prev_bit);
        cnt += mtb_popc(starts);
        prev_bit = v >> 31;
      }
      c.row_off[y + 1] = cnt;
    }
    MTB_SYNC();
    if (tid == 0) {
      c.row_off[0] = 0;
      for (int y = 0; y < J.ch; ++y) c.row_off[y + 1] += c.row_off[y];
      sh->n_runs = c.row_off[J.ch];
      if (sh->n_runs > J.max_runs) sh->status = ST_WORKSPACE_OVERFLOW;
    }
    MTB_SYNC();
    if (sh->status == ST_WORKSPACE_OVERFLOW) {
      if (tid == 0) R.status = ST_WORKSPACE_OVERFLOW;
      return;
    }
    const int n_runs = sh->n_runs;
    for (int y = tid; y < J.ch; y += nthr) {
      int k = c.row_off[y];
      int x = 0;
      bool in_run = false;
      for (int wx = 0; wx < c.cwords; ++wx) {
        uint32_t v = G[static_cast<size_t>(y) * c.cwords + wx];
        for (int b = 0; b < 32; ++b, ++x) {
          const bool bit = (v >> b) & 1u;
          if (bit && !in_run) {
            c.run_xs[k] = x;
            in_run = true;
          } else if (!bit && in_run) {
            c.run_xe[k] = x - 1;
            ++k;
            in_run = false;
          }
        }
      }
      if (in_run) {
        c.run_xe[k] = J.cw - 1;
        ++k;
      }
    }
    for (int i = tid; i < n_runs; i += nthr) c.run_parent[i] = i;
    MTB_SYNC();
    for (int iter = 0; iter < n_runs + 2; ++iter) {
      if (tid == 0) sh->changed = 0;
      MTB_SYNC();
      int local_changed = 0;
      for (int y = 1 + tid; y < J.ch; y += nthr) {
        int a = c.row_off[y], ae = c.row_off[y + 1];
        int b = c.row_off[y - 1];
        const int be = c.row_off[y];
        while (a < ae && b < be) {
          // 8-connectivity: runs touch if [xs-1, xe+1] overlaps
          if (c.run_xe[b] < c.run_xs[a] - 1) {
            ++b;
            continue;
          }
          if (c.run_xs[b] > c.run_xe[a] + 1) {
            ++a;
            continue;
          }
          const int ra = find_root(c.run_parent, a), rb = find_root(c.run_parent, b);
          if (ra != rb) {
            const int lo = ra < rb ? ra : rb, hi = ra < rb ? rb : ra;
            mtb_atomic_min(&c.run_parent[hi], lo);
            local_changed = 1;
          }
          if (c.run_xe[b] < c.run_xe[a]) ++b; else ++a;
        }
      }
      if (local_changed) sh->changed = 1;
      MTB_SYNC();
      const int ch_flag = sh->changed;
      MTB_SYNC();
      if (!ch_flag) break;
    }
    for (int i = tid; i < n_runs; i += nthr) c.run_parent[i] = find_root(c.run_parent, i);
    MTB_SYNC();
    // component list = root runs, in raster order of their first pixel (a root is the smallest run index)
    if (tid == 0) {
      int n = 0;
      for (int i = 0; i < n_runs; ++i)
        if (c.run_parent[i] == i) c.comp_root[n++] = i;
      sh->n_comp = n;
      sh->n_valid = 0;
      sh->best = -1;
    }
    MTB_SYNC();
    const int n_comp = sh->n_comp;

    // ---- P9: trace #1 on F: contourArea / moments centroid gate (cleaning.py:343-358) ----
    for (int k = tid; k < n_comp; k += nthr) {
      const int root = c.comp_root[k];
      // row of the root run
      int lo = 0, hi = J.ch;  // binary search row_off
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (c.row_off[mid] <= root) lo = mid; else hi = mid;
      }
      const int sy = lo, sx = c.run_xs[root];
      TraceOut t;
      trace_outer(c, F, sx, sy, t);
      const double a00 = static_cast<double>(t.a00);
      const double area = fabs(mtb_dmul(a00, 0.5));
      int valid = 0;
      if (area > P.min_area && fabs(a00) > 1.1920928955078125e-07) {
        const double m00 = t.a00 > 0 ? mtb_dmul(a00, 0.5) : mtb_dmul(a00, -0.5);
        const double s6 = t.a00 > 0 ? 0.16666666666666666666666666666667 : -0.16666666666666666666666666666667;
        const double m10 = mtb_dmul(static_cast<double>(t.a10), s6);
        const double m01 = mtb_dmul(static_cast<double>(t.a01), s6);
        if (m00 != 0.0) {
          const double qx = mtb_ddiv(m10, m00), qy = mtb_ddiv(m01, m00);
          const long long cxp = static_cast<long long>(qx), cyp = static_cast<long long>(qy);  // int(): truncation
          if (cxp >= 0 && cxp < J.img_w && cyp >= 0 && cyp < J.img_h) {
            const int ex = static_cast<int>(cxp) - J.wx0, ey = static_cast<int>(cyp) - J.wy0;
            if (get_bit(c, E, ex, ey)) valid = 1;
          }
        }
      }
      c.comp_flag[k] = valid;
      c.comp_a00[k] = t.a00;
      c.comp_area[k] = area;
      if (valid) mtb_atomic_add(&sh->n_valid, 1);
    }
    MTB_SYNC();

    if (sh->n_valid > 0) {
      // ---- P10: V = filled valid components (cleaning.py:367-370); map run -> component validity ----
      for (int i = tid; i < c.plane_words; i += nthr) V[i] = 0;
      MTB_SYNC();
      // mark valid roots in run_parent-indexed flag: reuse run_xs? keep a per-run lookup through comp index search
      for (int y = tid; y < J.ch; y += nthr) {
        for (int r = c.row_off[y]; r < c.row_off[y + 1]; ++r) {
          const int root = c.run_parent[r];
          // binary search component index of this root
          int lo = 0, hi = n_comp - 1;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c.comp_root[mid] < root) lo = mid + 1; else hi = mid;
          }
          if (!c.comp_flag[lo]) continue;
          for (int x = c.run_xs[r]; x <= c.run_xe[r]; ++x)
            V[static_cast<size_t>(y) * c.cwords + (x >> 5)] |= 1u << (x & 31);
        }
      }
      MTB_SYNC();
      // ---- trace #2 on V: the reference re-contours the validated mask and keeps the largest contour by
      //      contourArea; max() returns the first maximum in findContours order = reverse raster order of the
      //      start pixels (cleaning.py:373-377) ----
      for (int k = tid; k < n_comp; k += nthr) {
        if (!c.comp_flag[k]) continue;
        const int root = c.comp_root[k];
        int lo = 0, hi = J.ch;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (c.row_off[mid] <= root) lo = mid; else hi = mid;
        }
        TraceOut t;
        trace_outer(c, V, c.run_xs[root], lo, t);
        c.comp_area[k] = fabs(mtb_dmul(static_cast<double>(t.a00), 0.5));
      }
      MTB_SYNC();
      if (tid == 0) {
        int best = -1;
        double best_area = -1.0;
        for (int k = n_comp - 1; k >= 0; --k) {  // findContours order: last-found first
          if (!c.comp_flag[k]) continue;
          if (c.comp_area[k] > best_area) {
            best_area = c.comp_area[k];
            best = k;
          }
        }
        sh->best = best;
      }
      MTB_SYNC();
      break;  // success
    }
    // failure: retry once with Otsu (cleaning.py:690-734) unless already Otsu
    if (attempt_otsu || !P.retry_otsu) break;
    attempt_otsu = 1;
    MTB_SYNC();
  }

  if (sh->n_valid == 0 || sh->best < 0) {
    if (tid == 0) {
      R.status = ST_NO_CONTOUR;
      R.is_black = is_black;
      R.used_otsu = attempt_otsu;
      R.otsu_thr = attempt_otsu ? sh->thr : -1;
      R.n_components = sh->n_comp;
      R.n_valid = 0;
      R.gray_sum = sh->gray_sum;
      R.gray_cnt = sh->gray_cnt;
    }
    return;
  }

  // ---- P11: final mask = filled largest contour; text_bbox = its bounding rect (cleaning.py:377-383) ----
  const int best = sh->best;
  const int best_root = c.comp_root[best];
  if (tid == 0) {
    sh->bb[0] = J.cw;
    sh->bb[1] = J.ch;
    sh->bb[2] = -1;
    sh->bb[3] = -1;
    sh->gray_cnt = 0;  // reused: pixel count of the final mask
  }
  for (int i = tid; i < c.plane_words; i += nthr) FIN[i] = 0;
  MTB_SYNC();
  for (int y = tid; y < J.ch; y += nthr) {
    unsigned int cnt = 0;
    int xmin = J.cw, xmax = -1;
    for (int r = c.row_off[y]; r < c.row_off[y + 1]; ++r) {
      if (c.run_parent[r] != best_root) continue;
      for (int x = c.run_xs[r]; x <= c.run_xe[r]; ++x)
        FIN[static_cast<size_t>(y) * c.cwords + (x >> 5)] |= 1u << (x & 31);
      cnt += c.run_xe[r] - c.run_xs[r] + 1;
      if (c.run_xs[r] < xmin) xmin = c.run_xs[r];
      if (c.run_xe[r] > xmax) xmax = c.run_xe[r];
    }
    if (cnt) {
      mtb_atomic_add(&sh->gray_cnt, cnt);
      mtb_atomic_min(&sh->bb[0], xmin);
      mtb_atomic_max(&sh->bb[2], xmax);
      mtb_atomic_min(&sh->bb[1], y);
      mtb_atomic_max(&sh->bb[3], y);
    }
  }
  MTB_SYNC();

  // ---- P12: text colour (cleaning.py:472-503): median BGR of erode3x3(~thresholded & shrunk) ----
  for (int i = tid; i < c.plane_words; i += nthr) {
    const int wx = i % c.cwords;
    TXT[i] = ~F[i] & S[i] & valid_mask(c, wx);
  }
  MTB_SYNC();
  {
    const int hw3[3] = {1, 1, 1};
    morph_rows(c, TXT, TXE, hw3, 1, true);
  }
  if (tid == 0) sh->changed = 0;
  MTB_SYNC();
  {
    int any = 0;
    for (int i = tid; i < c.plane_words; i += nthr) any |= (TXE[i] != 0);
    if (any) sh->changed = 1;
  }
  MTB_SYNC();
  const uint32_t* sample = sh->changed ? TXE : TXT;  // fallback if erosion obliterates thin text
  for (int i = tid; i < 4 * 256; i += nthr) sh->hist[i >> 8][i & 255] = 0;
  MTB_SYNC();
  for (int i = tid; i < c.plane_words; i += nthr) {
    uint32_t v = sample[i];
    const int y = i / c.cwords, wx = i - y * c.cwords;
    while (v) {
      const int b = mtb_ffs(v) - 1;
      v &= v - 1;
      const uint8_t* px = J.img + static_cast<long long>(J.wy0 + y) * J.img_pitch +
                          static_cast<long long>(J.wx0 + wx * 32 + b) * J.img_c;
      for (int ch = 0; ch < J.img_c; ++ch) mtb_atomic_add(&sh->hist[ch][px[ch]], 1u);
    }
  }
  MTB_SYNC();
  if (tid == 0) {
    R.status = ST_OK;
    R.is_black = is_black;
    R.used_otsu = attempt_otsu;
    R.otsu_thr = attempt_otsu ? sh->thr : -1;
    R.fill_bgr[0] = R.fill_bgr[1] = R.fill_bgr[2] = is_black ? 0 : 255;
    R.text_bbox[0] = J.wx0 + sh->bb[0];
    R.text_bbox[1] = J.wy0 + sh->bb[1];
    R.text_bbox[2] = J.wx0 + sh->bb[2] + 1;
    R.text_bbox[3] = J.wy0 + sh->bb[3] + 1;
    R.n_components = sh->n_comp;
    R.n_valid = sh->n_valid;
    R.final_pixels = sh->gray_cnt;
    R.final_area = c.comp_area[best];
    {
      int lo = 0, hi = J.ch;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (c.row_off[mid] <= best_root) lo = mid; else hi = mid;
      }
      R.final_start = (J.wy0 + lo) * J.img_w + J.wx0 + c.run_xs[best_root];
    }
    R.gray_sum = sh->gray_sum;
    unsigned int total = 0;
    for (int i = 0; i < 256; ++i) total += sh->hist[0][i];
    R.has_text_color = 0;
    R.text_color[0] = R.text_color[1] = R.text_color[2] = R.text_color[3] = 0;
    if (total > 0) {
      int med[4] = {0, 0, 0, 0};
      for (int ch = 0; ch < J.img_c; ++ch) {
        // np.median: odd n -> middle; even n -> (a+b)/2 in float64, then .astype(int) truncates
        const unsigned int k1 = (total - 1) / 2, k2 = total / 2;
        unsigned int acc = 0;
        int v1 = -1, v2 = -1;
        for (int i = 0; i < 256; ++i) {
          acc += sh->hist[ch][i];
          if (v1 < 0 && acc > k1) v1 = i;
          if (v2 < 0 && acc > k2) {
            v2 = i;
            break;
          }
        }
        med[ch] = (v1 + v2) / 2;
      }
      const int sat = hsv_saturation(med[0], med[1], med[2]);
      R.has_text_color = 1;
      if (sat < 25) {
        const int tc = is_black ? 255 : 0;  // luminance of the fill colour >= 128 -> black text (cleaning.py:492-501)
        R.text_color[0] = R.text_color[1] = R.text_color[2] = tc;
        R.text_color[3] = -1;
      } else {
        for (int ch = 0; ch < 4; ++ch) R.text_color[ch] = ch < J.img_c ? med[ch] : -1;
      }
    }
  }
}

}  // namespace mtbclean
