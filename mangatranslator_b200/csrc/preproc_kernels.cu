// Input pre-processing on the device, bit-exact with the CPU libraries the reference goes through:
//   * mtb_letterbox_u8 : ultralytics LetterBox = cv2.resize(INTER_LINEAR) on uint8 (OpenCV's 11-bit fixed-point
//     bilinear: horizontal pass with short coefficients, vertical pass `(((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2`)
//     + constant 114 border + BGR->RGB                       (reference call: core/image/detection.py:1338-1345)
//   * mtb_resize_aa_u8 : torchvision/ATen bilinear antialias resize of a uint8 image (Sam2ImageProcessorFast), i.e.
//     Pillow-style separable resampling with int16 weights and a uint8 intermediate
//                                                             (reference call: core/image/detection.py:494-495)
#include <math.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

// dst[(top+dy), (left+dx)] = cv2-bilinear(src); everything else = pad value
__global__ void letterbox_kernel(const uint8_t* __restrict__ src, int sh, int sw, int sc, uint8_t* __restrict__ dst,
                                 int dh, int dw, int top, int left, int nh, int nw, int pad_value, int swap_rb,
                                 const int* __restrict__ xofs, const short* __restrict__ xa,
                                 const int* __restrict__ yofs, const short* __restrict__ ya) {
  const long long total = static_cast<long long>(dh) * dw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int oy = static_cast<int>(i / dw), ox = static_cast<int>(i - static_cast<long long>(oy) * dw);
    uint8_t* o = dst + i * 3;
    const int ry = oy - top, rx = ox - left;
    if (ry < 0 || ry >= nh || rx < 0 || rx >= nw) {
      o[0] = o[1] = o[2] = static_cast<uint8_t>(pad_value);
      continue;
    }
    const int sx = xofs[rx], a0 = xa[2 * rx], a1 = xa[2 * rx + 1];
    const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
    const int sy0 = yofs[2 * ry], sy1 = yofs[2 * ry + 1], b0 = ya[2 * ry], b1 = ya[2 * ry + 1];
    const uint8_t* r0 = src + static_cast<long long>(sy0) * sw * sc;
    const uint8_t* r1 = src + static_cast<long long>(sy1) * sw * sc;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = r0[sx * sc + c] * a0 + r0[sx1 * sc + c] * a1;
      const int h1 = r1[sx * sc + c] * a0 + r1[sx1 * sc + c] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[swap_rb ? 2 - c : c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

__global__ void copy_pad_kernel(const uint8_t* __restrict__ src, int sh, int sw, int sc, uint8_t* __restrict__ dst, int dh,
                                int dw, int top, int left, int pad_value, int swap_rb) {
  const long long total = static_cast<long long>(dh) * dw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int oy = static_cast<int>(i / dw), ox = static_cast<int>(i - static_cast<long long>(oy) * dw);
    uint8_t* o = dst + i * 3;
    const int ry = oy - top, rx = ox - left;
    if (ry < 0 || ry >= sh || rx < 0 || rx >= sw) {
      o[0] = o[1] = o[2] = static_cast<uint8_t>(pad_value);
      continue;
    }
    const uint8_t* s = src + (static_cast<long long>(ry) * sw + rx) * sc;
    o[0] = s[swap_rb ? 2 : 0];
    o[1] = s[1];
    o[2] = s[swap_rb ? 0 : 2];
  }
}

// Pillow `background.paste(image, mask=alpha)` (ImagingPaste with an "L" mask): per channel
// (bg * (255 - a) + src * a + 128 + ((... + 128) >> 8)) >> 8  -- checked exhaustively against Pillow in tests/test_preproc.py
__global__ void flatten_alpha_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long npix, int bg0,
                                     int bg1, int bg2) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uchar4 p = reinterpret_cast<const uchar4*>(src)[i];
    const int a = p.w, bg[3] = {bg0, bg1, bg2}, c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int t = bg[k] * (255 - a) + c[k] * a + 128;
      dst[i * 3 + k] = static_cast<uint8_t>((t + (t >> 8)) >> 8);
    }
  }
}

int cv_round_f(float v) { return static_cast<int>(lrintf(v)); }  // round half to even (default FP mode)

short sat_short(int v) { return static_cast<short>(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

// separable antialiased resample of one axis: out[o] = clip8((half + sum_k w[o][k] * in[start[o] + k]) >> prec)
// WT = short (ATen int16 tables) or int (Pillow's 22-bit tables; 255 * sum|w| stays below 2^31 by construction)
// WT_TRANSPOSED: the table is laid out [k][n_out] so that the threads of a warp (adjacent output samples of the
// horizontal pass) read adjacent words; the vertical pass shares one table row per warp and keeps [n_out][k].
template <typename WT, bool WT_TRANSPOSED = false>
__global__ void resample_axis_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int out_h, int out_w,
                                     int in_w_stride /* pixels per src row */, int c_in, int c_out, int horizontal,
                                     const int* __restrict__ start, const int* __restrict__ len,
                                     const WT* __restrict__ wts, int kmax, int prec) {
  const long long total = static_cast<long long>(out_h) * out_w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int oy = static_cast<int>(i / out_w), ox = static_cast<int>(i - static_cast<long long>(oy) * out_w);
    const int o = horizontal ? ox : oy;
    const int s0 = start[o], n = len[o];
    const WT* w = WT_TRANSPOSED ? wts + o : wts + static_cast<long long>(o) * kmax;
    const long long wstep = WT_TRANSPOSED ? (horizontal ? out_w : out_h) : 1;
    int acc[3] = {1 << (prec - 1), 1 << (prec - 1), 1 << (prec - 1)};
    for (int k = 0; k < n; ++k) {
      const uint8_t* px = horizontal ? src + (static_cast<long long>(oy) * in_w_stride + s0 + k) * c_in
                                     : src + (static_cast<long long>(s0 + k) * in_w_stride + ox) * c_in;
      const int wk = w[k * wstep];
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += wk * px[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = acc[c] >> prec;
      dst[i * c_out + c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

// ATen `_compute_indices_int16_weights_aa` for the bilinear (triangle) filter with antialias
void aa_weights(int in_size, int out_size, std::vector<int>& start, std::vector<int>& len, std::vector<short>& w16,
                int& kmax, int& prec) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double support = scale >= 1.0 ? 1.0 * scale : 1.0;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  kmax = ksize;
  std::vector<double> wd(static_cast<size_t>(out_size) * ksize, 0.0);
  start.assign(out_size, 0);
  len.assign(out_size, 0);
  double wmax = 0.0;
  const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
  for (int i = 0; i < out_size; ++i) {
    const double center = scale * (i + 0.5);
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double total = 0.0;
    for (int j = 0; j < n; ++j) {
      double x = (j + xmin - center + 0.5) * invscale;
      if (x < 0) x = -x;
      const double v = x < 1.0 ? 1.0 - x : 0.0;
      wd[static_cast<size_t>(i) * ksize + j] = v;
      total += v;
    }
    for (int j = 0; j < n; ++j) {
      if (total != 0.0) wd[static_cast<size_t>(i) * ksize + j] /= total;
      if (wd[static_cast<size_t>(i) * ksize + j] > wmax) wmax = wd[static_cast<size_t>(i) * ksize + j];
    }
    start[i] = xmin;
    len[i] = n;
  }
  prec = 0;
  for (prec = 0; prec < 22; ++prec) {
    const int next = static_cast<int>(0.5 + wmax * (1 << (prec + 1)));
    if (next >= (1 << 15)) break;
  }
  w16.assign(static_cast<size_t>(out_size) * ksize, 0);
  for (int i = 0; i < out_size; ++i)
    for (int j = 0; j < len[i]; ++j) {
      const double v = wd[static_cast<size_t>(i) * ksize + j];
      w16[static_cast<size_t>(i) * ksize + j] =
          static_cast<short>(v < 0 ? static_cast<int>(-0.5 + v * (1 << prec)) : static_cast<int>(0.5 + v * (1 << prec)));
    }
}

// Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` (libImaging/Resample.c) for the LANCZOS filter (support 3) over
// the full-image box: double-precision windowed sinc, normalised per output sample, rounded to 22-bit integers.
double pil_sinc(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
double pil_lanczos(double x) { return (-3.0 <= x && x < 3.0) ? pil_sinc(x) * pil_sinc(x / 3) : 0.0; }

constexpr int kPilPrecisionBits = 32 - 8 - 2;

void lanczos_weights(int in_size, int out_size, std::vector<int>& start, std::vector<int>& len, std::vector<int>& w32,
                     int& ksize) {
  double filterscale, scale;
  filterscale = scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 3.0 * filterscale;
  ksize = static_cast<int>(ceil(support)) * 2 + 1;
  start.assign(out_size, 0);
  len.assign(out_size, 0);
  w32.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = pil_lanczos((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x];
      w32[static_cast<size_t>(xx) * ksize + x] =
          v < 0 ? static_cast<int>(-0.5 + v * (1 << kPilPrecisionBits)) : static_cast<int>(0.5 + v * (1 << kPilPrecisionBits));
    }
    start[xx] = xmin;
    len[xx] = xmax;
  }
}

// Coefficient tables depend only on the geometry: build + upload once, keep on the device (no per-call H2D / sync,
// which also makes the calls CUDA-graph capturable).
std::mutex g_tab_mu;
std::map<std::tuple<int, int, int, int, int>, void*> g_tab_cache;

void* cached_table(int kind, int a, int b, int c, int d, const std::vector<char>& bytes) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_tuple(kind, a, b, c, d);
  auto it = g_tab_cache.find(key);
  if (it != g_tab_cache.end()) return it->second;
  void* dev = nullptr;
  if (cudaMalloc(&dev, bytes.size()) != cudaSuccess) return nullptr;
  if (cudaMemcpy(dev, bytes.data(), bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_tab_cache[key] = dev;
  return dev;
}
bool table_cached(int kind, int a, int b, int c, int d, void** out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto it = g_tab_cache.find(std::make_tuple(kind, a, b, c, d));
  if (it == g_tab_cache.end()) return false;
  *out = it->second;
  return true;
}

int sm_count3() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

}  // namespace

extern "C" {

// host-only helper exported for the CPU tests: the int16 antialias weight tables of one axis
int mtb_aa_weights_host(int in_size, int out_size, int* start /* [out] */, int* len /* [out] */,
                        short* weights /* [out][kmax_cap] */, int kmax_cap, int* kmax, int* prec) {
  std::vector<int> s, l;
  std::vector<short> w;
  int k = 0, p = 0;
  aa_weights(in_size, out_size, s, l, w, k, p);
  MTB_REQUIRE(k <= kmax_cap, "mtb_aa_weights_host: kmax %d > capacity %d", k, kmax_cap);
  for (int i = 0; i < out_size; ++i) {
    start[i] = s[i];
    len[i] = l[i];
    for (int j = 0; j < k; ++j) weights[static_cast<size_t>(i) * kmax_cap + j] = w[static_cast<size_t>(i) * k + j];
  }
  *kmax = k;
  *prec = p;
  return 0;
}

// tables: device scratch of at least 3*(nw + nh) ints (used as int / short tables), filled by this call
int mtb_letterbox_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* dst, int dh, int dw, int top, int left, int nh,
                     int nw, int pad_value, int swap_rb, int* tables_dev, void* stream) {
  MTB_REQUIRE(src && dst && tables_dev && (sc == 3 || sc == 4), "mtb_letterbox_u8: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = sm_count3() * 8;
  if (nh == sh && nw == sw) {
    copy_pad_kernel<<<grid, 256, 0, st>>>(src, sh, sw, sc, dst, dh, dw, top, left, pad_value, swap_rb);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
  }
  // coefficient tables exactly as cv::resize builds them (resize.cpp, INTER_LINEAR, 8-bit), cached per geometry
  void* tab = nullptr;
  if (!table_cached(0, sh, sw, nh, nw, &tab)) {
    std::vector<int> xofs(nw), yofs(2 * nh);
    std::vector<short> xa(2 * nw), ya(2 * nh);
    const double scale_x = 1.0 / (static_cast<double>(nw) / sw), scale_y = 1.0 / (static_cast<double>(nh) / sh);
    for (int dx = 0; dx < nw; ++dx) {
      float fx = static_cast<float>((dx + 0.5) * scale_x - 0.5);
      int sx = static_cast<int>(floorf(fx));
      fx -= sx;
      if (sx < 0) {
        fx = 0;
        sx = 0;
      }
      if (sx >= sw - 1) {
        fx = 0;
        sx = sw - 1;
      }
      xofs[dx] = sx;
      xa[2 * dx] = sat_short(cv_round_f((1.f - fx) * 2048.f));
      xa[2 * dx + 1] = sat_short(cv_round_f(fx * 2048.f));
    }
    for (int dy = 0; dy < nh; ++dy) {
      float fy = static_cast<float>((dy + 0.5) * scale_y - 0.5);
      int sy = static_cast<int>(floorf(fy));
      fy -= sy;
      yofs[2 * dy] = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
      yofs[2 * dy + 1] = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
      ya[2 * dy] = sat_short(cv_round_f((1.f - fy) * 2048.f));
      ya[2 * dy + 1] = sat_short(cv_round_f(fy * 2048.f));
    }
    std::vector<char> bytes(sizeof(int) * (nw + 2 * nh) + sizeof(short) * (2 * nw + 2 * nh));
    char* w = bytes.data();
    memcpy(w, xofs.data(), sizeof(int) * nw);
    w += sizeof(int) * nw;
    memcpy(w, yofs.data(), sizeof(int) * 2 * nh);
    w += sizeof(int) * 2 * nh;
    memcpy(w, xa.data(), sizeof(short) * 2 * nw);
    w += sizeof(short) * 2 * nw;
    memcpy(w, ya.data(), sizeof(short) * 2 * nh);
    tab = cached_table(0, sh, sw, nh, nw, bytes);
    MTB_REQUIRE(tab != nullptr, "mtb_letterbox_u8: table upload failed");
  }
  int* d_xofs = static_cast<int*>(tab);
  int* d_yofs = d_xofs + nw;
  short* d_xa = reinterpret_cast<short*>(d_yofs + 2 * nh);
  short* d_ya = d_xa + 2 * nw;
  letterbox_kernel<<<grid, 256, 0, st>>>(src, sh, sw, sc, dst, dh, dw, top, left, nh, nw, pad_value, swap_rb, d_xofs,
                                         d_xa, d_yofs, d_ya);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// antialiased bilinear resize HxW -> oh x ow of a uint8 image (first 3 channels), ATen CPU uint8 semantics:
// horizontal pass first (into `tmp`, sh x ow x 3), then vertical.
int mtb_resize_aa_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* tmp, uint8_t* dst, int oh, int ow,
                     int* tables_dev, long long tables_ints, void* stream) {
  MTB_REQUIRE(src && tmp && dst && tables_dev, "mtb_resize_aa_u8: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int hk = static_cast<int>(ceil(sw >= ow ? static_cast<double>(sw) / ow : 1.0)) * 2 + 1;
  int vk = static_cast<int>(ceil(sh >= oh ? static_cast<double>(sh) / oh : 1.0)) * 2 + 1;
  int hp = 0, vp = 0;
  void* tab = nullptr;
  if (!table_cached(1, sh, sw, oh, ow, &tab)) {
    std::vector<int> hs, hl, vs, vl;
    std::vector<short> hw, vw;
    aa_weights(sw, ow, hs, hl, hw, hk, hp);
    aa_weights(sh, oh, vs, vl, vw, vk, vp);
    std::vector<char> bytes(sizeof(int) * (2 + 2 * ow + 2 * oh) + sizeof(short) * (hw.size() + vw.size() + 2));
    char* w = bytes.data();
    const int precs[2] = {hp, vp};
    memcpy(w, precs, sizeof(precs));
    w += sizeof(precs);
    memcpy(w, hs.data(), sizeof(int) * ow);
    w += sizeof(int) * ow;
    memcpy(w, hl.data(), sizeof(int) * ow);
    w += sizeof(int) * ow;
    memcpy(w, vs.data(), sizeof(int) * oh);
    w += sizeof(int) * oh;
    memcpy(w, vl.data(), sizeof(int) * oh);
    w += sizeof(int) * oh;
    memcpy(w, hw.data(), sizeof(short) * hw.size());
    w += sizeof(short) * (hw.size() + (hw.size() & 1));
    memcpy(w, vw.data(), sizeof(short) * vw.size());
    tab = cached_table(1, sh, sw, oh, ow, bytes);
    MTB_REQUIRE(tab != nullptr, "mtb_resize_aa_u8: table upload failed");
    std::lock_guard<std::mutex> lk(g_tab_mu);
    g_tab_cache[std::make_tuple(2, sh, sw, oh, ow)] = reinterpret_cast<void*>(static_cast<intptr_t>(hp * 64 + vp));
  }
  {
    void* pv = nullptr;
    table_cached(2, sh, sw, oh, ow, &pv);
    const intptr_t code = reinterpret_cast<intptr_t>(pv);
    hp = static_cast<int>(code / 64);
    vp = static_cast<int>(code % 64);
  }
  int* d_hs = static_cast<int*>(tab) + 2;
  int* d_hl = d_hs + ow;
  int* d_vs = d_hl + ow;
  int* d_vl = d_vs + oh;
  short* d_hw = reinterpret_cast<short*>(d_vl + oh);
  const size_t hw_n = static_cast<size_t>(ow) * hk;
  short* d_vw = d_hw + hw_n + (hw_n & 1);
  const int grid = sm_count3() * 8;
  const uint8_t* cur = src;
  int cur_w = sw, cur_c = sc;
  if (ow != sw) {
    resample_axis_kernel<short><<<grid, 256, 0, st>>>(cur, tmp, sh, ow, sw, sc, 3, 1, d_hs, d_hl, d_hw, hk, hp);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    cur = tmp;
    cur_w = ow;
    cur_c = 3;
  }
  if (oh != sh) {
    resample_axis_kernel<short><<<grid, 256, 0, st>>>(cur, dst, oh, ow, cur_w, cur_c, 3, 0, d_vs, d_vl, d_vw, vk, vp);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
  } else {
    copy_pad_kernel<<<grid, 256, 0, st>>>(cur, sh, cur_w, cur_c, dst, oh, ow, 0, 0, 0, 0);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
  }
  return 0;
}

// host-only helper for the CPU tests: Pillow's 22-bit LANCZOS coefficient table of one axis
int mtb_lanczos_weights_host(int in_size, int out_size, int* start /* [out] */, int* len /* [out] */,
                             int* weights /* [out][ksize_cap] */, int ksize_cap, int* ksize) {
  MTB_REQUIRE(in_size > 0 && out_size > 0 && start && len && weights && ksize, "mtb_lanczos_weights_host: bad arguments");
  std::vector<int> s, l, w;
  int k = 0;
  lanczos_weights(in_size, out_size, s, l, w, k);
  MTB_REQUIRE(k <= ksize_cap, "mtb_lanczos_weights_host: ksize %d > capacity %d", k, ksize_cap);
  for (int i = 0; i < out_size; ++i) {
    start[i] = s[i];
    len[i] = l[i];
    for (int j = 0; j < k; ++j) weights[static_cast<size_t>(i) * ksize_cap + j] = w[static_cast<size_t>(i) * k + j];
  }
  *ksize = k;
  return 0;
}

long long mtb_resize_lanczos_table_ints(int sh, int sw, int oh, int ow) {
  const double fx = sw > ow ? static_cast<double>(sw) / ow : 1.0, fy = sh > oh ? static_cast<double>(sh) / oh : 1.0;
  const long long kx = static_cast<long long>(ceil(3.0 * fx)) * 2 + 1, ky = static_cast<long long>(ceil(3.0 * fy)) * 2 + 1;
  return 2LL * ow + 2LL * oh + ow * kx + oh * ky;
}

// Coefficient tables of one geometry, written into caller-owned device scratch (layout: h start/len, v start/len,
// h weights TRANSPOSED [k][ow], v weights [oh][k]).  Computing them costs ~35k sin() calls for a page, so callers keep
// the filled scratch per geometry and pass `tables_ready = 1` on later calls.
int mtb_resize_lanczos_tables(int sh, int sw, int oh, int ow, int* tables_dev, long long tables_ints, void* stream) {
  MTB_REQUIRE(tables_dev && sh > 0 && sw > 0 && oh > 0 && ow > 0, "mtb_resize_lanczos_tables: bad arguments");
  MTB_REQUIRE(tables_ints >= mtb_resize_lanczos_table_ints(sh, sw, oh, ow), "mtb_resize_lanczos_tables: table scratch too small");
  std::vector<int> hs, hl, hw, vs, vl, vw;
  int hk = 0, vk = 0;
  lanczos_weights(sw, ow, hs, hl, hw, hk);
  lanczos_weights(sh, oh, vs, vl, vw, vk);
  std::vector<int> all;
  all.reserve(2 * ow + 2 * oh + hw.size() + vw.size());
  all.insert(all.end(), hs.begin(), hs.end());
  all.insert(all.end(), hl.begin(), hl.end());
  all.insert(all.end(), vs.begin(), vs.end());
  all.insert(all.end(), vl.begin(), vl.end());
  for (int k = 0; k < hk; ++k)
    for (int o = 0; o < ow; ++o) all.push_back(hw[static_cast<size_t>(o) * hk + k]);
  all.insert(all.end(), vw.begin(), vw.end());
  MTB_REQUIRE(static_cast<long long>(all.size()) <= tables_ints, "mtb_resize_lanczos_tables: table scratch too small");
  // pageable source: the copy is staged before the call returns, so `all` may go out of scope
  MTB_CUDA_OK(cudaMemcpyAsync(tables_dev, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice,
                              static_cast<cudaStream_t>(stream)));
  return 0;
}

// PIL `Image.resize((ow, oh), Image.LANCZOS)` of a uint8 image (first 3 channels): horizontal pass into `tmp`
// (sh x ow x 3, uint8 like Pillow's intermediate image), then vertical; a pass whose size does not change is skipped
// as in ImagingResample.  `tables_dev` is caller-owned scratch of mtb_resize_lanczos_table_ints() ints, filled here
// unless `tables_ready` says a previous call (or mtb_resize_lanczos_tables) already did for this geometry.
int mtb_resize_lanczos_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* tmp, uint8_t* dst, int oh, int ow,
                          int* tables_dev, long long tables_ints, int tables_ready, void* stream) {
  MTB_REQUIRE(src && dst && tables_dev && (sc == 3 || sc == 4) && sh > 0 && sw > 0 && oh > 0 && ow > 0,
              "mtb_resize_lanczos_u8: bad arguments");
  MTB_REQUIRE(tables_ints >= mtb_resize_lanczos_table_ints(sh, sw, oh, ow), "mtb_resize_lanczos_u8: table scratch too small");
  MTB_REQUIRE(tmp || ow == sw || oh == sh, "mtb_resize_lanczos_u8: two passes need the intermediate buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!tables_ready) {
    const int rc = mtb_resize_lanczos_tables(sh, sw, oh, ow, tables_dev, tables_ints, stream);
    if (rc != 0) return rc;
  }
  const double fx = sw > ow ? static_cast<double>(sw) / ow : 1.0, fy = sh > oh ? static_cast<double>(sh) / oh : 1.0;
  const int hk = static_cast<int>(ceil(3.0 * fx)) * 2 + 1, vk = static_cast<int>(ceil(3.0 * fy)) * 2 + 1;
  const int* d_hs = tables_dev;
  const int* d_hl = d_hs + ow;
  const int* d_vs = d_hl + ow;
  const int* d_vl = d_vs + oh;
  const int* d_hw = d_vl + oh;
  const int* d_vw = d_hw + static_cast<size_t>(ow) * hk;
  const int grid = sm_count3() * 8;
  const uint8_t* cur = src;
  int cur_w = sw, cur_c = sc;
  const bool need_h = ow != sw, need_v = oh != sh;
  if (need_h) {
    uint8_t* out = need_v ? tmp : dst;
    resample_axis_kernel<int, true><<<grid, 256, 0, st>>>(cur, out, sh, ow, sw, sc, 3, 1, d_hs, d_hl, d_hw, hk, kPilPrecisionBits);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    cur = out;
    cur_w = ow;
    cur_c = 3;
  }
  if (need_v) {
    resample_axis_kernel<int><<<grid, 256, 0, st>>>(cur, dst, oh, ow, cur_w, cur_c, 3, 0, d_vs, d_vl, d_vw, vk, kPilPrecisionBits);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
  } else if (!need_h) {
    copy_pad_kernel<<<grid, 256, 0, st>>>(cur, sh, cur_w, cur_c, dst, oh, ow, 0, 0, 0, 0);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
  }
  return 0;
}

// RGBA -> RGB over a constant background, Pillow paste arithmetic (core/image/image_utils.py:598-675
// convert_image_to_target_mode: transparency is flattened onto white before JPEG-bound processing)
int mtb_flatten_alpha_u8(const uint8_t* src /* H x W x 4, alpha last */, int H, int W, const int* bg3 /* host */,
                         uint8_t* dst /* H x W x 3 */, void* stream) {
  MTB_REQUIRE(src && dst && bg3 && H > 0 && W > 0, "mtb_flatten_alpha_u8: bad arguments");
  MTB_REQUIRE(reinterpret_cast<uintptr_t>(src) % 4 == 0, "mtb_flatten_alpha_u8: source must be 4-byte aligned");
  const long long npix = static_cast<long long>(H) * W;
  long long g = (npix + 255) / 256;
  const long long cap = static_cast<long long>(sm_count3()) * 16;
  if (g > cap) g = cap;
  flatten_alpha_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, npix, bg3[0], bg3[1], bg3[2]);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
