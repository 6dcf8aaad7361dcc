// RT-DETRv2 glue kernels (the conjoined / fallback bubble detector of the reference, core/ml/rtdetr_adapter.py:61-113 ->
// transformers RTDetrV2ForObjectDetection): everything that is not a convolution / linear layer (those are tcgen05 conv
// plans), a LayerNorm or a dense attention (sam_kernels.cu).
//   * mtb_maxpool2d      : the ResNet stem's MaxPool2d(3, stride 2, pad 1) on NHWC hi/lo planes
//   * mtb_deform_attn    : RTDetrV2MultiscaleDeformableAttention's sampling core (`multi_scale_deformable_attention_v2`,
//                          method "default"): softmax over the levels x points logits of a head, bilinear grid_sample
//                          (align_corners=False, zero padding) of the projected value maps at
//                          ref_xy + offset * (1/n_points) * ref_wh * offset_scale, weighted sum.
//                          One warp per (query, head); a lane owns one of the head's 32 channels, so every corner fetch is
//                          one coalesced 64-byte row per plane.
#include <math.h>

#include <atomic>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

int sm_count_r() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

__global__ void maxpool2d_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int C,
                                 int Ho, int Wo, int k, int stride, int pad, int planes, long long ps_in, long long ps_out) {
  const int vec = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long p = i / vec;
    const int ox = static_cast<int>(p % Wo);
    p /= Wo;
    const int oy = static_cast<int>(p % Ho);
    const int n = static_cast<int>(p / Ho);
    float best[8];
    uint16_t bh[8], bl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bh[j] = 0xFF80;  // -inf in bf16 (a window that lies entirely in the padding cannot occur for pad < k)
      bl[j] = 0;
    }
    for (int dy = 0; dy < k; ++dy) {
      const int yy = oy * stride - pad + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int xx = ox * stride - pad + dx;
        if (xx < 0 || xx >= W) continue;
        const long long off = ((static_cast<long long>(n) * H + yy) * W + xx) * C + v * 8;
        const uint4 h4 = *reinterpret_cast<const uint4*>(x + off);
        uint4 l4 = make_uint4(0, 0, 0, 0);
        if (planes == 2) l4 = *reinterpret_cast<const uint4*>(x + ps_in + off);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint16_t hh = (j & 1) ? (hw[j >> 1] >> 16) : (hw[j >> 1] & 0xFFFF);
          const uint16_t ll = (j & 1) ? (lw[j >> 1] >> 16) : (lw[j >> 1] & 0xFFFF);
          const float val = bf16_to_f(hh) + bf16_to_f(ll);
          if (val > best[j]) {
            best[j] = val;
            bh[j] = hh;
            bl[j] = ll;
          }
        }
      }
    }
    const long long oo = ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + v * 8;
    uint4 oh, ol;
    oh.x = bh[0] | (static_cast<uint32_t>(bh[1]) << 16);
    oh.y = bh[2] | (static_cast<uint32_t>(bh[3]) << 16);
    oh.z = bh[4] | (static_cast<uint32_t>(bh[5]) << 16);
    oh.w = bh[6] | (static_cast<uint32_t>(bh[7]) << 16);
    ol.x = bl[0] | (static_cast<uint32_t>(bl[1]) << 16);
    ol.y = bl[2] | (static_cast<uint32_t>(bl[3]) << 16);
    ol.z = bl[4] | (static_cast<uint32_t>(bl[5]) << 16);
    ol.w = bl[6] | (static_cast<uint32_t>(bl[7]) << 16);
    *reinterpret_cast<uint4*>(y + oo) = oh;
    if (planes == 2) *reinterpret_cast<uint4*>(y + ps_out + oo) = ol;
  }
}

constexpr int kMaxLevels = 4;
constexpr int kMaxLP = 32;     // levels x points per head

struct DeformParams {
  int lh[kMaxLevels], lw[kMaxLevels], lstart[kMaxLevels];
  int n_levels, n_points, heads, Q, planes, ctotal;
  int off_stride, logit_stride;
  float offset_scale;
  long long v_ps, o_ps;
};

__device__ __forceinline__ float fetch_val(const uint16_t* v, long long ps, int planes, long long idx) {
  float r = bf16_to_f(v[idx]);
  if (planes == 2) r += bf16_to_f(v[ps + idx]);
  return r;
}

__global__ void deform_attn_kernel(const uint16_t* __restrict__ value, const float* __restrict__ offsets,
                                   const float* __restrict__ logits, const float* __restrict__ ref,
                                   uint16_t* __restrict__ out, const DeformParams P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.Q * P.heads) return;
  const int q = warp / P.heads, h = warp - q * P.heads;
  const int LP = P.n_levels * P.n_points;
  const float* lg = logits + static_cast<long long>(q) * P.logit_stride + h * LP;
  const float* of = offsets + static_cast<long long>(q) * P.off_stride + h * LP * 2;
  // softmax over the head's levels x points logits (fp32, like F.softmax)
  float w[kMaxLP];
  float m = -INFINITY;
  for (int i = 0; i < LP; ++i) {
    w[i] = lg[i];
    m = fmaxf(m, w[i]);
  }
  float s = 0.f;
  for (int i = 0; i < LP; ++i) {
    w[i] = expf(w[i] - m);
    s += w[i];
  }
  const float inv = 1.0f / s;
  const float rx = ref[q * 4 + 0], ry = ref[q * 4 + 1], rw = ref[q * 4 + 2], rh = ref[q * 4 + 3];
  const float pscale = 1.0f / static_cast<float>(P.n_points);
  float acc = 0.f;
  for (int l = 0; l < P.n_levels; ++l) {
    const int H = P.lh[l], W = P.lw[l];
    const long long base = static_cast<long long>(P.lstart[l]);
    for (int pnt = 0; pnt < P.n_points; ++pnt) {
      const int i = l * P.n_points + pnt;
      // sampling_locations = ref_xy + offset * n_points_scale * ref_wh * offset_scale; grid = 2*loc - 1;
      // grid_sample(align_corners=False): pixel = ((grid + 1) * size - 1) / 2
      const float lx = rx + of[2 * i] * pscale * rw * P.offset_scale;
      const float ly = ry + of[2 * i + 1] * pscale * rh * P.offset_scale;
      const float gx = 2.0f * lx - 1.0f, gy = 2.0f * ly - 1.0f;
      float ix = ((gx + 1.0f) * W - 1.0f) * 0.5f, iy = ((gy + 1.0f) * H - 1.0f) * 0.5f;
      // far-outside samples contribute zero either way; keep the float -> int conversion defined (NaN maps outside too)
      ix = (ix > -4.0f && ix < 1.0e6f) ? ix : -4.0f;
      iy = (iy > -4.0f && iy < 1.0e6f) ? iy : -4.0f;
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
      const float ax = ix - fx, ay = iy - fy;
      const float wnw = (1.0f - ax) * (1.0f - ay), wne = ax * (1.0f - ay), wsw = (1.0f - ax) * ay, wse = ax * ay;
      float sv = 0.f;
      const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
      const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
      const long long ch = static_cast<long long>(h) * 32 + lane;
      if (yin0 && xin0) sv += wnw * fetch_val(value, P.v_ps, P.planes, (base + static_cast<long long>(y0) * W + x0) * P.ctotal + ch);
      if (yin0 && xin1) sv += wne * fetch_val(value, P.v_ps, P.planes, (base + static_cast<long long>(y0) * W + x0 + 1) * P.ctotal + ch);
      if (yin1 && xin0) sv += wsw * fetch_val(value, P.v_ps, P.planes, (base + static_cast<long long>(y0 + 1) * W + x0) * P.ctotal + ch);
      if (yin1 && xin1) sv += wse * fetch_val(value, P.v_ps, P.planes, (base + static_cast<long long>(y0 + 1) * W + x0 + 1) * P.ctotal + ch);
      acc += sv * (w[i] * inv);
    }
  }
  uint16_t hi, lo;
  split_bf16(acc, hi, lo);
  const long long o = static_cast<long long>(q) * P.ctotal + h * 32 + lane;
  out[o] = hi;
  if (P.planes == 2) out[P.o_ps + o] = lo;
}

}  // namespace

extern "C" {

int mtb_maxpool2d(const void* x, void* y, int N, int H, int W, int C, int k, int stride, int pad, int planes, void* stream) {
  MTB_REQUIRE(x && y && C % 8 == 0 && k >= 1 && stride >= 1 && pad >= 0 && pad < k && (planes == 1 || planes == 2),
              "mtb_maxpool2d: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  MTB_REQUIRE(Ho > 0 && Wo > 0, "mtb_maxpool2d: empty output");
  const long long total = static_cast<long long>(N) * Ho * Wo * (C / 8);
  long long g = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count_r()) * 16;
  if (g > cap) g = cap;
  maxpool2d_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), N, H, W, C, Ho, Wo, k, stride, pad, planes,
      static_cast<long long>(N) * H * W * C, static_cast<long long>(N) * Ho * Wo * C);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_deform_attn(const void* value, long long value_plane_stride, int planes, int ctotal, int heads, int hd, int n_levels,
                    const int* level_h_w_start, int n_points, const float* offsets, int off_stride, const float* logits,
                    int logit_stride, const float* ref_cxcywh, int Q, float offset_scale, void* out,
                    long long out_plane_stride, void* stream) {
  MTB_REQUIRE(value && offsets && logits && ref_cxcywh && out && level_h_w_start, "mtb_deform_attn: null argument");
  MTB_REQUIRE(hd == 32 && heads * hd == ctotal && n_levels >= 1 && n_levels <= kMaxLevels && n_points >= 1 &&
                  n_levels * n_points <= kMaxLP && (planes == 1 || planes == 2) && Q >= 1,
              "mtb_deform_attn: unsupported geometry (head dim must be 32, levels*points <= 32)");
  DeformParams P;
  for (int l = 0; l < n_levels; ++l) {
    P.lh[l] = level_h_w_start[3 * l];
    P.lw[l] = level_h_w_start[3 * l + 1];
    P.lstart[l] = level_h_w_start[3 * l + 2];
  }
  P.n_levels = n_levels;
  P.n_points = n_points;
  P.heads = heads;
  P.Q = Q;
  P.planes = planes;
  P.ctotal = ctotal;
  P.off_stride = off_stride;
  P.logit_stride = logit_stride;
  P.offset_scale = offset_scale;
  P.v_ps = value_plane_stride;
  P.o_ps = out_plane_stride;
  const int warps = Q * heads;
  deform_attn_kernel<<<(warps + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(value), offsets, logits, ref_cxcywh, static_cast<uint16_t*>(out), P);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
