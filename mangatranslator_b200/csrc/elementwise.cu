// HBM-bound glue kernels around the tensor-core convolutions: image <-> plane conversion, RCAN channel attention
// (global average pool finish + squeeze/excite MLP + gate + residual), nearest upsample, concat, max-pool.
// All are pure streaming kernels: 16-byte vector loads/stores, grid = multiple of the SM count, one pass over HBM.
#include <atomic>

#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

inline int grid_for(long long work_items, int block, int sms) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sms) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

// u8 interleaved image (Cimg = 3 or 4) -> hi/lo planes NHWC, C = cpad (channels >= 3 zero)
__global__ void image_to_planes_kernel(const uint8_t* __restrict__ img, int H, int W, int cimg, int swap_rb, float mul,
                                       float s0, float s1, float s2, uint16_t* __restrict__ out, long long plane_stride,
                                       int cpad, int planes) {
  const long long npix = static_cast<long long>(H) * W;
  const int vec_per_pix = cpad / 8;  // 16-byte vectors per pixel per plane
  const long long total = npix * vec_per_pix;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / vec_per_pix;
    const int v = static_cast<int>(i - pix * vec_per_pix);
    uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
    if (v == 0) {
      const uint8_t* p = img + pix * cimg;
      float c0 = p[swap_rb ? 2 : 0], c1 = p[1], c2 = p[swap_rb ? 0 : 2];
      c0 = c0 * mul - s0;
      c1 = c1 * mul - s1;
      c2 = c2 * mul - s2;
      uint16_t h0, l0, h1, l1, h2, l2;
      split_bf16(c0, h0, l0);
      split_bf16(c1, h1, l1);
      split_bf16(c2, h2, l2);
      hi.x = h0 | (static_cast<uint32_t>(h1) << 16);
      hi.y = h2;
      lo.x = l0 | (static_cast<uint32_t>(l1) << 16);
      lo.y = l2;
    }
    reinterpret_cast<uint4*>(out + pix * cpad)[v] = hi;
    if (planes == 2) reinterpret_cast<uint4*>(out + plane_stride + pix * cpad)[v] = lo;
  }
}

// PixelUnshuffle(d) folded into the same conversion (RCAN "_PU" variants): out[yb][xb][c*d*d + dy*d + dx] =
// img[yb*d+dy][xb*d+dx][c]; rows/columns past the image edge (H or W not a multiple of d) are reflect-padded
__global__ void image_to_planes_unshuffle_kernel(const uint8_t* __restrict__ img, int H, int W, int cimg, int swap_rb,
                                                 float mul, float s0, float s1, float s2, int d,
                                                 uint16_t* __restrict__ out, long long plane_stride, int cpad, int planes) {
  const int Hb = (H + d - 1) / d, Wb = (W + d - 1) / d, dd = d * d;
  const int vec_per_pix = cpad / 8;
  const long long total = static_cast<long long>(Hb) * Wb * vec_per_pix;
  const float sub[3] = {s0, s1, s2};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / vec_per_pix;
    const int v = static_cast<int>(i - pix * vec_per_pix);
    const int yb = static_cast<int>(pix / Wb), xb = static_cast<int>(pix - static_cast<long long>(yb) * Wb);
    uint16_t h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = v * 8 + e;
      h[e] = l[e] = 0;
      if (ch < 3 * dd) {
        const int c = ch / dd, r = ch - c * dd;
        int y = yb * d + r / d, x = xb * d + r % d;
        if (y >= H) y = 2 * (H - 1) - y;
        if (x >= W) x = 2 * (W - 1) - x;
        y = y < 0 ? 0 : y;
        x = x < 0 ? 0 : x;
        const float val = img[(static_cast<long long>(y) * W + x) * cimg + (swap_rb ? 2 - c : c)] * mul - sub[c];
        split_bf16(val, h[e], l[e]);
      }
    }
    uint4 hi, lo;
    hi.x = h[0] | (static_cast<uint32_t>(h[1]) << 16);
    hi.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
    hi.z = h[4] | (static_cast<uint32_t>(h[5]) << 16);
    hi.w = h[6] | (static_cast<uint32_t>(h[7]) << 16);
    lo.x = l[0] | (static_cast<uint32_t>(l[1]) << 16);
    lo.y = l[2] | (static_cast<uint32_t>(l[3]) << 16);
    lo.z = l[4] | (static_cast<uint32_t>(l[5]) << 16);
    lo.w = l[6] | (static_cast<uint32_t>(l[7]) << 16);
    reinterpret_cast<uint4*>(out + pix * cpad)[v] = hi;
    if (planes == 2) reinterpret_cast<uint4*>(out + plane_stride + pix * cpad)[v] = lo;
  }
}

// finish the global average pool from the conv epilogue's per-warp partial sums and run the squeeze/excite MLP
// one block per image; C <= 256, R <= 64
__global__ void ca_scale_kernel(const float* __restrict__ sums, int parts, int C, float inv_hw,
                                const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2, int R,
                                float* __restrict__ scale) {
  __shared__ float mean[256];
  __shared__ float hid[64];
  __shared__ float part[8][256];
  const int n = blockIdx.x;
  const float* s = sums + static_cast<long long>(n) * parts * C;
  // deterministic two-level reduction: 8 row-groups of threads, fixed order inside each
  const int c = threadIdx.x % C;
  const int g = threadIdx.x / C;       // blockDim = 8*C (<= 1024 when C <= 128) or C
  const int groups = blockDim.x / C;
  float acc = 0.f;
  for (int p = g; p < parts; p += groups) acc += s[static_cast<long long>(p) * C + c];
  part[g][c] = acc;
  __syncthreads();
  if (g == 0) {
    float t = 0.f;
    for (int k = 0; k < groups; ++k) t += part[k][c];
    mean[c] = t * inv_hw;
  }
  __syncthreads();
  if (threadIdx.x < R) {
    float h = b1 ? b1[threadIdx.x] : 0.f;
    for (int k = 0; k < C; ++k) h += w1[threadIdx.x * C + k] * mean[k];
    hid[threadIdx.x] = fmaxf(h, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float o = b2 ? b2[threadIdx.x] : 0.f;
    for (int k = 0; k < R; ++k) o += w2[threadIdx.x * R + k] * hid[k];
    scale[n * C + threadIdx.x] = 1.0f / (1.0f + expf(-o));
  }
}


// RCAB channel-attention gate computed BEFORE the block's second convolution runs.
// The gate needs mean_pixels(conv2(u) + b2).  A zero-padded 3x3 convolution is linear, so that mean follows from sums of
// its INPUT: for tap (dy,dx) the sum over all output pixels of u[oy+dy][ox+dx] is the total minus the border row /
// column the shifted window never reaches (plus the doubly removed corner).  conv1's epilogue already delivers the total
// per channel; this kernel adds the four border lines, forms mean_y2 = b2 + W2 . S / HW in fp32 and runs the
// squeeze/excite MLP.  conv2's epilogue can then write x + gate*(conv2(u)+b2) directly and the separate
// read-scale-add pass over the page (1.2 GB per block) disappears.   One block of 1024 threads, C = 64.
__global__ void __launch_bounds__(1024, 1)
rcan_gate_kernel(const float* __restrict__ sums, int parts, const float* __restrict__ border,
                 const uint16_t* __restrict__ u, long long plane_stride,
                 int planes, int H, int W, const float* __restrict__ wconv /* [64][64][3][3] */,
                 const float* __restrict__ bconv, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, int R, float* __restrict__ scale,
                 float lo_inv /* planes == 3 (fp16c byte planes): scale of the e5m2 residual plane */,
                 long long* __restrict__ fixed /* optional [5][64]: totals + border lines in 2^-20 fixed point, accumulated
                                                  by the conv epilogue with integer atomics; consumed and zeroed here */) {
  constexpr int C = 64;
  __shared__ float red[32][C];                      // per-warp partials (border lines) / quarter sums (total)
  __shared__ float wide[64][C];                     // per-row-group partials of the two partial-row reductions
  __shared__ float tot[C], line[4][C], corner[4][C], shifted[9][C], mean[C], hid[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // second conv's weights for this thread's slice of the mean (16 threads per output channel, 36 terms each): issued
  // first so the L2 round trip overlaps everything below
  const int co = tid >> 4, part = tid & 15;
  float wreg[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) wreg[i] = __ldg(wconv + co * 576 + part + 16 * i);

  if (border == nullptr) {
    // border lines of u, one per group of 8 warps: 0 = row 0, 1 = row H-1, 2 = column 0, 3 = column W-1.
    // 8 lanes cover the 64 channels of a pixel (16-byte loads), a warp takes 4 pixels per step.
    {
      const int b = warp >> 3, v = lane & 7, slot = ((warp & 7) << 2) | (lane >> 3);   // 32 pixel slots per line
      const int len = b < 2 ? W : H;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const long long base = b == 1 ? static_cast<long long>(H - 1) * W : b == 3 ? W - 1 : 0;
      const long long step = b < 2 ? 1 : W;           // pixels between consecutive elements of the line
      const uint4 zero4 = make_uint4(0, 0, 0, 0);
  #pragma unroll 4
      for (int i = slot; i < len; i += 32) {
        const uint16_t* px = u + (base + i * step) * C + v * 8;
        const uint4 qh = *reinterpret_cast<const uint4*>(px);
        const uint4 ql = planes == 2 ? *reinterpret_cast<const uint4*>(px + plane_stride) : zero4;
        const uint32_t wh[4] = {qh.x, qh.y, qh.z, qh.w}, wl[4] = {ql.x, ql.y, ql.z, ql.w};
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += bf16_to_f(wh[j] & 0xFFFF) + bf16_to_f(wl[j] & 0xFFFF);
          acc[2 * j + 1] += bf16_to_f(wh[j] >> 16) + bf16_to_f(wl[j] >> 16);
        }
      }
  #pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
      }
      if (lane < 8) {
  #pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][v * 8 + j] = acc[j];
      }
    }
  }
  __syncthreads();
  if (tid < 4 * C) {
    const int b = tid >> 6, c = tid & 63;
    float t = 0.f;
    if (border == nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[b * 8 + k][c];
      line[b][c] = t;
    }
    // corners: 0 = (0,0), 1 = (0,W-1), 2 = (H-1,0), 3 = (H-1,W-1)
    const long long pix = (b & 2 ? static_cast<long long>(H - 1) * W : 0) + (b & 1 ? W - 1 : 0);
    float cv = 0.f;
    if (planes == 3) {
      // fp16c: fp16 plane (c >> 5) + e5m2 residual plane 2, 64 bytes per pixel each (plane_stride = bytes per plane)
      const uint8_t* ub = reinterpret_cast<const uint8_t*>(u);
      const __half h = *reinterpret_cast<const __half*>(ub + (c >> 5) * plane_stride + pix * 64 + (c & 31) * 2);
      const __half_raw lr = __nv_cvt_fp8_to_halfraw(ub[2 * plane_stride + pix * 64 + c], __NV_E5M2);
      cv = __half2float(h) + __half2float(*reinterpret_cast<const __half*>(&lr)) * lo_inv;
    } else {
      for (int pl = 0; pl < planes; ++pl) cv += bf16_to_f(u[pl * plane_stride + pix * C + c]);
    }
    corner[b][c] = cv;
  }
  __syncthreads();
  // Totals and border lines from the conv epilogue's partial rows.  Both reductions walk a few hundred L2-resident rows;
  // a plain `acc += row[p]` loop serialises one L2 round trip per row (the 148-iteration border loop alone cost ~30 us),
  // so every thread keeps four independent 16-byte loads in flight and the groups are folded through shared memory in a
  // fixed order (deterministic).
  if (fixed != nullptr) {
    if (tid < 5 * C) {
      const float v = static_cast<float>(static_cast<double>(fixed[tid]) * (1.0 / 1048576.0));
      fixed[tid] = 0;
      if (tid < C) tot[tid] = v;
      else line[(tid >> 6) - 1][tid & 63] = v;
    }
    __syncthreads();
  }
  if (fixed == nullptr) {  // total per channel: 16 float4 lanes x 64 row groups
    const int c4 = tid & 15, g = tid >> 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    const float4* src = reinterpret_cast<const float4*>(sums) + c4;
    int p = g;
    for (; p + 192 < parts; p += 256) {
      const float4 v0 = src[static_cast<long long>(p) * 16], v1 = src[static_cast<long long>(p + 64) * 16];
      const float4 v2 = src[static_cast<long long>(p + 128) * 16], v3 = src[static_cast<long long>(p + 192) * 16];
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; p < parts; p += 64) {
      const float4 v0 = src[static_cast<long long>(p) * 16];
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    float* dst = &wide[g][c4 * 4];
    dst[0] = (a0.x + a1.x) + (a2.x + a3.x);
    dst[1] = (a0.y + a1.y) + (a2.y + a3.y);
    dst[2] = (a0.z + a1.z) + (a2.z + a3.z);
    dst[3] = (a0.w + a1.w) + (a2.w + a3.w);
  }
  __syncthreads();
  if (fixed == nullptr && tid < 4 * C) {   // 64 groups -> 4 quarter sums per channel -> total
    const int c = tid & 63, qd = tid >> 6;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += wide[qd * 16 + k][c];
    red[qd][c] = t;
  }
  __syncthreads();
  if (fixed == nullptr && tid < C) tot[tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
  __syncthreads();
  if (border != nullptr && fixed == nullptr) {
    // the four border lines arrive as partial rows from the same conv epilogue: [parts][4 lines][64 channels];
    // 16 float4 lanes x 4 lines x 16 row groups
    const int c4 = tid & 15, b = (tid >> 4) & 3, g = tid >> 6;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    const float4* src = reinterpret_cast<const float4*>(border) + b * 16 + c4;
    int p = g;
    for (; p + 48 < parts; p += 64) {
      const float4 v0 = src[static_cast<long long>(p) * 64], v1 = src[static_cast<long long>(p + 16) * 64];
      const float4 v2 = src[static_cast<long long>(p + 32) * 64], v3 = src[static_cast<long long>(p + 48) * 64];
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; p < parts; p += 16) {
      const float4 v0 = src[static_cast<long long>(p) * 64];
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    float* dst = &wide[g * 4 + b][c4 * 4];
    dst[0] = (a0.x + a1.x) + (a2.x + a3.x);
    dst[1] = (a0.y + a1.y) + (a2.y + a3.y);
    dst[2] = (a0.z + a1.z) + (a2.z + a3.z);
    dst[3] = (a0.w + a1.w) + (a2.w + a3.w);
    __syncthreads();
    if (tid < 4 * C) {
      const int b2 = tid >> 6, c2 = tid & 63;
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += wide[k * 4 + b2][c2];
      line[b2][c2] = t;
    }
    __syncthreads();
  }
  if (tid < 9 * C) {
    const int tap = tid >> 6, c = tid & 63;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    float sft = tot[c];
    if (dy == 1) sft -= line[0][c];
    if (dy == -1) sft -= line[1][c];
    if (dx == 1) sft -= line[2][c];
    if (dx == -1) sft -= line[3][c];
    if (dy != 0 && dx != 0) sft += corner[(dy == -1 ? 2 : 0) + (dx == -1 ? 1 : 0)][c];
    shifted[tap][c] = sft;
  }
  __syncthreads();
  {  // mean of conv2's output
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 36; ++i) {
      const int t = part + 16 * i;
      const int ci = t / 9, tap = t - ci * 9;
      acc += wreg[i] * shifted[tap][ci];
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (part == 0) mean[co] = acc / (static_cast<float>(H) * static_cast<float>(W)) + (bconv ? bconv[co] : 0.f);
  }
  __syncthreads();
  if (tid < R) {
    float h = b1 ? b1[tid] : 0.f;
    for (int k = 0; k < C; ++k) h += w1[tid * C + k] * mean[k];
    hid[tid] = fmaxf(h, 0.f);
  }
  __syncthreads();
  if (tid < C) {
    float o = b2 ? b2[tid] : 0.f;
    for (int k = 0; k < R; ++k) o += w2[tid * R + k] * hid[k];
    scale[tid] = 1.0f / (1.0f + expf(-o));
  }
}

// y = x + t * scale[n][c]   (all hi/lo planes, NHWC, C multiple of 8)
__global__ void scale_residual_kernel(const uint16_t* __restrict__ t, const uint16_t* __restrict__ x,
                                      const float* __restrict__ scale, uint16_t* __restrict__ y,
                                      long long plane_stride, long long pix_per_image, int N, int C, int planes) {
  const int vec_per_pix = C / 8;
  const long long total = pix_per_image * N * vec_per_pix;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / vec_per_pix;
    const int v = static_cast<int>(i - pix * vec_per_pix);
    const int n = static_cast<int>(pix / pix_per_image);
    const long long off = pix * C + v * 8;
    const uint4 th = *reinterpret_cast<const uint4*>(t + off);
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const uint4 tl = planes == 2 ? *reinterpret_cast<const uint4*>(t + plane_stride + off) : zero4;
    const uint4 xh = *reinterpret_cast<const uint4*>(x + off);
    const uint4 xl = planes == 2 ? *reinterpret_cast<const uint4*>(x + plane_stride + off) : zero4;
    const uint32_t thw[4] = {th.x, th.y, th.z, th.w}, tlw[4] = {tl.x, tl.y, tl.z, tl.w};
    const uint32_t xhw[4] = {xh.x, xh.y, xh.z, xh.w}, xlw[4] = {xl.x, xl.y, xl.z, xl.w};
    const float* sc = scale + n * C + v * 8;
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t0 = bf16_to_f(thw[j] & 0xFFFF) + bf16_to_f(tlw[j] & 0xFFFF);
      const float t1 = bf16_to_f(thw[j] >> 16) + bf16_to_f(tlw[j] >> 16);
      const float x0 = bf16_to_f(xhw[j] & 0xFFFF) + bf16_to_f(xlw[j] & 0xFFFF);
      const float x1 = bf16_to_f(xhw[j] >> 16) + bf16_to_f(xlw[j] >> 16);
      const float y0 = x0 + t0 * sc[2 * j];
      const float y1 = x1 + t1 * sc[2 * j + 1];
      uint16_t h0, l0, h1, l1;
      split_bf16(y0, h0, l0);
      split_bf16(y1, h1, l1);
      oh[j] = h0 | (static_cast<uint32_t>(h1) << 16);
      ol[j] = l0 | (static_cast<uint32_t>(l1) << 16);
    }
    *reinterpret_cast<uint4*>(y + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(y + plane_stride + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// fp32 NHWC (C = cpad, first 3 channels used) -> (v + add[c]) * mul, clamp [0,1], *255, truncate -> u8 HxWx3;
// optionally also the float value before quantisation
__global__ void f32_to_u8_kernel(const float* __restrict__ in, long long npix, int cpad, float a0, float a1, float a2,
                                 float mul, uint8_t* __restrict__ out, float* __restrict__ out_f) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float* p = in + i * cpad;
    const float add[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (p[c] + add[c]) * mul;
      if (out_f) out_f[i * 3 + c] = v;
      const float q = fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f;
      out[i * 3 + c] = static_cast<uint8_t>(q);  // truncation like .astype(np.uint8) (image_utils.py:363-365)
    }
  }
}

// same as f32_to_u8_kernel for the top-left Hout x Wout window of a Hin x Win tensor (the "_PU" models compute on
// a frame padded to a multiple of the unshuffle factor)
__global__ void f32_to_u8_crop_kernel(const float* __restrict__ in, int Win, int cpad, int Hout, int Wout, float a0,
                                      float a1, float a2, float mul, uint8_t* __restrict__ out, float* __restrict__ out_f) {
  const long long npix = static_cast<long long>(Hout) * Wout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / Wout), x = static_cast<int>(i - static_cast<long long>(y) * Wout);
    const float* p = in + (static_cast<long long>(y) * Win + x) * cpad;
    const float add[3] = {a0, a1, a2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (p[c] + add[c]) * mul;
      if (out_f) out_f[i * 3 + c] = v;
      const float q = fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f;
      out[i * 3 + c] = static_cast<uint8_t>(q);
    }
  }
}

}  // namespace

extern "C" {

int mtb_image_to_planes(const uint8_t* img, int H, int W, int cimg, int swap_rb, float mul, const float* sub3,
                        void* planes_out, int cpad, int planes, void* stream) {
  MTB_REQUIRE(img && planes_out && (cimg == 3 || cimg == 4) && cpad % 8 == 0 && cpad >= 8 && (planes == 1 || planes == 2),
              "mtb_image_to_planes: bad arguments");
  const long long total = static_cast<long long>(H) * W * (cpad / 8);
  const float s0 = sub3 ? sub3[0] : 0.f, s1 = sub3 ? sub3[1] : 0.f, s2 = sub3 ? sub3[2] : 0.f;
  image_to_planes_kernel<<<grid_for(total, 256, sm_count()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, H, W, cimg, swap_rb, mul, s0, s1, s2, static_cast<uint16_t*>(planes_out),
      static_cast<long long>(H) * W * cpad, cpad, planes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_image_to_planes_unshuffle(const uint8_t* img, int H, int W, int cimg, int swap_rb, float mul, const float* sub3,
                                  int d, void* planes_out, int cpad, int planes, void* stream) {
  MTB_REQUIRE(img && planes_out && (cimg == 3 || cimg == 4) && cpad % 8 == 0 && (planes == 1 || planes == 2) && d >= 1 &&
                  3 * d * d <= cpad && H >= 1 && W >= 1,
              "mtb_image_to_planes_unshuffle: bad arguments");
  MTB_REQUIRE((H % d == 0 || H > d) && (W % d == 0 || W > d), "mtb_image_to_planes_unshuffle: image smaller than the reflect pad");
  const int Hb = (H + d - 1) / d, Wb = (W + d - 1) / d;
  const long long total = static_cast<long long>(Hb) * Wb * (cpad / 8);
  const float s0 = sub3 ? sub3[0] : 0.f, s1 = sub3 ? sub3[1] : 0.f, s2 = sub3 ? sub3[2] : 0.f;
  image_to_planes_unshuffle_kernel<<<grid_for(total, 256, sm_count()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, H, W, cimg, swap_rb, mul, s0, s1, s2, d, static_cast<uint16_t*>(planes_out),
      static_cast<long long>(Hb) * Wb * cpad, cpad, planes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_f32_to_u8_crop(const float* in, int Hin, int Win, int cpad, int Hout, int Wout, const float* add3, float mul,
                       uint8_t* out, float* out_f, void* stream) {
  MTB_REQUIRE(in && out && cpad >= 3 && Hout >= 1 && Wout >= 1 && Hout <= Hin && Wout <= Win, "mtb_f32_to_u8_crop: bad arguments");
  const float a0 = add3 ? add3[0] : 0.f, a1 = add3 ? add3[1] : 0.f, a2 = add3 ? add3[2] : 0.f;
  f32_to_u8_crop_kernel<<<grid_for(static_cast<long long>(Hout) * Wout, 256, sm_count()), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(in, Win, cpad, Hout, Wout, a0, a1, a2, mul, out, out_f);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_ca_scale(const float* sums, int n_images, int parts_per_image, int C, float inv_hw, const float* w1,
                 const float* b1, const float* w2, const float* b2, int R, float* scale_out, void* stream) {
  MTB_REQUIRE(sums && w1 && w2 && scale_out, "mtb_ca_scale: null argument");
  MTB_REQUIRE(C <= 256 && R <= 64 && C >= R, "mtb_ca_scale: C<=256, R<=64 required (C=%d R=%d)", C, R);
  int groups = 1024 / C;
  if (groups > 8) groups = 8;
  if (groups < 1) groups = 1;
  ca_scale_kernel<<<n_images, groups * C, 0, static_cast<cudaStream_t>(stream)>>>(sums, parts_per_image, C, inv_hw, w1,
                                                                                b1, w2, b2, R, scale_out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_rcan_gate(const float* sums, int parts, const float* border_sums, const void* u, int planes, int H, int W,
                  const float* conv_w, const float* conv_b, const float* w1, const float* b1, const float* w2,
                  const float* b2, int R, float* scale_out, void* stream) {
  MTB_REQUIRE(sums && u && conv_w && w1 && w2 && scale_out, "mtb_rcan_gate: null argument");
  MTB_REQUIRE((planes == 1 || planes == 2) && H > 0 && W > 0 && R > 0 && R <= 64 && parts > 0,
              "mtb_rcan_gate: bad arguments");
  rcan_gate_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, parts, border_sums, static_cast<const uint16_t*>(u), static_cast<long long>(H) * W * 64, planes, H, W, conv_w, conv_b, w1,
      b1, w2, b2, R, scale_out, 1.0f, nullptr);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_rcan_gate_fp16c(const float* sums, int parts, const float* border_sums, long long* fixed_sums, const void* u,
                        int lo_shift, int H, int W, const float* conv_w, const float* conv_b, const float* w1,
                        const float* b1, const float* w2, const float* b2, int R, float* scale_out, void* stream) {
  MTB_REQUIRE(u && conv_w && w1 && w2 && scale_out, "mtb_rcan_gate_fp16c: null argument");
  MTB_REQUIRE(fixed_sums || (sums && border_sums && parts > 0), "mtb_rcan_gate_fp16c: needs fixed_sums or sums + border_sums rows");
  MTB_REQUIRE(H > 0 && W > 0 && R > 0 && R <= 64, "mtb_rcan_gate_fp16c: bad arguments");
  rcan_gate_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, parts, border_sums, static_cast<const uint16_t*>(u), static_cast<long long>(H) * W * 64, 3, H, W, conv_w, conv_b,
      w1, b1, w2, b2, R, scale_out, ldexpf(1.0f, -lo_shift), fixed_sums);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_scale_residual(const void* t, const void* x, const float* scale, void* y, long long pix_per_image, int N, int C,
                       int planes, void* stream) {
  MTB_REQUIRE(t && x && scale && y && C % 8 == 0 && (planes == 1 || planes == 2), "mtb_scale_residual: bad arguments");
  const long long total = pix_per_image * N * (C / 8);
  scale_residual_kernel<<<grid_for(total, 256, sm_count()), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(t), static_cast<const uint16_t*>(x), scale, static_cast<uint16_t*>(y),
      pix_per_image * N * C, pix_per_image, N, C, planes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_f32_to_u8(const float* in, long long npix, int cpad, const float* add3, float mul, uint8_t* out, float* out_f,
                  void* stream) {
  MTB_REQUIRE(in && out && cpad >= 3, "mtb_f32_to_u8: bad arguments");
  const float a0 = add3 ? add3[0] : 0.f, a1 = add3 ? add3[1] : 0.f, a2 = add3 ? add3[2] : 0.f;
  f32_to_u8_kernel<<<grid_for(npix, 256, sm_count()), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, npix, cpad, a0,
                                                                                                 a1, a2, mul, out,
                                                                                                 out_f);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
