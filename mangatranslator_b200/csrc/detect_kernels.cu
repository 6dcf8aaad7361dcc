// Detection glue on the device: SPPF max-pool, nearest 2x upsample (both on channel slices of NHWC plane tensors),
// YOLO head decode (DFL + sigmoid + confidence filter), NMS, and the reference's own post-NMS box logic
// (core/image/detection.py:219-295 `_deduplicate_primary_boxes`, `_remove_contained_boxes`) so that the kept INDICES
// are produced on the GPU without a host round trip.
#include <atomic>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

int sm_count2() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}
inline int grid_for2(long long items, int block) {
  long long g = (items + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count2()) * 16;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

// k x k max-pool, stride 1, pad k/2 (padding never wins), on channels [ci, ci+c) -> [co, co+c); 8 channels per thread
__global__ void maxpool_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int ct_in,
                               int ci, int ct_out, int co, int c, int k, int planes, long long ps_in, long long ps_out) {
  const int vec = c / 8;
  const long long total = static_cast<long long>(N) * H * W * vec;
  const int r = k / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long p = i / vec;
    const int px = static_cast<int>(p % W);
    p /= W;
    const int py = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    float best[8];
    uint16_t bh[8], bl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bh[j] = 0xFF80;  // -inf in bf16
      bl[j] = 0;
    }
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = py + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = px + dx;
        if (xx < 0 || xx >= W) continue;
        const long long off = ((static_cast<long long>(n) * H + yy) * W + xx) * ct_in + ci + v * 8;
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
    const long long oo = ((static_cast<long long>(n) * H + py) * W + px) * ct_out + co + v * 8;
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

// Depthwise k x k convolution (groups = channels), stride 1, pad k/2, + bias (+ SiLU) on channels [ci, ci+c) -> [co, co+c):
// the DWConv of the YOLO11 class branch and the 3x3 / 7x7 positional convolution of the PSA / area-attention blocks
// (ultralytics nn/modules/conv.py DWConv, block.py Attention.pe / AAttn.pe).  No contraction over channels, so no tensor
// cores: 8 channels per thread (16-byte loads of both planes), fp32 accumulation in tap order (ky, kx) like a direct conv.
// w: fp32 [k*k][c] (tap-major, so a thread's 8 weights are contiguous), bias fp32 [c].
__global__ void dwconv_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int ct_in, int ci,
                              int ct_out, int co, int c, int k, const float* __restrict__ w, const float* __restrict__ bias,
                              int act, int planes, long long ps_in, long long ps_out) {
  const int vec = c / 8;
  const long long total = static_cast<long long>(N) * H * W * vec;
  const int r = k / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long p = i / vec;
    const int px = static_cast<int>(p % W);
    p /= W;
    const int py = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = py + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = px + dx;
        if (xx < 0 || xx >= W) continue;
        const long long off = ((static_cast<long long>(n) * H + yy) * W + xx) * ct_in + ci + v * 8;
        const uint4 h4 = *reinterpret_cast<const uint4*>(x + off);
        uint4 l4 = make_uint4(0, 0, 0, 0);
        if (planes == 2) l4 = *reinterpret_cast<const uint4*>(x + ps_in + off);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
        const float* wt = w + static_cast<long long>((dy + r) * k + (dx + r)) * c + v * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wt), w1 = *reinterpret_cast<const float4*>(wt + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint16_t hh = (j & 1) ? (hw[j >> 1] >> 16) : (hw[j >> 1] & 0xFFFF);
          const uint16_t ll = (j & 1) ? (lw[j >> 1] >> 16) : (lw[j >> 1] & 0xFFFF);
          acc[j] = fmaf(bf16_to_f(hh) + bf16_to_f(ll), wv[j], acc[j]);
        }
      }
    }
    uint16_t oh[8], ol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = acc[j] + bias[v * 8 + j];
      if (act) o = o / (1.0f + expf(-o));
      split_bf16(o, oh[j], ol[j]);
    }
    const long long oo = ((static_cast<long long>(n) * H + py) * W + px) * ct_out + co + v * 8;
    uint4 qh, ql;
    qh.x = oh[0] | (static_cast<uint32_t>(oh[1]) << 16);
    qh.y = oh[2] | (static_cast<uint32_t>(oh[3]) << 16);
    qh.z = oh[4] | (static_cast<uint32_t>(oh[5]) << 16);
    qh.w = oh[6] | (static_cast<uint32_t>(oh[7]) << 16);
    ql.x = ol[0] | (static_cast<uint32_t>(ol[1]) << 16);
    ql.y = ol[2] | (static_cast<uint32_t>(ol[3]) << 16);
    ql.z = ol[4] | (static_cast<uint32_t>(ol[5]) << 16);
    ql.w = ol[6] | (static_cast<uint32_t>(ol[7]) << 16);
    *reinterpret_cast<uint4*>(y + oo) = qh;
    if (planes == 2) *reinterpret_cast<uint4*>(y + ps_out + oo) = ql;
  }
}

// nearest 2x upsample of a channel slice: out[n, 2y+a, 2x+b, co+..] = in[n, y, x, ci+..]
__global__ void upsample2x_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int ct_in,
                                  int ci, int ct_out, int co, int c, int planes, long long ps_in, long long ps_out) {
  const int vec = c / 8;
  const long long total = static_cast<long long>(N) * (2 * H) * (2 * W) * vec * planes;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long p = i;
    const int v = static_cast<int>(p % vec);
    p /= vec;
    const int ox = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int oy = static_cast<int>(p % (2 * H));
    p /= 2 * H;
    const int n = static_cast<int>(p % N);
    const int pl = static_cast<int>(p / N);
    const long long src = ((static_cast<long long>(n) * H + (oy >> 1)) * W + (ox >> 1)) * ct_in + ci + v * 8 + pl * ps_in;
    const long long dst = ((static_cast<long long>(n) * 2 * H + oy) * 2 * W + ox) * ct_out + co + v * 8 + pl * ps_out;
    *reinterpret_cast<uint4*>(y + dst) = *reinterpret_cast<const uint4*>(x + src);
  }
}

// ---- YOLO head decode -----------------------------------------------------------------------------------------
struct Level {
  const float* box;  // [N][H][W][64]
  const float* cls;  // [N][H][W][ncp]
  int H, W, stride, anchor0;
};
struct DecodeParams {
  Level lv[3];
  int n_levels, N, nc, ncp, total_anchors;
  float conf;
  int max_cand;
};

// one thread per anchor: DFL (softmax over 16 bins, expectation), dist2bbox, sigmoid class scores, confidence filter.
// candidate record: [x1,y1,x2,y2 (letterboxed px), score, cls] + anchor index
__global__ void decode_kernel(DecodeParams P, float* __restrict__ cand, int* __restrict__ cand_anchor,
                              int* __restrict__ count) {
  const long long total = static_cast<long long>(P.N) * P.total_anchors;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / P.total_anchors);
    const int a = static_cast<int>(i - static_cast<long long>(n) * P.total_anchors);
    int li = 0;
    while (li + 1 < P.n_levels && a >= P.lv[li + 1].anchor0) ++li;
    const Level& L = P.lv[li];
    const int la = a - L.anchor0;
    const int ay = la / L.W, ax = la - ay * L.W;
    const long long pix = (static_cast<long long>(n) * L.H + ay) * L.W + ax;
    // class scores first (cheap reject)
    const float* cp = L.cls + pix * P.ncp;
    float best = -1.0f;
    int bc = 0;
    for (int c = 0; c < P.nc; ++c) {
      const float s = 1.0f / (1.0f + expf(-cp[c]));
      if (s > best) {
        best = s;
        bc = c;
      }
    }
    if (!(best > P.conf)) continue;
    const float* bp = L.box + pix * 64;
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
      float mx = bp[side * 16];
      for (int k = 1; k < 16; ++k) mx = fmaxf(mx, bp[side * 16 + k]);
      float e[16], sum = 0.f;
      for (int k = 0; k < 16; ++k) {
        e[k] = expf(bp[side * 16 + k] - mx);
        sum += e[k];
      }
      float acc = 0.f;
      for (int k = 0; k < 16; ++k) acc += (e[k] / sum) * static_cast<float>(k);
      d[side] = acc;
    }
    const float cx = static_cast<float>(ax) + 0.5f, cy = static_cast<float>(ay) + 0.5f;
    const float x1 = cx - d[0], y1 = cy - d[1], x2 = cx + d[2], y2 = cy + d[3];
    const float st = static_cast<float>(L.stride);
    // xywh in input pixels (as the head emits), then xyxy as non_max_suppression converts it back
    const float bx = (x1 + x2) / 2.0f * st, by = (y1 + y2) / 2.0f * st, bw = (x2 - x1) * st, bh = (y2 - y1) * st;
    const int slot = atomicAdd(&count[n], 1);
    if (slot < P.max_cand) {
      float* o = cand + (static_cast<long long>(n) * P.max_cand + slot) * 6;
      o[0] = bx - bw / 2.0f;
      o[1] = by - bh / 2.0f;
      o[2] = bx + bw / 2.0f;
      o[3] = by + bh / 2.0f;
      o[4] = best;
      o[5] = static_cast<float>(bc);
      cand_anchor[static_cast<long long>(n) * P.max_cand + slot] = a;
    }
  }
}

// ---- NMS + reference post-processing, one block per image ---------------------------------------------------------
struct NmsParams {
  int N, max_cand, max_det;
  float iou_thr;
  float max_wh;
  // scale_boxes (letterboxed -> original) parameters
  float gain;
  int pad_x, pad_y;
  int img_w, img_h;
  // reference post-NMS logic
  double dedup_iou;     // IOU_DUPLICATE_THRESHOLD = 0.7 (detection.py:19)
  double contain_ioa;   // 0.9 (detection.py:260)
  int apply_dedup;
};

__device__ __forceinline__ bool cand_before(float sa, int aa, float sb, int ab) {
  // descending score, ties by ascending anchor index (== stable argsort of the head's anchor order)
  return sa > sb || (sa == sb && aa < ab);
}

__device__ double box_inter_d(const double* a, const double* b) {
  const double x0 = a[0] > b[0] ? a[0] : b[0], y0 = a[1] > b[1] ? a[1] : b[1];
  const double x1 = a[2] < b[2] ? a[2] : b[2], y1 = a[3] < b[3] ? a[3] : b[3];
  const double w = __dadd_rn(x1, -x0), h = __dadd_rn(y1, -y0);
  return __dmul_rn(w > 0.0 ? w : 0.0, h > 0.0 ? h : 0.0);
}
__device__ double box_area_d(const double* a) {
  const double w = __dadd_rn(a[2], -a[0]), h = __dadd_rn(a[3], -a[1]);
  return __dmul_rn(w > 0.0 ? w : 0.0, h > 0.0 ? h : 0.0);
}


// rank of every candidate in (score desc, anchor asc) order: order[n][rank] = candidate slot.  Grid = (chunks, N).
__global__ void rank_kernel(const float* __restrict__ cand, const int* __restrict__ cand_anchor,
                            const int* __restrict__ count, int max_cand, int* __restrict__ order_ws) {
  const int n = blockIdx.y;
  const int m = count[n] < max_cand ? count[n] : max_cand;
  const float* c = cand + static_cast<long long>(n) * max_cand * 6;
  const int* ca = cand_anchor + static_cast<long long>(n) * max_cand;
  int* order = order_ws + static_cast<long long>(n) * max_cand;
  __shared__ float s_sc[256];
  __shared__ int s_an[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float si = i < m ? c[i * 6 + 4] : 0.f;
  const int ai = i < m ? ca[i] : 0;
  int r = 0;
  for (int j0 = 0; j0 < m; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    s_sc[threadIdx.x] = j < m ? c[j * 6 + 4] : -1.f;
    s_an[threadIdx.x] = j < m ? ca[j] : 0x7fffffff;
    __syncthreads();
    const int lim = min(256, m - j0);
    if (i < m)
      for (int t = 0; t < lim; ++t)
        if (cand_before(s_sc[t], s_an[t], si, ai)) ++r;
  }
  if (i < m) order[r] = i;
}

// out_det: [N][max_det][8] = x1,y1,x2,y2 (original px), score, cls, anchor, kept_by_reference_logic(0/1)
// out_count: [N][2] = (n after NMS, n after dedup+containment)
__global__ void nms_kernel(NmsParams P, const float* __restrict__ cand, const int* __restrict__ cand_anchor,
                           const int* __restrict__ count, int* __restrict__ order_ws, unsigned char* __restrict__ dead_ws,
                           float* __restrict__ out_det, int* __restrict__ out_count, int* __restrict__ final_idx) {
  const int n = blockIdx.x;
  const int m = count[n] < P.max_cand ? count[n] : P.max_cand;
  const float* c = cand + static_cast<long long>(n) * P.max_cand * 6;
  const int* ca = cand_anchor + static_cast<long long>(n) * P.max_cand;
  int* order = order_ws + static_cast<long long>(n) * P.max_cand;
  unsigned char* dead = dead_ws + static_cast<long long>(n) * P.max_cand;
  __shared__ int s_keep[512];
  __shared__ int s_nkeep;
  // `order` was filled by rank_kernel (one thread per candidate, many blocks)
  for (int i = threadIdx.x; i < m; i += blockDim.x) dead[i] = 0;
  if (threadIdx.x == 0) s_nkeep = 0;
  __syncthreads();
  // greedy NMS in sorted order, class offset trick (boxes + cls * max_wh), fp32 like torchvision
  for (int oi = 0; oi < m; ++oi) {
    const int i = order[oi];
    __syncthreads();
    if (dead[oi]) continue;
    if (s_nkeep >= P.max_det) break;
    if (threadIdx.x == 0) s_keep[s_nkeep] = i;
    const float off_i = c[i * 6 + 5] * P.max_wh;
    const float ix1 = c[i * 6 + 0] + off_i, iy1 = c[i * 6 + 1] + off_i, ix2 = c[i * 6 + 2] + off_i, iy2 = c[i * 6 + 3] + off_i;
    const float iarea = (ix2 - ix1) * (iy2 - iy1);
    for (int oj = oi + 1 + threadIdx.x; oj < m; oj += blockDim.x) {
      if (dead[oj]) continue;
      const int j = order[oj];
      const float off_j = c[j * 6 + 5] * P.max_wh;
      const float jx1 = c[j * 6 + 0] + off_j, jy1 = c[j * 6 + 1] + off_j, jx2 = c[j * 6 + 2] + off_j, jy2 = c[j * 6 + 3] + off_j;
      const float xx1 = fmaxf(ix1, jx1), yy1 = fmaxf(iy1, jy1), xx2 = fminf(ix2, jx2), yy2 = fminf(iy2, jy2);
      const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
      const float inter = w * h;
      const float jarea = (jx2 - jx1) * (jy2 - jy1);
      const float ovr = inter / (iarea + jarea - inter);
      if (ovr > P.iou_thr) dead[oj] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) ++s_nkeep;
    __syncthreads();
  }
  __syncthreads();
  const int nk = s_nkeep;
  float* od = out_det + static_cast<long long>(n) * P.max_det * 8;
  // scale_boxes: remove letterbox padding, divide by gain, clip (fp32 like the predictor)
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const int i = s_keep[k];
    float x1 = (c[i * 6 + 0] - static_cast<float>(P.pad_x)) / P.gain;
    float y1 = (c[i * 6 + 1] - static_cast<float>(P.pad_y)) / P.gain;
    float x2 = (c[i * 6 + 2] - static_cast<float>(P.pad_x)) / P.gain;
    float y2 = (c[i * 6 + 3] - static_cast<float>(P.pad_y)) / P.gain;
    x1 = fminf(fmaxf(x1, 0.f), static_cast<float>(P.img_w));
    x2 = fminf(fmaxf(x2, 0.f), static_cast<float>(P.img_w));
    y1 = fminf(fmaxf(y1, 0.f), static_cast<float>(P.img_h));
    y2 = fminf(fmaxf(y2, 0.f), static_cast<float>(P.img_h));
    float* o = od + k * 8;
    o[0] = x1;
    o[1] = y1;
    o[2] = x2;
    o[3] = y2;
    o[4] = c[i * 6 + 4];
    o[5] = c[i * 6 + 5];
    o[6] = static_cast<float>(ca[i]);
    o[7] = 1.0f;
  }
  __syncthreads();
  // reference logic, sequential and order dependent, in double like the Python floats it runs on
  if (threadIdx.x == 0) {
    int nfinal = nk;
    int* fi = final_idx + static_cast<long long>(n) * P.max_det;
    for (int k = 0; k < nk; ++k) fi[k] = k;
    if (P.apply_dedup && nk > 1) {
      // _deduplicate_primary_boxes: stable sort by confidence (descending), greedy IoU > thr
      int ord[512];
      for (int k = 0; k < nk; ++k) ord[k] = k;
      for (int a = 1; a < nk; ++a) {  // insertion sort keeps equal scores in index order (Python sorted(reverse=True))
        const int v = ord[a];
        int b = a - 1;
        while (b >= 0 && od[ord[b] * 8 + 4] < od[v * 8 + 4]) {
          ord[b + 1] = ord[b];
          --b;
        }
        ord[b + 1] = v;
      }
      int keep[512];
      int nkeep = 0;
      for (int a = 0; a < nk; ++a) {
        const int i = ord[a];
        const double bi[4] = {od[i * 8 + 0], od[i * 8 + 1], od[i * 8 + 2], od[i * 8 + 3]};
        bool dup = false;
        for (int q = 0; q < nkeep; ++q) {
          const int j = keep[q];
          const double bj[4] = {od[j * 8 + 0], od[j * 8 + 1], od[j * 8 + 2], od[j * 8 + 3]};
          const double inter = box_inter_d(bi, bj);
          const double uni = __dadd_rn(__dadd_rn(box_area_d(bi), box_area_d(bj)), -inter);
          const double iou = uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
          if (iou > P.dedup_iou) {
            dup = true;
            break;
          }
        }
        if (!dup) keep[nkeep++] = i;
      }
      // _remove_contained_boxes on boxes[keep] (in `keep` order): drop i if IoA(i in j) > thr for a still-kept j
      bool alive[512];
      for (int a = 0; a < nkeep; ++a) alive[a] = true;
      for (int a = 0; a < nkeep; ++a) {
        if (!alive[a]) continue;
        const int i = keep[a];
        const double bi[4] = {od[i * 8 + 0], od[i * 8 + 1], od[i * 8 + 2], od[i * 8 + 3]};
        const double ai = box_area_d(bi);
        for (int b = 0; b < nkeep; ++b) {
          if (a == b || !alive[b]) continue;
          const int j = keep[b];
          const double bj[4] = {od[j * 8 + 0], od[j * 8 + 1], od[j * 8 + 2], od[j * 8 + 3]};
          const double ioa = ai > 0.0 ? __ddiv_rn(box_inter_d(bi, bj), ai) : 0.0;
          if (ioa > P.contain_ioa) {
            alive[a] = false;
            break;
          }
        }
      }
      nfinal = 0;
      for (int k = 0; k < nk; ++k) od[k * 8 + 7] = 0.0f;
      for (int a = 0; a < nkeep; ++a)
        if (alive[a]) {
          fi[nfinal++] = keep[a];
          od[keep[a] * 8 + 7] = 1.0f;
        }
    }
    out_count[n * 2 + 0] = nk;
    out_count[n * 2 + 1] = nfinal;
  }
}


// ---- retina masks (ultralytics process_mask_native): coeff . proto -> strip letterbox pad -> bilinear to the page ->
//      zero outside the box -> > 0.  One thread per page pixel per detection; the 32-term dot product is evaluated at the
//      four neighbouring prototype pixels.
struct MaskLevel {
  const float* mc;  // [H][W][32]
  int H, W, anchor0;
};
__global__ void yolo_masks_kernel(const float* __restrict__ proto, int mh, int mw, int nm, MaskLevel l0, MaskLevel l1,
                                  MaskLevel l2, const float* __restrict__ det /* [n][8] */, const int* __restrict__ rows,
                                  int n, int top, int left, int ch, int cw, int H, int W, uint8_t* __restrict__ out) {
  const long long per = static_cast<long long>(H) * W;
  const long long total = per * n;
  const float sy = static_cast<float>(ch) / static_cast<float>(H), sx = static_cast<float>(cw) / static_cast<float>(W);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / per);
    const long long r = i - static_cast<long long>(k) * per;
    const int y = static_cast<int>(r / W), x = static_cast<int>(r - static_cast<long long>(y) * W);
    const float* d = det + static_cast<long long>(rows ? rows[k] : k) * 8;
    uint8_t o = 0;
    const float fxp = static_cast<float>(x), fyp = static_cast<float>(y);
    if (fxp >= d[0] && fxp < d[2] && fyp >= d[1] && fyp < d[3]) {
      const int a = static_cast<int>(d[6]);
      const MaskLevel& L = a >= l2.anchor0 ? l2 : (a >= l1.anchor0 ? l1 : l0);
      const float* c = L.mc + static_cast<long long>(a - L.anchor0) * nm;
      float fy = sy * (fyp + 0.5f) - 0.5f, fx = sx * (fxp + 0.5f) - 0.5f;
      if (fy < 0.f) fy = 0.f;
      if (fx < 0.f) fx = 0.f;
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = y0 + (y0 < ch - 1 ? 1 : 0), x1 = x0 + (x0 < cw - 1 ? 1 : 0);
      const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
      const float* p00 = proto + (static_cast<long long>(y0 + top) * mw + x0 + left) * nm;
      const float* p01 = proto + (static_cast<long long>(y0 + top) * mw + x1 + left) * nm;
      const float* p10 = proto + (static_cast<long long>(y1 + top) * mw + x0 + left) * nm;
      const float* p11 = proto + (static_cast<long long>(y1 + top) * mw + x1 + left) * nm;
      float m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f;
      for (int q = 0; q < nm; ++q) {
        const float cq = c[q];
        m00 += cq * p00[q];
        m01 += cq * p01[q];
        m10 += cq * p10[q];
        m11 += cq * p11[q];
      }
      const float v = (1.f - ly) * ((1.f - lx) * m00 + lx * m01) + ly * ((1.f - lx) * m10 + lx * m11);
      o = v > 0.0f ? 1 : 0;
    }
    out[i] = o;
  }
}

}  // namespace

extern "C" {

int mtb_maxpool(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int k,
                int planes, void* stream) {
  MTB_REQUIRE(x && y && c % 8 == 0 && ci % 8 == 0 && co % 8 == 0 && (k & 1), "mtb_maxpool: bad arguments");
  const long long total = static_cast<long long>(N) * H * W * (c / 8);
  maxpool_kernel<<<grid_for2(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), N, H, W, ct_in, ci, ct_out, co, c, k, planes,
      static_cast<long long>(N) * H * W * ct_in, static_cast<long long>(N) * H * W * ct_out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_dwconv(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int k,
               const float* w, const float* bias, int act, int planes, void* stream) {
  MTB_REQUIRE(x && y && w && bias && x != y, "mtb_dwconv: null or aliased argument");
  MTB_REQUIRE(c > 0 && c % 8 == 0 && ci % 8 == 0 && co % 8 == 0 && ct_in % 8 == 0 && ct_out % 8 == 0 && (k & 1) && k <= 15,
              "mtb_dwconv: channels must come in multiples of 8 and the kernel size must be odd (c %d ci %d co %d k %d)", c,
              ci, co, k);
  MTB_REQUIRE(planes == 1 || planes == 2, "mtb_dwconv: planes must be 1 or 2");
  const long long total = static_cast<long long>(N) * H * W * (c / 8);
  dwconv_kernel<<<grid_for2(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), N, H, W, ct_in, ci, ct_out, co, c, k, w, bias, act, planes,
      static_cast<long long>(N) * H * W * ct_in, static_cast<long long>(N) * H * W * ct_out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_upsample2x(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int planes,
                   void* stream) {
  MTB_REQUIRE(x && y && c % 8 == 0 && ci % 8 == 0 && co % 8 == 0, "mtb_upsample2x: bad arguments");
  const long long total = static_cast<long long>(N) * 4 * H * W * (c / 8) * planes;
  upsample2x_kernel<<<grid_for2(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), N, H, W, ct_in, ci, ct_out, co, c, planes,
      static_cast<long long>(N) * H * W * ct_in, static_cast<long long>(N) * 4 * H * W * ct_out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_yolo_decode(const mtb_yolo_level* levels, int n_levels, int N, int nc, int ncp, float conf, int max_cand,
                    float* cand, int* cand_anchor, int* count, void* stream) {
  MTB_REQUIRE(levels && n_levels >= 1 && n_levels <= 3 && cand && cand_anchor && count, "mtb_yolo_decode: bad arguments");
  DecodeParams P;
  int a0 = 0;
  for (int i = 0; i < n_levels; ++i) {
    P.lv[i].box = levels[i].box;
    P.lv[i].cls = levels[i].cls;
    P.lv[i].H = levels[i].H;
    P.lv[i].W = levels[i].W;
    P.lv[i].stride = levels[i].stride;
    P.lv[i].anchor0 = a0;
    a0 += levels[i].H * levels[i].W;
  }
  P.n_levels = n_levels;
  P.N = N;
  P.nc = nc;
  P.ncp = ncp;
  P.total_anchors = a0;
  P.conf = conf;
  P.max_cand = max_cand;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTB_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int) * N, st));
  decode_kernel<<<grid_for2(static_cast<long long>(N) * a0, 128), 128, 0, st>>>(P, cand, cand_anchor, count);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_nms(const mtb_nms_params* p, const float* cand, const int* cand_anchor, const int* count, int* order_ws,
            unsigned char* dead_ws, float* out_det, int* out_count, int* final_idx, void* stream) {
  MTB_REQUIRE(p && cand && cand_anchor && count && order_ws && dead_ws && out_det && out_count && final_idx,
              "mtb_nms: null argument");
  MTB_REQUIRE(p->max_det <= 512, "mtb_nms: max_det must be <= 512");
  NmsParams P;
  P.N = p->N;
  P.max_cand = p->max_cand;
  P.max_det = p->max_det;
  P.iou_thr = p->iou_thr;
  P.max_wh = p->max_wh;
  P.gain = p->gain;
  P.pad_x = p->pad_x;
  P.pad_y = p->pad_y;
  P.img_w = p->img_w;
  P.img_h = p->img_h;
  P.dedup_iou = p->dedup_iou;
  P.contain_ioa = p->contain_ioa;
  P.apply_dedup = p->apply_dedup;
  {
    dim3 rgrid(static_cast<unsigned>((P.max_cand + 255) / 256), static_cast<unsigned>(P.N));
    rank_kernel<<<rgrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(cand, cand_anchor, count, P.max_cand, order_ws);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
  }
  nms_kernel<<<P.N, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, cand, cand_anchor, count, order_ws, dead_ws, out_det,
                                                                out_count, final_idx);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}


int mtb_yolo_masks(const float* proto, int mh, int mw, int nm, const float* const* mc_levels, const int* level_hw /* 3x2 */,
                   const float* det, const int* rows, int n, int top, int left, int ch, int cw, int H, int W, uint8_t* out,
                   void* stream) {
  MTB_REQUIRE(proto && mc_levels && level_hw && det && out, "mtb_yolo_masks: null argument");
  if (n <= 0) return 0;
  MaskLevel L[3];
  int a0 = 0;
  for (int i = 0; i < 3; ++i) {
    L[i].mc = mc_levels[i];
    L[i].H = level_hw[2 * i];
    L[i].W = level_hw[2 * i + 1];
    L[i].anchor0 = a0;
    a0 += L[i].H * L[i].W;
  }
  yolo_masks_kernel<<<grid_for2(static_cast<long long>(n) * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      proto, mh, mw, nm, L[0], L[1], L[2], det, rows, n, top, left, ch, cw, H, W, out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
