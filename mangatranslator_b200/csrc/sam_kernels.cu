// SAM 2.1 glue kernels (reference: transformers Sam2Model behind core/image/detection.py:475-511):
//   layer norm (+GELU), 2x2 max-pool, plane add with broadcast, multi-head attention with the Hiera window
//   partition / zero-padding / query-pooling folded into its addressing, box-prompt encoding, the hyper-network
//   mask product, the stability-based mask selection and the fused "bilinear 256^2 -> HxW, > 0, clip to the box,
//   write uint8 0/255" mask writer (core/image/detection.py:504-511,1732-1750).
// All softmax / normalisation math is fp32; activations travel as bf16 hi/lo planes like everywhere else.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace mtb {
long long attention_tc_workspace_bytes(int B, int heads, int hd, int nk);
int launch_attention_tc(const mtb_attn_desc* d, cudaStream_t st);
}  // namespace mtb

namespace {

int sm_count4() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}
inline int grid_for4(long long items, int block) {
  long long g = (items + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count4()) * 16;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

__device__ __forceinline__ float ld_planes(const uint16_t* p, long long idx, long long ps, int planes) {
  float v = bf16_to_f(p[idx]);
  if (planes == 2) v += bf16_to_f(p[ps + idx]);
  return v;
}
__device__ __forceinline__ void st_planes(uint16_t* p, long long idx, long long ps, int planes, float v) {
  uint16_t h, l;
  split_bf16(v, h, l);
  p[idx] = h;
  if (planes == 2) p[ps + idx] = l;
}

// ---- layer norm over the channel dimension of each row (pixel / token), optional GELU --------------------
// one warp per row; two-pass mean / variance in fp32 like torch.nn.LayerNorm
__global__ void layernorm_kernel(const uint16_t* __restrict__ x, long long rows, int C, int ct_in, int ci, long long ps_in,
                                 int planes_in, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 uint16_t* __restrict__ y, int ct_out, int co, long long ps_out, int planes_out, int gelu) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * warps_per_block) {
    const long long base = r * ct_in + ci;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += ld_planes(x, base + c, ps_in, planes_in);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / static_cast<float>(C);
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = ld_planes(x, base + c, ps_in, planes_in) - mean;
      v += d * d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / static_cast<float>(C) + eps);
    const long long ob = r * ct_out + co;
    for (int c = lane; c < C; c += 32) {
      float o = (ld_planes(x, base + c, ps_in, planes_in) - mean) * rstd * gamma[c] + beta[c];
      if (gelu) o = 0.5f * o * (1.0f + erff(o * 0.70710678118654752440f));
      st_planes(y, ob + c, ps_out, planes_out, o);
    }
  }
}

// ---- 2x2 / stride 2 max-pool of an NHWC plane tensor --------------------------------------------------------
__global__ void maxpool2x2_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int N, int H, int W, int C,
                                  long long ps_in, long long ps_out, int planes) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long p = i / C;
    const int ox = static_cast<int>(p % Wo);
    p /= Wo;
    const int oy = static_cast<int>(p % Ho);
    const int n = static_cast<int>(p / Ho);
    float best = -INFINITY;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const long long idx = ((static_cast<long long>(n) * H + 2 * oy + a) * W + 2 * ox + b) * C + c;
        const float v = ld_planes(x, idx, ps_in, planes);
        best = fmaxf(best, v);
      }
    st_planes(y, i, ps_out, planes, best);
  }
}

// ---- out = a + b (b broadcast over the batch when b_rows < rows) -------------------------------------------------
__global__ void add_planes_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, uint16_t* __restrict__ o,
                                  long long rows, int C, long long b_rows, long long ps_a, long long ps_b, long long ps_o,
                                  int planes) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i - r * C);
    const long long bi = (r % b_rows) * C + c;
    st_planes(o, i, ps_o, planes, ld_planes(a, i, ps_a, planes) + ld_planes(b, bi, ps_b, planes));
  }
}

// ---- attention ----------------------------------------------------------------------------------------------------
struct AttnParams {
  int B, heads, hd, nq, nk;
  float scale;
  const uint16_t *q, *k, *v;
  uint16_t* out;
  int q_ct, q_off, k_ct, k_off, v_ct, v_off, o_ct, o_off;
  long long q_ps, k_ps, v_ps, o_ps;
  int planes;
  int mode;  // 0: rows b*n + t ; 1: Hiera windows on a grid
  int grid_h, grid_w, ws, pool, nwx;
  const float *pad_q, *pad_k, *pad_v;  // value of a padded token (= the qkv bias) [heads*hd] each
};

// memory row of token t of batch b, or -1 for a padded (out-of-grid) window position
__device__ __forceinline__ long long tok_row(const AttnParams& P, int b, int t, int n) {
  if (P.mode == 0) return static_cast<long long>(b) * n + t;
  const int by = b / P.nwx, bx = b - by * P.nwx;
  const int ty = t / P.ws, tx = t - ty * P.ws;
  const int y = by * P.ws + ty, x = bx * P.ws + tx;
  if (y >= P.grid_h || x >= P.grid_w) return -1;
  return static_cast<long long>(y) * P.grid_w + x;
}

constexpr int kAttnThreads = 256;
constexpr int kQT = 64;   // queries per block (8 per warp)
constexpr int kKT = 64;   // keys per shared-memory tile
constexpr int kMaxHD = 128;

// 8 consecutive channels of one token as fp32 (16-byte loads of the hi and lo planes); padded tokens read `pad`
__device__ __forceinline__ void load_tok8(const AttnParams& P, const uint16_t* base, int ct, int off, long long ps,
                                          const float* pad, long long row, int ch, float (&v)[8]) {
  if (row < 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = pad ? pad[ch + j] : 0.f;
    return;
  }
  const long long idx = row * ct + off + ch;
  const uint4 h = *reinterpret_cast<const uint4*>(base + idx);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(hw[j] << 16);
    v[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u);
  }
  if (P.planes == 2) {
    const uint4 l = *reinterpret_cast<const uint4*>(base + ps + idx);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] += __uint_as_float(lw[j] << 16);
      v[2 * j + 1] += __uint_as_float(lw[j] & 0xFFFF0000u);
    }
  }
}

// Flash-style attention on CUDA cores in fp32 (exact softmax, online rescaling).  Block = 64 queries x one (batch, head);
// each warp owns 8 queries processed as two register-blocked groups of 4 so every K/V shared-memory read feeds 4 FMAs.
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(AttnParams P) {
  extern __shared__ float sm[];
  const int hd = P.hd, hdp = hd + 1;
  float* sK = sm;                     // [kKT][hd+1]
  float* sV = sK + kKT * hdp;         // [kKT][hd]
  float* sQ = sV + kKT * hd;          // [kQT][hd]
  const int bh = blockIdx.x;
  const int b = bh / P.heads, h = bh - b * P.heads;
  const int q0 = blockIdx.y * kQT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq_tile = min(kQT, P.nq - q0);
  const int wso = P.pool ? P.ws / 2 : P.ws;
  const int hv = hd >> 3;             // 8-channel vectors per token

  // stage the block's queries, applying the 2x2 max-pool inside the window when requested
  for (int i = threadIdx.x; i < kQT * hv; i += kAttnThreads) {
    const int qi = i / hv, dv = i - qi * hv;
    float v[8];
    if (qi >= nq_tile) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else {
      const int t = q0 + qi;
      if (P.mode == 1 && P.pool) {
        const int qy = t / wso, qx = t - qy * wso;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
        for (int a = 0; a < 2; ++a)
          for (int c = 0; c < 2; ++c) {
            float u[8];
            const int tt = (2 * qy + a) * P.ws + 2 * qx + c;
            load_tok8(P, P.q, P.q_ct, P.q_off, P.q_ps, P.pad_q, tok_row(P, b, tt, P.nk), h * hd + dv * 8, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], u[j]);
          }
      } else {
        load_tok8(P, P.q, P.q_ct, P.q_off, P.q_ps, P.pad_q, tok_row(P, b, t, P.nq), h * hd + dv * 8, v);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sQ[qi * hd + dv * 8 + j] = v[j] * P.scale;
  }
  float m[8], l[8], acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = -INFINITY;
    l[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  }
  for (int k0 = 0; k0 < P.nk; k0 += kKT) {
    const int nk_tile = min(kKT, P.nk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kKT * hv; i += kAttnThreads) {
      const int ki = i / hv, dv = i - ki * hv;
      float kv[8], vv[8];
      if (ki < nk_tile) {
        const long long row = tok_row(P, b, k0 + ki, P.nk);
        load_tok8(P, P.k, P.k_ct, P.k_off, P.k_ps, P.pad_k, row, h * hd + dv * 8, kv);
        load_tok8(P, P.v, P.v_ct, P.v_off, P.v_ps, P.pad_v, row, h * hd + dv * 8, vv);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) kv[j] = vv[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sK[ki * hdp + dv * 8 + j] = kv[j];
        sV[ki * hd + dv * 8 + j] = vv[j];
      }
    }
    __syncthreads();
    const float* k0r = sK + lane * hdp;
    const float* k1r = sK + (lane + 32) * hdp;
    const bool v0 = lane < nk_tile, v1 = lane + 32 < nk_tile;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (warp * 8 + g * 4 >= nq_tile) break;
      const float* q0p = sQ + (warp * 8 + g * 4) * hd;
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int d = 0; d < hd; ++d) {
        const float ka = k0r[d], kb = k1r[d];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float qv = q0p[j * hd + d];
          s0[j] += qv * ka;
          s1[j] += qv * kb;
        }
      }
      float p0[4], p1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = g * 4 + j;
        const float a0 = v0 ? s0[j] : -INFINITY, a1 = v1 ? s1[j] : -INFINITY;
        float tm = fmaxf(a0, a1);
#pragma unroll
        for (int o = 16; o; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
        const float nm = fmaxf(m[jj], tm);
        const float corr = expf(m[jj] - nm);
        p0[j] = v0 ? expf(a0 - nm) : 0.f;
        p1[j] = v1 ? expf(a1 - nm) : 0.f;
        float ps = p0[j] + p1[j];
#pragma unroll
        for (int o = 16; o; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
        l[jj] = l[jj] * corr + ps;
        m[jj] = nm;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[jj][e] *= corr;
      }
      for (int kk = 0; kk < nk_tile; ++kk) {
        const float* vr = sV + kk * hd;
        float ve[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) ve[e] = (lane + 32 * e < hd) ? vr[lane + 32 * e] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p = __shfl_sync(0xffffffffu, kk < 32 ? p0[j] : p1[j], kk & 31);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[g * 4 + j][e] += p * ve[e];
        }
      }
    }
  }
  // write: out[b, t, h*hd + d]; window mode writes at the un-partitioned grid position and drops the padding
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int qi = warp * 8 + j;
    if (qi >= nq_tile) break;
    const int t = q0 + qi;
    long long row;
    if (P.mode == 0) {
      row = static_cast<long long>(b) * P.nq + t;
    } else {
      const int by = b / P.nwx, bx = b - by * P.nwx;
      const int ty = t / wso, tx = t - ty * wso;
      const int y = by * wso + ty, x = bx * wso + tx;
      const int gh = P.pool ? P.grid_h / 2 : P.grid_h, gw = P.pool ? P.grid_w / 2 : P.grid_w;
      if (y >= gh || x >= gw) continue;
      row = static_cast<long long>(y) * gw + x;
    }
    const float inv = 1.0f / l[j];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = lane + 32 * e;
      if (d < hd) st_planes(P.out, row * P.o_ct + P.o_off + h * hd + d, P.o_ps, P.planes, acc[j][e] * inv);
    }
  }
}

// 8 consecutive channels of one token back into the bf16 planes (16-byte stores)
__device__ __forceinline__ void store_tok8(const AttnParams& P, long long idx, const float (&v)[8]) {
  uint16_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_bf16(v[j], hi[j], lo[j]);
  uint4 h4, l4;
  h4.x = hi[0] | (static_cast<uint32_t>(hi[1]) << 16);
  h4.y = hi[2] | (static_cast<uint32_t>(hi[3]) << 16);
  h4.z = hi[4] | (static_cast<uint32_t>(hi[5]) << 16);
  h4.w = hi[6] | (static_cast<uint32_t>(hi[7]) << 16);
  *reinterpret_cast<uint4*>(P.out + idx) = h4;
  if (P.planes == 2) {
    l4.x = lo[0] | (static_cast<uint32_t>(lo[1]) << 16);
    l4.y = lo[2] | (static_cast<uint32_t>(lo[3]) << 16);
    l4.z = lo[4] | (static_cast<uint32_t>(lo[5]) << 16);
    l4.w = lo[6] | (static_cast<uint32_t>(lo[7]) << 16);
    *reinterpret_cast<uint4*>(P.out + P.o_ps + idx) = l4;
  }
}

// Window / small-sequence attention with ONE THREAD PER QUERY (round 2).  attention_kernel above spreads a query's keys
// over the lanes of a warp: every K element is a separate shared-memory wavefront, probabilities travel by shuffle, and
// the block needs 74 KB of shared memory whatever the window size (three blocks per SM) — the Hiera window blocks ran at
// 0.13-0.24 ms per launch, one of them for 4 queries x 16 keys per block.  Here a thread keeps its query and its HD
// output accumulators in registers and walks the K / V tile with 16-byte BROADCAST loads (all lanes read the same key
// row: one wavefront feeds 4 x 32 FMAs), the online softmax needs no cross-lane traffic (rescaling once per 8 keys),
// and shared memory is sized by min(nk, 64) keys, so small windows run many blocks per SM.  Exact fp32 softmax.
// raw 16-byte loads of 8 channels (hi and lo plane) through the read-only path; a padded token (row < 0) reads row 0 and
// is patched by the caller, so that no branch sits between the loads and many stay in flight per thread
__device__ __forceinline__ void ldg_tok8(const AttnParams& P, const uint16_t* base, int ct, int off, long long ps, long long row,
                                         int ch, uint4& hi, uint4& lo) {
  const long long idx = (row < 0 ? 0 : row) * ct + off + ch;
  hi = __ldg(reinterpret_cast<const uint4*>(base + idx));
  lo = P.planes == 2 ? __ldg(reinterpret_cast<const uint4*>(base + ps + idx)) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void cvt_tok8(const uint4& hi, const uint4& lo, const float* pad, bool padded, int ch, float (&v)[8]) {
  const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
    v[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
  }
  if (padded) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = pad ? pad[ch + j] : 0.f;
  }
}

constexpr int kLpqQueries = 64;   // queries per block
constexpr int kLpqKeys = 64;      // keys per shared-memory tile
// SP = threads per query: with SP = 2 the two lanes of a pair own one half of the head's channels each (half of q and of
// the output accumulators: ~130 registers instead of 250, twice the warps per SM); the pair's partial dot products are
// added with one shuffle per key.
template <int HD, int SP>
__global__ void __launch_bounds__(kLpqQueries * SP, SP == 2 ? 3 : 1) attention_lpq_kernel(AttnParams P, int kt) {
  constexpr int kThreadsB = kLpqQueries * SP;
  constexpr int HH = HD / SP;                 // channels owned by a thread
  constexpr int hv = HD / 8, hvh = HH / 8;
  extern __shared__ __align__(16) float sm[];
  float* sK = sm;                 // [kt][HD]
  float* sV = sm + kt * HD;       // [kt][HD]
  const int bh = blockIdx.x;
  const int b = bh / P.heads, h = bh - b * P.heads;
  const int q0 = blockIdx.y * kLpqQueries;
  const int tid = threadIdx.x;
  const int t = q0 + tid / SP;
  const int c_base = (tid % SP) * HH;         // first channel of this thread inside the head
  const bool active = t < P.nq;
  const int wso = P.pool ? P.ws / 2 : P.ws;
  float q[HH], acc[HH];
#pragma unroll
  for (int d = 0; d < HH; ++d) {
    q[d] = 0.f;
    acc[d] = 0.f;
  }
  if (active) {
    if (P.mode == 1 && P.pool) {            // Hiera query pooling: 2x2 max inside the window
      const int qy = t / wso, qx = t - qy * wso;
#pragma unroll
      for (int d = 0; d < HH; ++d) q[d] = -INFINITY;
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          const long long row = tok_row(P, b, (2 * qy + a) * P.ws + 2 * qx + c, P.nk);
          uint4 rh[hvh], rl[hvh];
#pragma unroll
          for (int dv = 0; dv < hvh; ++dv)
            ldg_tok8(P, P.q, P.q_ct, P.q_off, P.q_ps, row, h * HD + c_base + dv * 8, rh[dv], rl[dv]);
#pragma unroll
          for (int dv = 0; dv < hvh; ++dv) {
            float u[8];
            cvt_tok8(rh[dv], rl[dv], P.pad_q, row < 0, h * HD + c_base + dv * 8, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) q[dv * 8 + j] = fmaxf(q[dv * 8 + j], u[j]);
          }
        }
    } else {
      const long long row = tok_row(P, b, t, P.nq);
      uint4 rh[hvh], rl[hvh];
#pragma unroll
      for (int dv = 0; dv < hvh; ++dv)
        ldg_tok8(P, P.q, P.q_ct, P.q_off, P.q_ps, row, h * HD + c_base + dv * 8, rh[dv], rl[dv]);
#pragma unroll
      for (int dv = 0; dv < hvh; ++dv) {
        float u[8];
        cvt_tok8(rh[dv], rl[dv], P.pad_q, row < 0, h * HD + c_base + dv * 8, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) q[dv * 8 + j] = u[j];
      }
    }
#pragma unroll
    for (int d = 0; d < HH; ++d) q[d] *= P.scale;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < P.nk; k0 += kt) {
    const int nk_tile = min(kt, P.nk - k0);
    __syncthreads();
    constexpr int U = (HH > 72 || SP == 2) ? 2 : 4;   // 8-16 x 16-byte loads in flight per thread (q and acc stay live)
    const int items = nk_tile * hv;
    for (int i0 = tid; i0 < items; i0 += U * kThreadsB) {
      uint4 kh[U], kl[U], vh[U], vl[U];
      long long rows[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = min(i0 + u * kThreadsB, items - 1);
        const int ki = i / hv, dv = i - ki * hv;
        rows[u] = tok_row(P, b, k0 + ki, P.nk);
        ldg_tok8(P, P.k, P.k_ct, P.k_off, P.k_ps, rows[u], h * HD + dv * 8, kh[u], kl[u]);
        ldg_tok8(P, P.v, P.v_ct, P.v_off, P.v_ps, rows[u], h * HD + dv * 8, vh[u], vl[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * kThreadsB;
        if (i >= items) break;
        const int ki = i / hv, dv = i - ki * hv;
        float kv[8], vv[8];
        cvt_tok8(kh[u], kl[u], P.pad_k, rows[u] < 0, h * HD + dv * 8, kv);
        cvt_tok8(vh[u], vl[u], P.pad_v, rows[u] < 0, h * HD + dv * 8, vv);
        float4* dk = reinterpret_cast<float4*>(sK + ki * HD + dv * 8);
        float4* dvp = reinterpret_cast<float4*>(sV + ki * HD + dv * 8);
        dk[0] = make_float4(kv[0], kv[1], kv[2], kv[3]);
        dk[1] = make_float4(kv[4], kv[5], kv[6], kv[7]);
        dvp[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        dvp[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
      }
    }
    __syncthreads();
    // (inactive threads walk the tile too: with SP = 2 the pair's shuffle needs both lanes; their q is 0)
    for (int c0 = 0; c0 < nk_tile; c0 += 8) {
      float sc8[8];
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const int key = min(c0 + j, nk_tile - 1);
        const float4* kr = reinterpret_cast<const float4*>(sK + key * HD + c_base);
#pragma unroll
        for (int d4 = 0; d4 < HH / 4; ++d4) {
          const float4 kk = kr[d4];
          s0 = fmaf(q[4 * d4], kk.x, s0);
          s1 = fmaf(q[4 * d4 + 1], kk.y, s1);
          s2 = fmaf(q[4 * d4 + 2], kk.z, s2);
          s3 = fmaf(q[4 * d4 + 3], kk.w, s3);
        }
        float sd = (s0 + s1) + (s2 + s3);
        if (SP == 2) sd += __shfl_xor_sync(0xffffffffu, sd, 1);
        sc8[j] = (c0 + j < nk_tile) ? sd : -INFINITY;
        cm = fmaxf(cm, sc8[j]);
      }
      const float nm = fmaxf(m, cm);          // finite: key c0 is always valid
      const float corr = expf(m - nm);        // 0 on the first chunk (m = -inf)
      l *= corr;
#pragma unroll
      for (int d = 0; d < HH; ++d) acc[d] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < nk_tile) {
          const float pw = expf(sc8[j] - nm);
          l += pw;
          const float4* vr = reinterpret_cast<const float4*>(sV + (c0 + j) * HD + c_base);
#pragma unroll
          for (int d4 = 0; d4 < HH / 4; ++d4) {
            const float4 vv = vr[d4];
            acc[4 * d4] = fmaf(pw, vv.x, acc[4 * d4]);
            acc[4 * d4 + 1] = fmaf(pw, vv.y, acc[4 * d4 + 1]);
            acc[4 * d4 + 2] = fmaf(pw, vv.z, acc[4 * d4 + 2]);
            acc[4 * d4 + 3] = fmaf(pw, vv.w, acc[4 * d4 + 3]);
          }
        }
      }
      m = nm;
    }
  }
  if (!active) return;
  long long row;
  if (P.mode == 0) {
    row = static_cast<long long>(b) * P.nq + t;
  } else {                                    // un-partition: the window's token goes back to its grid position
    const int by = b / P.nwx, bx = b - by * P.nwx;
    const int ty = t / wso, tx = t - ty * wso;
    const int y = by * wso + ty, x = bx * wso + tx;
    const int gh = P.pool ? P.grid_h / 2 : P.grid_h, gw = P.pool ? P.grid_w / 2 : P.grid_w;
    if (y >= gh || x >= gw) return;           // padding of the window partition is dropped
    row = static_cast<long long>(y) * gw + x;
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int dv = 0; dv < hvh; ++dv) {
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = acc[dv * 8 + j] * inv;
    store_tok8(P, row * P.o_ct + P.o_off + h * HD + c_base + dv * 8, o8);
  }
}

template <int HD, int SP>
int launch_lpq(const AttnParams& P, cudaStream_t st) {
  const int kt = P.nk < kLpqKeys ? ((P.nk + 7) / 8) * 8 : kLpqKeys;
  const size_t smem = sizeof(float) * 2 * static_cast<size_t>(kt) * HD;
  if (cudaFuncSetAttribute(attention_lpq_kernel<HD, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem)) != cudaSuccess)
    return -1;
  dim3 grid(static_cast<unsigned>(P.B * P.heads), static_cast<unsigned>((P.nq + kLpqQueries - 1) / kLpqQueries));
  attention_lpq_kernel<HD, SP><<<grid, kLpqQueries * SP, smem, st>>>(P, kt);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Few queries against many keys (the mask decoder's token -> image cross attention: 9 queries x 4096 keys per box and
// head).  The general kernel above parallelises over queries and would leave one warp walking all keys; here one block
// owns a (batch, head), keeps the whole score matrix [nq][nk] in shared memory and parallelises over KEYS.
// Exact softmax in fp32 (max-subtracted), same arithmetic as the general kernel up to summation order.
constexpr int kFewQThreads = 512;
template <int HD>
__global__ void __launch_bounds__(kFewQThreads) attention_fewq_kernel(AttnParams P) {
  extern __shared__ float sm[];
  const int nq = P.nq, nk = P.nk;
  constexpr int hd = HD;
  float* sc = sm;                                   // [nq][nk] scores, then unnormalised probabilities
  float* qs = sc + static_cast<size_t>(nq) * nk;    // [nq][hd] scaled queries
  float* lsum = qs + nq * hd;                       // [nq]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / P.heads, h = blockIdx.x - b * P.heads;
  for (int i = tid; i < nq * hd; i += kFewQThreads) {
    const int qi = i / hd, d = i - qi * hd;
    const long long idx = (static_cast<long long>(b) * nq + qi) * P.q_ct + P.q_off + h * hd + d;
    float v = bf16_to_f(P.q[idx]);
    if (P.planes == 2) v += bf16_to_f(P.q[P.q_ps + idx]);
    qs[i] = v * P.scale;
  }
  __syncthreads();
  // scores: a thread takes whole keys (hd <= 32 values in registers) against all queries
  for (int key = tid; key < nk; key += kFewQThreads) {
    float kv[HD];
    const long long row = static_cast<long long>(b) * nk + key;
#pragma unroll
    for (int c = 0; c < hd; c += 8) {
      float t8[8];
      load_tok8(P, P.k, P.k_ct, P.k_off + h * hd, P.k_ps, nullptr, row, c, t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) kv[c + j] = t8[j];
    }
    for (int qi = 0; qi < nq; ++qi) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < hd; ++d) sdot += qs[qi * hd + d] * kv[d];
      sc[static_cast<size_t>(qi) * nk + key] = sdot;
    }
  }
  __syncthreads();
  // softmax statistics: warp w owns query w
  if (warp < nq) {
    float* row = sc + static_cast<size_t>(warp) * nk;
    float m = -INFINITY;
    for (int k2 = lane; k2 < nk; k2 += 32) m = fmaxf(m, row[k2]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int k2 = lane; k2 < nk; k2 += 32) {
      const float pv = __expf(row[k2] - m);
      row[k2] = pv;
      l += pv;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) lsum[warp] = l;
  }
  __syncthreads();
  // output: thread = (key split, query, 8-channel group); 16-byte loads of V, partial sums folded through shared memory
  constexpr int DG = HD / 8;
  const int per = nq * DG;                          // threads per key split
  const int nsplit = min(kFewQThreads / per, nk / hd);   // >= 1; the partials must fit in the score storage
  const int ks = tid / per, r = tid - ks * per;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int qi = r / DG, dg = r - qi * DG;
  if (ks < nsplit) {
    const int k0 = static_cast<int>(static_cast<long long>(nk) * ks / nsplit);
    const int k1 = static_cast<int>(static_cast<long long>(nk) * (ks + 1) / nsplit);
    const float* row = sc + static_cast<size_t>(qi) * nk;
#pragma unroll 2
    for (int key = k0; key < k1; ++key) {
      float v8[8];
      load_tok8(P, P.v, P.v_ct, P.v_off + h * hd, P.v_ps, nullptr, static_cast<long long>(b) * nk + key, dg * 8, v8);
      const float pw = row[key];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += pw * v8[j];
    }
  }
  __syncthreads();                                  // scores are dead from here: reuse their storage for the partials
  float* part = sc;                                 // [nsplit][nq][hd]
  if (ks < nsplit) {
#pragma unroll
    for (int j = 0; j < 8; ++j) part[(ks * nq + qi) * hd + dg * 8 + j] = acc[j];
  }
  __syncthreads();
  if (tid < nq * hd) {
    float o = 0.f;
    for (int s2 = 0; s2 < nsplit; ++s2) o += part[s2 * nq * hd + tid];
    const int q2 = tid / hd, d = tid - q2 * hd;
    o /= lsum[q2];
    const long long idx = (static_cast<long long>(b) * nq + q2) * P.o_ct + P.o_off + h * hd + d;
    uint16_t hi, lo;
    split_bf16(o, hi, lo);
    P.out[idx] = hi;
    if (P.planes == 2) P.out[P.o_ps + idx] = lo;
  }
}


// Many queries against a handful of keys (the mask decoder's image -> token cross attention: 4096 queries x 9 keys per
// box and head).  The general kernel pads the 9 keys to a 64-key tile and keeps 23 of 32 lanes idle; here the K / V rows
// of a batch entry (all heads) sit in shared memory as fp32 and ONE THREAD owns a (query, head): consecutive threads
// read consecutive 2 * HD-byte pieces of a query row (coalesced), the scores never leave registers (the dot products
// are evaluated twice: once for the maximum, once for the weights), output written with 16-byte stores.
constexpr int kFewKThreads = 256;
template <int HD>
__global__ void __launch_bounds__(kFewKThreads) attention_fewk_kernel(AttnParams P) {
  extern __shared__ float sm[];
  const int H = P.heads, Wd = H * HD, nk = P.nk;
  float* sK = sm;                // [nk][H*HD]
  float* sV = sm + nk * Wd;
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < nk * (Wd / 8); i += kFewKThreads) {
    const int key = i / (Wd / 8), c8 = i - key * (Wd / 8);
    float t8[8];
    load_tok8(P, P.k, P.k_ct, P.k_off, P.k_ps, nullptr, static_cast<long long>(b) * nk + key, c8 * 8, t8);
#pragma unroll
    for (int j = 0; j < 8; ++j) sK[key * Wd + c8 * 8 + j] = t8[j];
    load_tok8(P, P.v, P.v_ct, P.v_off, P.v_ps, nullptr, static_cast<long long>(b) * nk + key, c8 * 8, t8);
#pragma unroll
    for (int j = 0; j < 8; ++j) sV[key * Wd + c8 * 8 + j] = t8[j];
  }
  __syncthreads();
  const int qpb = kFewKThreads / H;
  const int q = blockIdx.x * qpb + tid / H, h = tid % H;
  if (tid >= qpb * H || q >= P.nq) return;
  const long long row = static_cast<long long>(b) * P.nq + q;
  float qv[HD];
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float t8[8];
    load_tok8(P, P.q, P.q_ct, P.q_off + h * HD, P.q_ps, nullptr, row, c, t8);
#pragma unroll
    for (int j = 0; j < 8; ++j) qv[c + j] = t8[j] * P.scale;
  }
  float m = -INFINITY;
  for (int key = 0; key < nk; ++key) {
    const float* kr = sK + key * Wd + h * HD;
    float sdot = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) sdot += qv[d] * kr[d];
    m = fmaxf(m, sdot);
  }
  float l = 0.f, acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  for (int key = 0; key < nk; ++key) {
    const float* kr = sK + key * Wd + h * HD;
    const float* vr = sV + key * Wd + h * HD;
    float sdot = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) sdot += qv[d] * kr[d];
    const float pw = expf(sdot - m);
    l += pw;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] += pw * vr[d];
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = acc[c + j] * inv;
    store_tok8(P, row * P.o_ct + P.o_off + h * HD + c, o8);
  }
}

// Few queries against many keys, keys split over a THREAD-BLOCK CLUSTER.  attention_fewq_kernel above gives a whole
// (batch, head) to one CTA: 96 CTAs for 12 boxes x 8 heads, each walking 4096 keys through four block-wide phases —
// 0.19 ms per launch at 0.3 TB/s.  Here the CTAs of a cluster take nk / S keys each, keep their partial softmax state
// (local maximum, local sum, unnormalised weighted V sums) in shared memory, and rank 0 folds the S partials through
// distributed shared memory (exact: every partial is rescaled by exp(m_r - M)).
namespace cg = cooperative_groups;
constexpr int kFewQCThreads = 256;
constexpr int kFewQCSplit = 8;
template <int HD>
__global__ void __launch_bounds__(kFewQCThreads) attention_fewq_cluster_kernel(AttnParams P) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int S = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int bh = blockIdx.x / S;
  const int b = bh / P.heads, h = bh - b * P.heads;
  const int nq = P.nq, nk = P.nk;
  constexpr int hd = HD;
  const int k0 = static_cast<int>(static_cast<long long>(nk) * rank / S);
  const int k1 = static_cast<int>(static_cast<long long>(nk) * (rank + 1) / S);
  const int nkl = k1 - k0, nkmax = (nk + S - 1) / S + 1;
  float* part = sm;                                  // [nq][hd] unnormalised sums, then  [nq] max, [nq] sum
  float* pm = part + nq * hd;
  float* pl = pm + nq;
  float* qs = pl + nq;                               // [nq][hd]
  float* sc = qs + nq * hd;                          // [nq][nkmax]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < nq * hd; i += kFewQCThreads) {
    const int qi = i / hd, d = i - qi * hd;
    const long long idx = (static_cast<long long>(b) * nq + qi) * P.q_ct + P.q_off + h * hd + d;
    float v = bf16_to_f(P.q[idx]);
    if (P.planes == 2) v += bf16_to_f(P.q[P.q_ps + idx]);
    qs[i] = v * P.scale;
    part[i] = 0.f;
  }
  __syncthreads();
  for (int kl = tid; kl < nkl; kl += kFewQCThreads) {
    float kv[HD];
    const long long row = static_cast<long long>(b) * nk + k0 + kl;
#pragma unroll
    for (int c = 0; c < hd; c += 8) {
      float t8[8];
      load_tok8(P, P.k, P.k_ct, P.k_off + h * hd, P.k_ps, nullptr, row, c, t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) kv[c + j] = t8[j];
    }
    for (int qi = 0; qi < nq; ++qi) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < hd; ++d) sdot += qs[qi * hd + d] * kv[d];
      sc[qi * nkmax + kl] = sdot;
    }
  }
  __syncthreads();
  for (int qi = warp; qi < nq; qi += kFewQCThreads / 32) {
    float* row = sc + qi * nkmax;
    float m = -INFINITY;
    for (int k2 = lane; k2 < nkl; k2 += 32) m = fmaxf(m, row[k2]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int k2 = lane; k2 < nkl; k2 += 32) {
      const float pv = expf(row[k2] - m);
      row[k2] = pv;
      l += pv;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) {
      pm[qi] = m;
      pl[qi] = l;
    }
  }
  __syncthreads();
  // weighted V sums: thread = (key split, query, 8-channel group)
  constexpr int DG = HD / 8;
  const int per = nq * DG;
  const int nsplit = kFewQCThreads / per;            // >= 1 (checked by the launcher)
  const int ks = tid / per, r = tid - ks * per;
  if (ks < nsplit) {
    const int qi = r / DG, dg = r - qi * DG;
    const int a0 = static_cast<int>(static_cast<long long>(nkl) * ks / nsplit);
    const int a1 = static_cast<int>(static_cast<long long>(nkl) * (ks + 1) / nsplit);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* row = sc + qi * nkmax;
#pragma unroll 2
    for (int kl = a0; kl < a1; ++kl) {
      float v8[8];
      load_tok8(P, P.v, P.v_ct, P.v_off + h * hd, P.v_ps, nullptr, static_cast<long long>(b) * nk + k0 + kl, dg * 8, v8);
      const float pw = row[kl];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += pw * v8[j];
    }
    // one slot per (split, query, channel) behind the scores; folded in split order below (deterministic)
    float* slot = sc + nq * nkmax + (ks * nq + qi) * hd + dg * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) slot[j] = acc[j];
  }
  __syncthreads();
  for (int i = tid; i < nq * hd; i += kFewQCThreads) {
    float o = 0.f;
    for (int s2 = 0; s2 < nsplit; ++s2) o += sc[nq * nkmax + s2 * nq * hd + i];
    part[i] = o;
  }
  cluster.sync();
  if (rank == 0) {
    for (int i = tid; i < nq * hd; i += kFewQCThreads) {
      const int qi = i / hd, d = i - qi * hd;
      float M = -INFINITY;
      for (int r2 = 0; r2 < S; ++r2) M = fmaxf(M, cluster.map_shared_rank(pm, r2)[qi]);
      float num = 0.f, den = 0.f;
      for (int r2 = 0; r2 < S; ++r2) {
        const float w = expf(cluster.map_shared_rank(pm, r2)[qi] - M);
        num += w * cluster.map_shared_rank(part, r2)[i];
        den += w * cluster.map_shared_rank(pl, r2)[qi];
      }
      const float o = num / den;
      const long long idx = (static_cast<long long>(b) * nq + qi) * P.o_ct + P.o_off + h * hd + d;
      uint16_t hi, lo;
      split_bf16(o, hi, lo);
      P.out[idx] = hi;
      if (P.planes == 2) P.out[P.o_ps + idx] = lo;
    }
  }
  cluster.sync();                                    // the other ranks' shared memory stays alive until rank 0 has read it
}

template <int HD>
int launch_fewq_cluster(const AttnParams& P, cudaStream_t st) {
  const int nkmax = (P.nk + kFewQCSplit - 1) / kFewQCSplit + 1;
  const int nsplit = kFewQCThreads / (P.nq * (HD / 8));
  const size_t smem = sizeof(float) * (static_cast<size_t>(P.nq) * HD * 2 + 2 * P.nq + static_cast<size_t>(P.nq) * nkmax +
                                       static_cast<size_t>(nsplit) * P.nq * HD);
  if (smem > 96 * 1024) return 1;
  cudaError_t e = cudaFuncSetAttribute(attention_fewq_cluster_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(P.B * P.heads * kFewQCSplit));
  cfg.blockDim = dim3(kFewQCThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kFewQCSplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, attention_fewq_cluster_kernel<HD>, P);
  return e == cudaSuccess ? 0 : -1;
}

// ---- patch embedding: (u8 - mean')/std' -> conv k x k / stride / pad (3 -> C) + bias + positional embedding -------
// one thread per (output pixel, 8 output channels); weights in shared memory
// G output channels per thread; a warp = 32 consecutive output pixels of ONE channel group, so the weights (stored
// tap-major in shared memory: [3*k*k][C]) are read with 16-byte broadcast loads and every normalised input value
// ((u8 - mean) / std: a division) feeds G FMAs instead of 8.  Per channel the taps are accumulated in the same order
// (c, ky, kx) as before: results are bit-identical to the round-1 kernel (G = 8, lane = channel group).
template <int G>
__global__ void patch_embed_kernel(const uint8_t* __restrict__ img, int H, int W, const float* __restrict__ mean,
                                   const float* __restrict__ stdv, const float* __restrict__ w, const float* __restrict__ b,
                                   const float* __restrict__ pos, int C, int k, int stride, int pad, int Ho, int Wo,
                                   uint16_t* __restrict__ out, long long ps, int planes) {
  extern __shared__ __align__(16) float sw[];  // [3*k*k][C]
  const int kk = 3 * k * k;
  for (int i = threadIdx.x; i < C * kk; i += blockDim.x) {
    const int ch = i / kk, wi = i - ch * kk;
    sw[wi * C + ch] = w[i];
  }
  __syncthreads();
  const int groups = C / G;
  const long long npix = static_cast<long long>(Ho) * Wo;
  const long long chunks = (npix + 31) / 32 * groups;        // warp-sized work items
  const float m0 = mean[0], m1 = mean[1], m2 = mean[2], s0 = stdv[0], s1 = stdv[1], s2 = stdv[2];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long wk = warp0; wk < chunks; wk += nwarps) {
    const int g = static_cast<int>(wk % groups);
    const long long pix = (wk / groups) * 32 + lane;
    if (pix >= npix) continue;
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>(pix / Wo);
    float acc[G];
#pragma unroll
    for (int j = 0; j < G; ++j) acc[j] = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float mm = c == 0 ? m0 : (c == 1 ? m1 : m2), ss = c == 0 ? s0 : (c == 1 ? s1 : s2);
      for (int ky = 0; ky < k; ++ky) {
        const int iy = oy * stride + ky - pad;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int ix = ox * stride + kx - pad;
          if (ix < 0 || ix >= W) continue;
          const float v = (static_cast<float>(img[(static_cast<long long>(iy) * W + ix) * 3 + c]) - mm) / ss;
          const float4* wr = reinterpret_cast<const float4*>(sw + ((c * k + ky) * k + kx) * C + g * G);
#pragma unroll
          for (int j4 = 0; j4 < G / 4; ++j4) {
            const float4 ww = wr[j4];
            acc[4 * j4] += v * ww.x;
            acc[4 * j4 + 1] += v * ww.y;
            acc[4 * j4 + 2] += v * ww.z;
            acc[4 * j4 + 3] += v * ww.w;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int ch = g * G + j;
      st_planes(out, pix * C + ch, ps, planes, acc[j] + b[ch] + pos[pix * C + ch]);
    }
  }
}

// ---- box prompts -> 3 sparse tokens per box (Sam2PromptEncoder._embed_boxes) -------------------------------------
// boxes: [P][4] already scaled to the 1024 input frame; gauss: [2][half]; out fp32 [P][3][2*half]
__global__ void prompt_boxes_kernel(const float* __restrict__ boxes, float sx, float sy, int Pn,
                                    const float* __restrict__ gauss, int half,
                                    const float* __restrict__ pe2, const float* __restrict__ pe3,
                                    const float* __restrict__ not_a_point, float input_size, float* __restrict__ out) {
  const int total = Pn * 3 * 2 * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % (2 * half);
    const int t = (i / (2 * half)) % 3;
    const int p = i / (6 * half);
    float v;
    if (t == 2) {
      v = not_a_point[c];
    } else {
      // processor: coords * (1024 / W) in fp32 (processing_sam2.py:196-197); prompt encoder: + 0.5, / 1024
      float x = (boxes[p * 4 + 2 * t] * sx + 0.5f) / input_size;
      float y = (boxes[p * 4 + 2 * t + 1] * sy + 0.5f) / input_size;
      x = 2.0f * x - 1.0f;
      y = 2.0f * y - 1.0f;
      const int cc = c < half ? c : c - half;
      float proj = x * gauss[cc] + y * gauss[half + cc];
      proj = 6.283185307179586f * proj;
      v = (c < half ? sinf(proj) : cosf(proj)) + (t == 0 ? pe2[c] : pe3[c]);
    }
    out[i] = v;
  }
}

// ---- masks[p][k][pix] = sum_c hyper[p][k][c] * up[p][pix][c] -------------------------------------------------------
__global__ void hyper_masks_kernel(const uint16_t* __restrict__ up, long long ps, int planes, const float* __restrict__ hyper,
                                   int Pn, int K, int C, long long npix, float* __restrict__ out) {
  const long long total = static_cast<long long>(Pn) * npix;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i / npix);
    const long long pix = i - static_cast<long long>(p) * npix;
    float u[32];
    for (int c = 0; c < C; ++c) u[c] = ld_planes(up, i * C + c, ps, planes);
    for (int k = 0; k < K; ++k) {
      const float* hk = hyper + (static_cast<long long>(p) * K + k) * C;
      float a = 0.f;
      for (int c = 0; c < C; ++c) a += hk[c] * u[c];
      out[(static_cast<long long>(p) * K + k) * npix + pix] = a;
    }
  }
}

// ---- _dynamic_multimask_via_stability: one block per box ----------------------------------------------------------
__global__ void select_mask_kernel(const float* __restrict__ logits, const float* __restrict__ iou, int K, long long npix,
                                   float delta, float thresh, int* __restrict__ sel) {
  const int p = blockIdx.x;
  const float* m0 = logits + static_cast<long long>(p) * K * npix;
  __shared__ unsigned int s_i, s_u;
  if (threadIdx.x == 0) {
    s_i = 0;
    s_u = 0;
  }
  __syncthreads();
  unsigned int ai = 0, au = 0;
  for (long long i = threadIdx.x; i < npix; i += blockDim.x) {
    const float v = m0[i];
    ai += v > delta;
    au += v > -delta;
  }
  atomicAdd(&s_i, ai);
  atomicAdd(&s_u, au);
  __syncthreads();
  if (threadIdx.x == 0) {
    const float area_i = static_cast<float>(s_i), area_u = static_cast<float>(s_u);
    const float stab = area_u > 0.f ? area_i / area_u : 1.0f;
    int best = 0;
    if (!(stab >= thresh)) {
      best = 1;
      for (int k = 2; k < K; ++k)
        if (iou[p * K + k] > iou[p * K + best]) best = k;  // torch.argmax: first maximum
    }
    sel[p] = best;
  }
}

// ---- bilinear 256^2 -> HxW (align_corners=False), > 0, AND the floor/ceil box, write uint8 {0,255} --------------
// One thread per output pixel of the box window; pixels outside every box stay 0 (the mask buffer is zeroed here too).
__global__ void mask_write_kernel(const float* __restrict__ logits, const int* __restrict__ sel, int K, int S,
                                  const float* __restrict__ boxes /* [P][4] original px */, int Pn, int H, int W,
                                  uint8_t* __restrict__ masks /* [P][H][W] */, float* __restrict__ logit_out) {
  const long long per = static_cast<long long>(H) * W;
  const long long total = per * Pn;
  const float sy = static_cast<float>(S) / static_cast<float>(H), sx = static_cast<float>(S) / static_cast<float>(W);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i / per);
    const long long r = i - static_cast<long long>(p) * per;
    const int y = static_cast<int>(r / W), x = static_cast<int>(r - static_cast<long long>(y) * W);
    const float* b = boxes + p * 4;
    // clip rectangle: floor(x0), floor(y0), ceil(x1), ceil(y1) clamped to the page (detection.py:1733-1737)
    const int bx0 = static_cast<int>(floorf(fminf(fmaxf(b[0], 0.f), static_cast<float>(W))));
    const int by0 = static_cast<int>(floorf(fminf(fmaxf(b[1], 0.f), static_cast<float>(H))));
    const int bx1 = static_cast<int>(ceilf(fminf(fmaxf(b[2], 0.f), static_cast<float>(W))));
    const int by1 = static_cast<int>(ceilf(fminf(fmaxf(b[3], 0.f), static_cast<float>(H))));
    uint8_t o = 0;
    const bool inside = x >= bx0 && x < bx1 && y >= by0 && y < by1;
    if (inside || logit_out) {
      float fy = sy * (static_cast<float>(y) + 0.5f) - 0.5f;
      float fx = sx * (static_cast<float>(x) + 0.5f) - 0.5f;
      if (fy < 0.f) fy = 0.f;
      if (fx < 0.f) fx = 0.f;
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
      const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
      const float* m = logits + (static_cast<long long>(p) * K + sel[p]) * S * S;
      const float top = (1.f - lx) * m[y0 * S + x0] + lx * m[y0 * S + x1];
      const float bot = (1.f - lx) * m[y1 * S + x0] + lx * m[y1 * S + x1];
      const float v = (1.f - ly) * top + ly * bot;
      if (logit_out) logit_out[i] = v;
      if (inside && v > 0.0f) o = 255;
    }
    masks[i] = o;
  }
}

}  // namespace

extern "C" {

int mtb_layernorm(const void* x, long long rows, int C, int ct_in, int ci, int planes_in, const float* gamma,
                  const float* beta, float eps, void* y, int ct_out, int co, int planes_out, int gelu, void* stream) {
  MTB_REQUIRE(x && y && gamma && beta && rows >= 0, "mtb_layernorm: bad arguments");
  if (rows == 0) return 0;
  const int block = 256;
  layernorm_kernel<<<grid_for4(rows, block / 32), block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), rows, C, ct_in, ci, rows * ct_in, planes_in, gamma, beta, eps,
      static_cast<uint16_t*>(y), ct_out, co, rows * ct_out, planes_out, gelu);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_maxpool2x2(const void* x, void* y, int N, int H, int W, int C, int planes, void* stream) {
  MTB_REQUIRE(x && y && (H % 2 == 0) && (W % 2 == 0), "mtb_maxpool2x2: bad arguments");
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * C;
  maxpool2x2_kernel<<<grid_for4(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), N, H, W, C, static_cast<long long>(N) * H * W * C, total,
      planes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_add_planes(const void* a, const void* b, void* out, long long rows, int C, long long b_rows, int planes,
                   void* stream) {
  MTB_REQUIRE(a && b && out && b_rows > 0, "mtb_add_planes: bad arguments");
  add_planes_kernel<<<grid_for4(rows * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(a), static_cast<const uint16_t*>(b), static_cast<uint16_t*>(out), rows, C, b_rows,
      rows * C, b_rows * C, rows * C, planes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

long long mtb_attention_workspace_bytes(int B, int heads, int hd, int nk) {
  return mtb::attention_tc_workspace_bytes(B, heads, hd, nk);
}

int mtb_attention(const mtb_attn_desc* d, void* stream) {
  MTB_REQUIRE(d && d->q && d->k && d->v && d->out, "mtb_attention: null argument");
  {
    const int rc = mtb::launch_attention_tc(d, static_cast<cudaStream_t>(stream));
    if (rc < 0) return rc;
    if (rc == 0) {
      g_launches.fetch_add(2);   // V transpose + attention
      return 0;
    }
  }
  MTB_REQUIRE(d->hd >= 1 && d->hd <= kMaxHD, "mtb_attention: head dim %d not supported (<= %d)", d->hd, kMaxHD);
  MTB_REQUIRE(!(d->mode == 1 && d->pool && (d->ws & 1)), "mtb_attention: pooled windows need an even window size");
  AttnParams P;
  P.B = d->B;
  P.heads = d->heads;
  P.hd = d->hd;
  P.nq = d->nq;
  P.nk = d->nk;
  P.scale = d->scale;
  P.q = static_cast<const uint16_t*>(d->q);
  P.k = static_cast<const uint16_t*>(d->k);
  P.v = static_cast<const uint16_t*>(d->v);
  P.out = static_cast<uint16_t*>(d->out);
  P.q_ct = d->q_ct; P.q_off = d->q_off; P.k_ct = d->k_ct; P.k_off = d->k_off;
  P.v_ct = d->v_ct; P.v_off = d->v_off; P.o_ct = d->o_ct; P.o_off = d->o_off;
  P.q_ps = d->q_ps; P.k_ps = d->k_ps; P.v_ps = d->v_ps; P.o_ps = d->o_ps;
  P.planes = d->planes;
  P.mode = d->mode;
  P.grid_h = d->grid_h; P.grid_w = d->grid_w; P.ws = d->ws; P.pool = d->pool;
  P.nwx = d->ws > 0 ? (d->grid_w + d->ws - 1) / d->ws : 1;
  P.pad_q = d->pad_q; P.pad_k = d->pad_k; P.pad_v = d->pad_v;
  const bool aligned8 = ((d->q_ct | d->k_ct | d->v_ct | d->o_ct | d->q_off | d->k_off | d->v_off | d->o_off) % 8) == 0 &&
                        ((d->q_ps | d->k_ps | d->v_ps | d->o_ps) % 8) == 0;
  static const bool use_fewk = !getenv("MTB200_ATTN_FEWK") || atoi(getenv("MTB200_ATTN_FEWK")) != 0;
  static const bool use_cluster = !getenv("MTB200_ATTN_CLUSTER") || atoi(getenv("MTB200_ATTN_CLUSTER")) != 0;
  if (use_fewk && d->mode == 0 && d->nk <= 32 && d->nq >= 256 && (d->hd == 16 || d->hd == 32) && aligned8 &&
      d->heads >= 1 && d->heads <= kFewKThreads &&
      static_cast<size_t>(d->nk) * d->heads * d->hd * 2 * sizeof(float) <= 48 * 1024) {
    const size_t ksmem = static_cast<size_t>(d->nk) * d->heads * d->hd * 2 * sizeof(float);
    const int qpb = kFewKThreads / d->heads;
    dim3 grid(static_cast<unsigned>((d->nq + qpb - 1) / qpb), static_cast<unsigned>(d->B));
    if (d->hd == 16) attention_fewk_kernel<16><<<grid, kFewKThreads, ksmem, static_cast<cudaStream_t>(stream)>>>(P);
    else attention_fewk_kernel<32><<<grid, kFewKThreads, ksmem, static_cast<cudaStream_t>(stream)>>>(P);
    MTB_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
  }
  if (use_cluster && d->mode == 0 && d->nq <= 16 && d->nk >= 512 && (d->hd == 16 || d->hd == 32) && aligned8 &&
      d->nq * (d->hd / 8) <= kFewQCThreads) {
    const int rc = d->hd == 16 ? launch_fewq_cluster<16>(P, static_cast<cudaStream_t>(stream))
                               : launch_fewq_cluster<32>(P, static_cast<cudaStream_t>(stream));
    MTB_REQUIRE(rc >= 0, "mtb_attention: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == 0) {
      g_launches.fetch_add(1);
      return 0;
    }
  }
  if (d->mode == 0 && d->nq <= 16 && d->nk >= 512 && (d->hd == 16 || d->hd == 32) && d->nq * d->hd <= kFewQThreads) {
    const size_t fsmem = sizeof(float) * (static_cast<size_t>(d->nq) * d->nk + d->nq * d->hd + 16);
    if (fsmem <= 200 * 1024) {
      if (d->hd == 16) {
        MTB_CUDA_OK(cudaFuncSetAttribute(attention_fewq_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(fsmem)));
        attention_fewq_kernel<16><<<d->B * d->heads, kFewQThreads, fsmem, static_cast<cudaStream_t>(stream)>>>(P);
      } else {
        MTB_CUDA_OK(cudaFuncSetAttribute(attention_fewq_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(fsmem)));
        attention_fewq_kernel<32><<<d->B * d->heads, kFewQThreads, fsmem, static_cast<cudaStream_t>(stream)>>>(P);
      }
      MTB_CUDA_OK(cudaGetLastError());
      g_launches.fetch_add(1);
      return 0;
    }
  }
  static const bool use_lpq = !getenv("MTB200_ATTN_LPQ") || atoi(getenv("MTB200_ATTN_LPQ")) != 0;
  if (use_lpq && aligned8 && !(d->mode == 1 && d->pool && (d->ws & 1)) &&
      (d->hd == 32 || d->hd == 64 || d->hd == 72 || d->hd == 96)) {
    const cudaStream_t cs = static_cast<cudaStream_t>(stream);
    static const int split = getenv("MTB200_ATTN_LPQ_SPLIT") ? atoi(getenv("MTB200_ATTN_LPQ_SPLIT")) : 2;
    const int rc = d->hd == 32 ? launch_lpq<32, 1>(P, cs)
                 : d->hd == 64 ? (split == 2 ? launch_lpq<64, 2>(P, cs) : launch_lpq<64, 1>(P, cs))
                 : d->hd == 72 ? launch_lpq<72, 1>(P, cs)
                               : (split == 2 ? launch_lpq<96, 2>(P, cs) : launch_lpq<96, 1>(P, cs));
    MTB_REQUIRE(rc == 0, "mtb_attention: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    return 0;
  }
  const size_t smem = sizeof(float) * (static_cast<size_t>(kKT) * (d->hd + 1) + static_cast<size_t>(kKT) * d->hd +
                                       static_cast<size_t>(kQT) * d->hd);
  MTB_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>(d->B * d->heads), static_cast<unsigned>((d->nq + kQT - 1) / kQT));
  attention_kernel<<<grid, kAttnThreads, smem, static_cast<cudaStream_t>(stream)>>>(P);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}


int mtb_sam_patch_embed(const uint8_t* img, int H, int W, const float* mean3, const float* std3, const float* w,
                        const float* b, const float* pos, int C, int k, int stride, int pad, void* out, int planes,
                        void* stream) {
  MTB_REQUIRE(img && mean3 && std3 && w && b && pos && out && C % 8 == 0, "mtb_sam_patch_embed: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t smem = sizeof(float) * static_cast<size_t>(C) * 3 * k * k;
  const long long npix = static_cast<long long>(Ho) * Wo;
  const cudaStream_t cs = static_cast<cudaStream_t>(stream);
  uint16_t* o = static_cast<uint16_t*>(out);
  if (C % 32 == 0) {
    MTB_CUDA_OK(cudaFuncSetAttribute(patch_embed_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    patch_embed_kernel<32><<<grid_for4(npix * (C / 32), 128), 128, smem, cs>>>(img, H, W, mean3, std3, w, b, pos, C, k, stride,
                                                                              pad, Ho, Wo, o, npix * C, planes);
  } else if (C % 16 == 0) {
    MTB_CUDA_OK(cudaFuncSetAttribute(patch_embed_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    patch_embed_kernel<16><<<grid_for4(npix * (C / 16), 128), 128, smem, cs>>>(img, H, W, mean3, std3, w, b, pos, C, k, stride,
                                                                              pad, Ho, Wo, o, npix * C, planes);
  } else {
    MTB_CUDA_OK(cudaFuncSetAttribute(patch_embed_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    patch_embed_kernel<8><<<grid_for4(npix * (C / 8), 128), 128, smem, cs>>>(img, H, W, mean3, std3, w, b, pos, C, k, stride,
                                                                            pad, Ho, Wo, o, npix * C, planes);
  }
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_sam_prompt_boxes(const float* boxes, float sx, float sy, int P, const float* gauss, int half, const float* pe2, const float* pe3,
                         const float* not_a_point, float input_size, float* out, void* stream) {
  MTB_REQUIRE(boxes && gauss && pe2 && pe3 && not_a_point && out, "mtb_sam_prompt_boxes: null argument");
  if (P <= 0) return 0;
  prompt_boxes_kernel<<<grid_for4(static_cast<long long>(P) * 6 * half, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, sx, sy, P, gauss, half, pe2, pe3, not_a_point, input_size, out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_sam_hyper_masks(const void* up, int planes, const float* hyper, int P, int K, int C, long long npix, float* out,
                        void* stream) {
  MTB_REQUIRE(up && hyper && out && C <= 32, "mtb_sam_hyper_masks: bad arguments");
  if (P <= 0) return 0;
  hyper_masks_kernel<<<grid_for4(static_cast<long long>(P) * npix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(up), static_cast<long long>(P) * npix * C, planes, hyper, P, K, C, npix, out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_sam_select_mask(const float* logits, const float* iou, int P, int K, long long npix, float delta, float thresh,
                        int* sel, void* stream) {
  MTB_REQUIRE(logits && iou && sel, "mtb_sam_select_mask: null argument");
  if (P <= 0) return 0;
  select_mask_kernel<<<P, 512, 0, static_cast<cudaStream_t>(stream)>>>(logits, iou, K, npix, delta, thresh, sel);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_sam_mask_write(const float* logits, const int* sel, int K, int S, const float* boxes, int P, int H, int W,
                       uint8_t* masks, float* logit_out, void* stream) {
  MTB_REQUIRE(logits && sel && boxes && masks, "mtb_sam_mask_write: null argument");
  if (P <= 0) return 0;
  mask_write_kernel<<<grid_for4(static_cast<long long>(P) * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, sel, K, S, boxes, P, H, W, masks, logit_out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
