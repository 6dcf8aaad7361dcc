// Shared conv epilogue: 16 accumulator columns of one pixel (one TMEM lane) -> bias, activation, residual,
// per-tile channel sums, hi/lo bf16 split, vector stores.  Used by conv_gemm.cu and conv_halo.cu.
// The activation is a template parameter so the hot loop has no switch; ACT == -1 takes it from ConvParams.
#pragma once
#include "conv_gemm.cuh"

namespace mtb {

template <int ACT>
__device__ __forceinline__ float epi_act(float v, int act_rt) {
  const int a = (ACT >= 0) ? ACT : act_rt;
  if (a == ACT_RELU) return fmaxf(v, 0.0f);
  if (a == ACT_SILU) return v / (1.0f + expf(-v));
  if (a == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (a == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

// two floats -> packed bf16 pair (round to nearest even), element 0 in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int ACT>
__device__ __forceinline__ void epilogue_chunk16(const ConvParams& p, float (&v)[16], bool valid, long long pix, int cbase,
                                                 int n, int oy, int ox, int lane, int q, long long mtile,
                                                 float& cta_sum) {
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + cbase);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j] += b.x;
      v[4 * j + 1] += b.y;
      v[4 * j + 2] += b.z;
      v[4 * j + 3] += b.w;
    }
  }
  if (p.chan_scale) {
    const float4* s4 = reinterpret_cast<const float4*>(p.chan_scale + cbase);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b = __ldg(s4 + j);
      v[4 * j] *= b.x;
      v[4 * j + 1] *= b.y;
      v[4 * j + 2] *= b.z;
      v[4 * j + 3] *= b.w;
    }
  }
  // output location (element offset inside a plane) and the matching residual location
  long long off = pix * p.out_cstride + p.out_coff + cbase;
  long long roff = ((p.res_bcast ? (static_cast<long long>(oy) * p.Wo + ox) : pix)) * p.res_cstride + p.res_coff + cbase;
  if (p.pixel_shuffle) {
    // PixelShuffle(2) / ConvTranspose2d(2,2) fused into the store: channel block b = dy*2+dx lands on the 2x grid
    const int cq = p.Cout >> 2;
    const int blk = cbase / cq, cc = cbase - blk * cq;
    const long long hy = 2 * oy + (blk >> 1), hx = 2 * ox + (blk & 1);
    off = ((static_cast<long long>(n) * (2 * p.Ho) + hy) * (2 * p.Wo) + hx) * cq + cc;
    roff = ((static_cast<long long>(p.res_bcast ? 0 : n) * (2 * p.Ho) + hy) * (2 * p.Wo) + hx) * p.res_cstride + p.res_coff + cc;
  }
  if (!p.act_after_res) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = epi_act<ACT>(v[j], p.act);
  }
  if (p.residual && valid) {
    const uint16_t* rp = p.residual + roff;
    for (int pl = 0; pl < p.res_planes; ++pl) {
      const uint4* r4 = reinterpret_cast<const uint4*>(rp + pl * p.res_plane_stride);
      const uint4 a = __ldg(r4), b = __ldg(r4 + 1);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        v[2 * j] += f.x;
        v[2 * j + 1] += f.y;
      }
    }
  }
  if (p.act_after_res) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = epi_act<ACT>(v[j], p.act);
  }
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.0f;
  }
  if (p.tile_sums) {
    // channel sums over the warp's 32 pixels: fold the two half-warps, then a 16-lane transpose-reduce; lane l < 16
    // ends with the sum of channel cbase + l
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = v[j] + __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
      const bool upper = (lane & w) != 0;
#pragma unroll
      for (int j = 0; j < w; ++j) {
        const float send = upper ? s[j] : s[j + w];
        const float keep = upper ? s[j + w] : s[j];
        s[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
      }
    }
    if (p.sums_per_cta) cta_sum += s[0];   // fixed tile order per CTA -> deterministic
    else if (lane < 16) p.tile_sums[(mtile * 4 + q) * p.Cout + cbase + lane] = s[0];
  }
  if (!valid) return;
  if (p.out_f32) {
    float4* o4 = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    return;
  }
  uint32_t hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
  o4[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  o4[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  if (p.planes_out == 2) {
    uint32_t lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 h = unpack_bf16x2(hi[j]);
      lo[j] = pack_bf16x2(v[2 * j] - h.x, v[2 * j + 1] - h.y);
    }
    uint4* l4 = reinterpret_cast<uint4*>(p.out + p.out_plane_stride + off);
    l4[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    l4[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
}

constexpr int kEpiWarps = 16;                       // 4 per TMEM lane quarter
constexpr int kConvThreads = (2 + kEpiWarps) * 32;  // TMA warp + MMA warp + epilogue warps

}  // namespace mtb
