// Full (global) softmax attention on the tensor cores: SAM 2.1 Hiera global-attention blocks
// (reference: transformers Sam2MultiScaleAttention behind core/image/detection.py:475-511; 64x64 = 4096 tokens, 4 heads
// of 96 channels in the tiny model), fp32-grade via bf16x3 operands.
//
//   S = Q K^T   : tcgen05.mma M=128 queries x N=64 keys, K = head dim; three products (Qh Kh, Qh Kl, Ql Kh)
//   P = exp2((S - m) * scale*log2e)    softmax warps: one thread per query row, S read from TMEM
//   O += P V    : M=128 x N=head dim, K = 64 keys; P is written to shared memory as bf16 hi/lo in the 128B-swizzled
//                 K-major layout the MMA reads; V comes pre-transposed ([d][token], K-major) from attn_vt_kernel
//
// Exact softmax without rescaling the accumulator: TWO passes over the keys.  Pass A only finds each row's maximum m
// (S tiles are produced and max-reduced, double-buffered in TMEM); pass B recomputes S, forms P with the final m and
// accumulates l = sum(P) and O.  That costs the QK^T MMAs twice — cheap next to the 2 ms the CUDA-core kernel needs
// for such a block — and keeps the pipeline free of the correction step.
//
// CTA = one 128-query tile of one (batch, head); 6 warps: TMA producer, MMA issuer, 4 softmax warps (TMEM lane
// quarters).  Shared memory (head dim 96): Q 64 KB + K ring 64 KB + V^T ring 48 KB + P 32 KB.
#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {

namespace {

constexpr int kQTile = 128;
constexpr int kKTile = 64;
constexpr int kTcThreads = 192;
constexpr int kQBlk = kQTile * 128;   // one 64-channel block of the Q tile (bytes)
constexpr int kKBlk = kKTile * 128;   // one 64-channel block of a K tile
constexpr int kPBlk = kQTile * 128;   // one plane of P (128 rows x 64 keys)

struct TcParams {
  int B, heads, nq, nk;
  float c;  // scale * log2(e)
  int q_off, k_off;
  uint16_t* out;
  int o_ct, o_off;
  long long o_ps;
};

// v planes [token][ct] -> vt [plane][b*heads+h][d][token] (keys contiguous), zero beyond nk
__global__ void attn_vt_kernel(const uint16_t* __restrict__ v, int v_ct, int v_off, long long v_ps, int planes, int heads,
                               int hd, int nk, int nk_pad, uint16_t* __restrict__ vt) {
  extern __shared__ uint16_t tile[];   // [hd][64 + 2]
  const int tok0 = blockIdx.x * 64;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  const int BH = gridDim.y;
  for (int pl = 0; pl < planes; ++pl) {
    for (int i = threadIdx.x; i < 64 * hd; i += blockDim.x) {
      const int tk = i / hd, d = i - tk * hd;
      const int tok = tok0 + tk;
      uint16_t val = 0;
      if (tok < nk) val = v[pl * v_ps + (static_cast<long long>(b) * nk + tok) * v_ct + v_off + h * hd + d];
      tile[d * 66 + tk] = val;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * hd; i += blockDim.x) {
      const int d = i >> 6, tk = i & 63;
      if (tok0 + tk < nk_pad)
        vt[((static_cast<long long>(pl) * BH + bh) * hd + d) * nk_pad + tok0 + tk] = tile[d * 66 + tk];
    }
    __syncthreads();
  }
}

template <int HD>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const TcParams p) {
  constexpr int NKB = (HD + 63) / 64;          // 64-channel blocks of the head dim
  constexpr int kVPlane = HD * 128;            // one plane of a V^T tile: HD rows x 64 keys
  constexpr int kKStage = 2 * NKB * kKBlk;
  constexpr int kVStage = 2 * kVPlane;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  uint8_t* sQ = smem;                                   // [plane][kb][128 x 128 B]
  uint8_t* sK = sQ + 2 * NKB * kQBlk;                   // [stage][plane][kb][64 x 128 B]
  uint8_t* sV = sK + 2 * kKStage;                       // [stage][plane][HD x 128 B]
  uint8_t* sP = sV + 2 * kVStage;                       // [plane][128 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBlk);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* p_empty = bars + 14;
  uint64_t* o_full = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.nk + kKTile - 1) / kKTile;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // S buffers at columns 0 and 64, O at column 128

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * NKB * kQBlk);
      for (int pl = 0; pl < 2; ++pl)
        for (int kb = 0; kb < NKB; ++kb)
          tma_load_3d(sQ + (pl * NKB + kb) * kQBlk, &tmQ, q_full, p.q_off + h * HD + kb * 64, b * p.nq + q0, pl);
    }
    __syncwarp();
    int ks = 0, vs = 0;
    uint32_t kphase = 0, vphase = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < T; ++j) {
        mbar_wait(&k_empty[ks], kphase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[ks], kKStage);
          for (int pl = 0; pl < 2; ++pl)
            for (int kb = 0; kb < NKB; ++kb)
              tma_load_3d(sK + ks * kKStage + (pl * NKB + kb) * kKBlk, &tmK, &k_full[ks], p.k_off + h * HD + kb * 64,
                          b * p.nk + j * kKTile, pl);
        }
        __syncwarp();
        if (++ks == 2) {
          ks = 0;
          kphase ^= 1;
        }
        if (pass == 1) {
          mbar_wait(&v_empty[vs], vphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&v_full[vs], kVStage);
            for (int pl = 0; pl < 2; ++pl)
              tma_load_4d(sV + vs * kVStage + pl * kVPlane, &tmV, &v_full[vs], j * kKTile, 0, b * p.heads + h, pl);
          }
          __syncwarp();
          if (++vs == 2) {
            vs = 0;
            vphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    constexpr uint32_t idS = make_idesc_bf16(kQTile, kKTile);
    constexpr uint32_t idO = make_idesc_bf16(kQTile, HD);
    const uint32_t aQ = sbase, aK = aQ + 2 * NKB * kQBlk, aV = aK + 2 * kKStage, aP = aV + 2 * kVStage;
    mbar_wait(q_full, 0);
    tc_fence_after();
    int ks = 0, vs = 0;
    uint32_t kphase = 0, vphase = 0, pphase = 0;
    int g = 0;   // S tiles issued so far (buffer g & 1, use count g >> 1)
    auto issue_s = [&]() {
      const int sb = g & 1;
      mbar_wait(&k_full[ks], kphase);
      mbar_wait(&s_empty[sb], ((g >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + sb * kKTile;
        bool first = true;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const int pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;   // (Qh,Kh) (Qh,Kl) (Ql,Kh)
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            constexpr int kFull = 4;
            const int steps = (HD - kb * 64) >= 64 ? kFull : (HD - kb * 64) / 16;
#pragma unroll
            for (int k = 0; k < kFull; ++k) {
              if (k < steps) {
                const uint64_t da = make_sdesc_sw128(aQ + (pa * NKB + kb) * kQBlk + k * 32, 1024, 0);
                const uint64_t db = make_sdesc_sw128(aK + ks * kKStage + (pb * NKB + kb) * kKBlk + k * 32, 1024, 0);
                umma_bf16(d_tmem, da, db, idS, first ? 0u : 1u);
                first = false;
              }
            }
          }
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[ks]);
      }
      __syncwarp();
      if (++ks == 2) {
        ks = 0;
        kphase ^= 1;
      }
      ++g;
    };
    for (int j = 0; j < T; ++j) issue_s();      // pass A: row maxima
    issue_s();                                  // pass B, tile 0
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_s();                 // S of the next tile runs under this tile's softmax
      mbar_wait(p_full, pphase);
      mbar_wait(&v_full[vs], vphase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + 128;
        bool first = true;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const int pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;   // (Ph,Vh) (Ph,Vl) (Pl,Vh)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_sdesc_sw128(aP + pa * kPBlk + k * 32, 1024, 0);
            const uint64_t db = make_sdesc_sw128(aV + vs * kVStage + pb * kVPlane + k * 32, 1024, 0);
            umma_bf16(d_tmem, da, db, idO, (j > 0 || !first) ? 1u : 0u);
            first = false;
          }
        }
        umma_commit(p_empty);
        umma_commit(&v_empty[vs]);
      }
      __syncwarp();
      pphase ^= 1;
      if (++vs == 2) {
        vs = 0;
        vphase ^= 1;
      }
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else {
    // ------------------------------- softmax / epilogue warps -------------------------------
    const int q = warp & 3;                   // TMEM lane quarter
    const int r = q * 32 + lane;              // row of the query tile
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m = -INFINITY;
    int g = 0;
    for (int j = 0; j < T; ++j, ++g) {        // pass A
      const int sb = g & 1;
      mbar_wait(&s_full[sb], (g >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t a[16];
        tmem_ld16(tlane + sb * kKTile + c * 16, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (j * kKTile + c * 16 + i < p.nk) m = fmaxf(m, __uint_as_float(a[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    float l = 0.f;
    uint8_t* prow = sP + r * 128;
    for (int j = 0; j < T; ++j, ++g) {        // pass B
      const int sb = g & 1;
      mbar_wait(&s_full[sb], (g >> 1) & 1);
      tc_fence_after();
      float pv[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t a[16];
        tmem_ld16(tlane + sb * kKTile + c * 16, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) pv[c * 16 + i] = __uint_as_float(a[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float e = (j * kKTile + i < p.nk) ? exp2f((pv[i] - m) * p.c) : 0.f;
        pv[i] = e;
        l += e;
      }
      mbar_wait(p_empty, (j & 1) ^ 1);        // the previous tile's P V MMAs have read the buffer
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float x0 = pv[cc * 8 + 2 * t], x1 = pv[cc * 8 + 2 * t + 1];
          __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
          const float2 hf = __bfloat1622float2(hb);
          __nv_bfloat162 lb = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
          hi[t] = *reinterpret_cast<uint32_t*>(&hb);
          lo[t] = *reinterpret_cast<uint32_t*>(&lb);
        }
        const int phys = (cc ^ (r & 7)) << 4;   // 128B swizzle: 16-byte chunk index XOR row-in-atom
        *reinterpret_cast<uint4*>(prow + phys) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(prow + kPBlk + phys) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qrow = q0 + r;
    const float inv = 1.0f / l;
    uint16_t* orow = p.out + (static_cast<long long>(b) * p.nq + qrow) * p.o_ct + p.o_off + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t a[16];
      tmem_ld16(tlane + 128 + c * 16, a);
      tmem_ld_wait();
      if (qrow < p.nq) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float x0 = __uint_as_float(a[2 * t]) * inv, x1 = __uint_as_float(a[2 * t + 1]) * inv;
          __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
          const float2 hf = __bfloat1622float2(hb);
          __nv_bfloat162 lb = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
          hi[t] = *reinterpret_cast<uint32_t*>(&hb);
          lo[t] = *reinterpret_cast<uint32_t*>(&lb);
        }
        uint4* oh = reinterpret_cast<uint4*>(orow + c * 16);
        uint4* ol = reinterpret_cast<uint4*>(orow + p.o_ps + c * 16);
        oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int HD>
int launch_tc(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const TcParams& p, cudaStream_t st) {
  constexpr int NKB = (HD + 63) / 64;
  const size_t smem = 1024 + 2 * NKB * kQBlk + 2 * (2 * NKB * kKBlk) + 2 * (2 * HD * 128) + 2 * kPBlk + 17 * 8 + 16;
  MTB_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((p.nq + kQTile - 1) / kQTile), static_cast<unsigned>(p.heads), static_cast<unsigned>(p.B));
  attn_tc_kernel<HD><<<grid, kTcThreads, smem, st>>>(tmQ, tmK, tmV, p);
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

long long attention_tc_workspace_bytes(int B, int heads, int hd, int nk) {
  const long long nk_pad = (nk + 63) / 64 * 64;
  return 2ll * B * heads * hd * nk_pad * 2;
}

// 1: not eligible (caller falls back to the CUDA-core kernel), 0: launched, < 0: error
int launch_attention_tc(const mtb_attn_desc* d, cudaStream_t st) {
  if (d->mode != 0 || d->planes != 2 || !(d->hd == 64 || d->hd == 96 || d->hd == 128)) return 1;
  if (d->nq < 128 || d->nk < 256 || d->workspace == nullptr) return 1;
  if (d->workspace_bytes < attention_tc_workspace_bytes(d->B, d->heads, d->hd, d->nk)) return 1;
  if ((d->q_ct | d->k_ct | d->v_ct | d->o_ct | d->o_off) % 8 != 0 || (d->q_ps | d->k_ps | d->o_ps) % 8 != 0) return 1;
  const int nk_pad = (d->nk + 63) / 64 * 64;
  uint16_t* vt = static_cast<uint16_t*>(d->workspace);
  {
    const size_t smem = static_cast<size_t>(d->hd) * 66 * 2;
    dim3 grid(static_cast<unsigned>(nk_pad / 64), static_cast<unsigned>(d->B * d->heads));
    attn_vt_kernel<<<grid, 256, smem, st>>>(static_cast<const uint16_t*>(d->v), d->v_ct, d->v_off, d->v_ps, 2, d->heads,
                                            d->hd, d->nk, nk_pad, vt);
    MTB_CUDA_OK(cudaGetLastError());
  }
  CUtensorMap tmQ, tmK, tmV;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d->q_ct), static_cast<uint64_t>(d->B) * d->nq, 2};
    const uint64_t strides[2] = {static_cast<uint64_t>(d->q_ct) * 2, static_cast<uint64_t>(d->q_ps) * 2};
    const uint32_t box[3] = {64, kQTile, 1};
    if (encode_tmap(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d->q, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d->k_ct), static_cast<uint64_t>(d->B) * d->nk, 2};
    const uint64_t strides[2] = {static_cast<uint64_t>(d->k_ct) * 2, static_cast<uint64_t>(d->k_ps) * 2};
    const uint32_t box[3] = {64, kKTile, 1};
    if (encode_tmap(&tmK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d->k, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  {
    const uint64_t bh = static_cast<uint64_t>(d->B) * d->heads;
    const uint64_t dims[4] = {static_cast<uint64_t>(nk_pad), static_cast<uint64_t>(d->hd), bh, 2};
    const uint64_t strides[3] = {static_cast<uint64_t>(nk_pad) * 2, static_cast<uint64_t>(d->hd) * nk_pad * 2,
                                 bh * d->hd * nk_pad * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(d->hd), 1, 1};
    if (encode_tmap(&tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, vt, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  TcParams p;
  p.B = d->B;
  p.heads = d->heads;
  p.nq = d->nq;
  p.nk = d->nk;
  p.c = d->scale * 1.4426950408889634f;
  p.q_off = d->q_off;
  p.k_off = d->k_off;
  p.out = static_cast<uint16_t*>(d->out);
  p.o_ct = d->o_ct;
  p.o_off = d->o_off;
  p.o_ps = d->o_ps;
  switch (d->hd) {
    case 64: return launch_tc<64>(tmQ, tmK, tmV, p, st);
    case 96: return launch_tc<96>(tmQ, tmK, tmV, p, st);
    default: return launch_tc<128>(tmQ, tmK, tmV, p, st);
  }
}

}  // namespace mtb
