// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.cuh for the design).
#include "conv_gemm.cuh"

namespace mtb {

namespace {

constexpr int kThreads = 192;
constexpr int kTileM = 128;         // pixels per tile == TMEM lanes
constexpr int kChunkBytes = 128;    // 64 bf16 channels == one 128B swizzle row
constexpr int kATileBytes = kTileM * kChunkBytes;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_SILU: return v / (1.0f + expf(-v));
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

struct TileCoord {
  int n, y0, x0, tyi, txi, nt;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, const ConvParams& p) {
  TileCoord t;
  t.nt = tile % p.n_tiles_n;
  int m = tile / p.n_tiles_n;
  t.txi = m % p.tiles_x;
  m /= p.tiles_x;
  t.tyi = m % p.tiles_y;
  t.n = m / p.tiles_y;
  t.x0 = t.txi * p.TW;
  t.y0 = t.tyi * p.TH;
  return t;
}

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvParams p) {
  constexpr int PLANES = (NSPLIT == 3) ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int b_tile_bytes = p.BN * kChunkBytes;
  const int stage_bytes = PLANES * (kATileBytes + b_tile_bytes);
  const int S = p.num_stages;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(S) * stage_bytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.KH * p.KW;
  const int KS = taps * p.cin_chunks;
  const int total_tiles = p.N * p.tiles_y * p.tiles_x * p.n_tiles_n;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, p);
        const int n0 = t.nt * p.BN;
        for (int ks = 0; ks < KS; ++ks) {
          const int tap = ks / p.cin_chunks;
          const int cc = ks - tap * p.cin_chunks;
          const int ky = tap / p.KW;
          const int kx = tap - ky * p.KW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          uint8_t* sb = sa + PLANES * kATileBytes;
          const int xi = t.x0 * p.stride + kx - p.pad;
          const int yi = t.y0 * p.stride + ky - p.pad;
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            tma_load_4d(sa + pl * kATileBytes, &tmA, &full_bar[stage], cc * 64, xi, yi, pl * p.N + t.n);
          }
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            tma_load_2d(sb + pl * b_tile_bytes, &tmB, &full_bar[stage], cc * 64, (pl * taps + tap) * p.Cout + n0);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = make_idesc_bf16(kTileM, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.BN);
      for (int ks = 0; ks < KS; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + PLANES * kATileBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_hi = make_sdesc_sw128(sa + k * 32, 1024, 0);
            const uint64_t b_hi = make_sdesc_sw128(sb + k * 32, 1024, 0);
            umma_bf16(d_tmem, a_hi, b_hi, idesc, (ks > 0 || k > 0) ? 1u : 0u);
            if (NSPLIT == 3) {
              const uint64_t a_lo = make_sdesc_sw128(sa + kATileBytes + k * 32, 1024, 0);
              const uint64_t b_lo = make_sdesc_sw128(sb + b_tile_bytes + k * 32, 1024, 0);
              umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (lane == 0) umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..5) -------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int ty = r / p.TW;
    const int tx = r - ty * p.TW;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p);
      const int n0 = t.nt * p.BN;
      const int oy = t.y0 + ty;
      const int ox = t.x0 + tx;
      const bool valid = (oy < p.Ho) && (ox < p.Wo);
      const long long pix = (static_cast<long long>(t.n) * p.Ho + oy) * p.Wo + ox;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + c0, acc);
        tmem_ld_wait();
        float v[16];
        const int cbase = n0 + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(acc[j]);
          if (p.bias) x += __ldg(p.bias + cbase + j);
          v[j] = apply_act(x, p.act);
        }
        if (p.residual && valid) {
          const uint16_t* rp = p.residual + pix * p.Cout + cbase;
          for (int pl = 0; pl < p.res_planes; ++pl) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rp + pl * p.res_plane_stride);
            uint4 a = __ldg(r4), b = __ldg(r4 + 1);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[2 * j] += bf16_to_f(static_cast<uint16_t>(w[j] & 0xFFFF));
              v[2 * j + 1] += bf16_to_f(static_cast<uint16_t>(w[j] >> 16));
            }
          }
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0f;
        }
        if (p.tile_sums) {
          // channel sums over the warp's 32 pixels: 16 values x 32 lanes -> lanes 0..15 hold channel sums
          float s[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] = v[j];
          // fold lanes 16..31 onto 0..15
#pragma unroll
          for (int j = 0; j < 16; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
          // transpose-reduce across the remaining 16 lanes: 8+4+2+1 exchanges
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
          // lane l (< 16) now holds the sum for channel index bitrev-free mapping: channel = l's bits select halves
          // channel owned by lane l: at step w the lane kept the upper half iff (l & w) -> channel = l & 15
          if (lane < 16) {
            const long long mt = (static_cast<long long>(t.n) * p.tiles_y + t.tyi) * p.tiles_x + t.txi;
            p.tile_sums[(mt * 4 + q) * p.Cout + cbase + lane] = s[0];
          }
        }
        if (valid) {
          if (p.out_f32) {
            float4* o4 = reinterpret_cast<float4*>(p.out_f32 + pix * p.Cout + cbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint16_t h0, l0, h1, l1;
              split_bf16(v[2 * j], h0, l0);
              split_bf16(v[2 * j + 1], h1, l1);
              hi[j] = static_cast<uint32_t>(h0) | (static_cast<uint32_t>(h1) << 16);
              lo[j] = static_cast<uint32_t>(l0) | (static_cast<uint32_t>(l1) << 16);
            }
            long long off = pix * p.Cout + cbase;
            if (p.pixel_shuffle) {
              // PixelShuffle(2) fused into the store: channel block b = dy*2+dx lands on the 2x finer grid
              const int cq = p.Cout >> 2;
              const int blk = cbase / cq, cc = cbase - blk * cq;
              const long long hp = (static_cast<long long>(t.n) * (2 * p.Ho) + 2 * oy + (blk >> 1)) * (2 * p.Wo) +
                                   2 * ox + (blk & 1);
              off = hp * cq + cc;
            }
            uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
            o4[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            o4[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            if (p.planes_out == 2) {
              uint4* l4 = reinterpret_cast<uint4*>(p.out + p.out_plane_stride + off);
              l4[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              l4[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

}  // namespace

size_t conv_gemm_smem_bytes(const ConvParams& p, int nsplit) {
  const int planes = (nsplit == 3) ? 2 : 1;
  const size_t stage = static_cast<size_t>(planes) * (kATileBytes + p.BN * kChunkBytes);
  return 1024 + stage * p.num_stages + (2 * p.num_stages + 4) * 8 + 16;
}

int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int nsplit,
                     cudaStream_t stream) {
  MTB_REQUIRE(nsplit == 1 || nsplit == 3, "conv_gemm: nsplit must be 1 or 3 (got %d)", nsplit);
  MTB_REQUIRE(p.TW * p.TH == kTileM, "conv_gemm: tile must hold 128 pixels (TW=%d TH=%d)", p.TW, p.TH);
  MTB_REQUIRE(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256, "conv_gemm: bad BN %d", p.BN);
  MTB_REQUIRE(p.num_stages >= 2, "conv_gemm: need >= 2 stages");
  const size_t smem = conv_gemm_smem_bytes(p, nsplit);
  MTB_REQUIRE(smem <= 227 * 1024, "conv_gemm: smem %zu too large", smem);
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = static_cast<long long>(p.N) * p.tiles_y * p.tiles_x * p.n_tiles_n;
  const int grid = static_cast<int>(total < sms ? total : sms);
  if (grid <= 0) return 0;
  if (nsplit == 3) {
    MTB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    conv_gemm_kernel<3><<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
  } else {
    MTB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    conv_gemm_kernel<1><<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
  }
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mtb
