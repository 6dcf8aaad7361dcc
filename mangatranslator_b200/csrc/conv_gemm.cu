// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.cuh for the design).
#include "conv_gemm.cuh"
#include "epilogue.cuh"

namespace mtb {

namespace {

constexpr int kThreads = kConvThreads;
constexpr int kTileM = 128;         // pixels per tile == TMEM lanes
constexpr int kChunkBytes = 128;    // 64 bf16 channels == one 128B swizzle row
constexpr int kATileBytes = kTileM * kChunkBytes;

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

template <int NSPLIT, int ACT>
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
      mbar_init(&tempty_bar[s], kEpiWarps);
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
    // the whole warp walks the loop; one elected lane issues (see elect_one() in common.cuh)
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
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          uint8_t* sb = sa + PLANES * kATileBytes;
          const int xi = t.x0 * p.stride + kx - p.pad;
          const int yi = t.y0 * p.stride + ky - p.pad;
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            tma_load_4d(sa + pl * kATileBytes, &tmA, &full_bar[stage], p.in_coff + cc * 64, xi, yi, pl * p.N + t.n);
          }
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            tma_load_2d(sb + pl * b_tile_bytes, &tmB, &full_bar[stage], cc * 64, (pl * taps + tap) * p.Cout + n0);
          }
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
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
        if (elect_one()) {
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
      if (elect_one()) umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..17) -------------------------------
    // warp w may touch TMEM lanes [32*(w%4), +32); the four warps of a lane quarter split the channel chunks
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r / p.TW;
    const int tx = r - ty * p.TW;
    int as = 0;
    uint32_t aphase = 0;
    float cta_sums[4] = {0.f, 0.f, 0.f, 0.f};   // per-CTA channel sums of this warp's chunks (sums_per_cta mode)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p);
      const int n0 = t.nt * p.BN;
      const int oy = t.y0 + ty;
      const int ox = t.x0 + tx;
      const bool valid = (oy < p.Ho) && (ox < p.Wo);
      const long long pix = (static_cast<long long>(t.n) * p.Ho + oy) * p.Wo + ox;
      const long long mtile = (static_cast<long long>(t.n) * p.tiles_y + t.tyi) * p.tiles_x + t.txi;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.BN);
      for (int c0 = cg * 16; c0 < p.BN; c0 += 64) {
        uint32_t acc[16];
        tmem_ld16(taddr + c0, acc);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        epilogue_chunk16<ACT>(p, v, valid, pix, n0 + c0, t.n, oy, ox, lane, q, mtile, cta_sums[(c0 >> 6) & 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (p.tile_sums && p.sums_per_cta && lane < 16) {
      for (int c0 = cg * 16, k = 0; c0 < p.BN; c0 += 64, ++k)
        p.tile_sums[(static_cast<long long>(blockIdx.x) * 4 + q) * p.Cout + c0 + lane] = cta_sums[k];
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
#define MTB_LAUNCH_GEMM(NS, ACT)                                                                              \
  do {                                                                                                        \
    MTB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<NS, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     static_cast<int>(smem)));                                                \
    conv_gemm_kernel<NS, ACT><<<grid, kThreads, smem, stream>>>(tmA, tmB, p);                                 \
  } while (0)
  if (nsplit == 3) {
    switch (p.act) {
      case ACT_NONE: MTB_LAUNCH_GEMM(3, ACT_NONE); break;
      case ACT_RELU: MTB_LAUNCH_GEMM(3, ACT_RELU); break;
      case ACT_SILU: MTB_LAUNCH_GEMM(3, ACT_SILU); break;
      default: MTB_LAUNCH_GEMM(3, -1); break;
    }
  } else {
    switch (p.act) {
      case ACT_NONE: MTB_LAUNCH_GEMM(1, ACT_NONE); break;
      case ACT_RELU: MTB_LAUNCH_GEMM(1, ACT_RELU); break;
      case ACT_SILU: MTB_LAUNCH_GEMM(1, ACT_SILU); break;
      default: MTB_LAUNCH_GEMM(1, -1); break;
    }
  }
#undef MTB_LAUNCH_GEMM
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mtb
