// 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels — the RCAN residual-channel-attention body
// layer that runs ~400 times per page (reference: spandrel RCAN behind core/image/image_utils.py:369-374) — as a
// HALO-TILE implicit GEMM on tcgen05:
//
//   * the 9 filter taps are 9 SHIFTED VIEWS of one shared-memory tile: TMA loads the (16+2) x (8+2) pixel halo of a
//     16x8 output tile once (128B-swizzled, one 128-byte row per pixel) and the UMMA shared-memory descriptor of tap
//     (ky,kx) just starts (ky*10+kx) rows later with a 10-row stride between 8-pixel groups (SBO = 1280 B).
//     (Verified on B200 by mtb_exp_shifted_desc: the 128B swizzle is a function of absolute smem address bits, so
//     row-shifted descriptors with base_offset = 0 read the rows TMA wrote.)  L2->SM traffic per tile drops from
//     9 x 32 KB (one TMA box per tap) to 46 KB.
//   * all weights (9 taps x [hi|lo] x 64x64 bf16 = 144 KB) are loaded once per persistent CTA and stay resident.
//   * bf16x3 with ONE N=128 MMA per (tap, k16): A_hi x [W_hi ; W_lo] -> columns 0..63 = Ahi*Whi, 64..127 = Ahi*Wlo,
//     then A_lo x W_hi accumulates into columns 0..63; the epilogue adds the two halves.  Shared-memory operand
//     traffic per MMA is then 128 B/clk for the N=128 instruction, i.e. matched to the SMEM bandwidth.
//   * ring of 3 shared-memory slots, one (tile, plane) halo per slot, so the next tile's hi plane streams in while the
//     current tile's lo plane is being multiplied; 2 TMEM accumulator stages overlap the epilogue with the next tile.
#include <stdlib.h>

#include "conv_gemm.cuh"
#include "epilogue.cuh"

// The clock64() / skip-stage instrumentation of this kernel (ConvParams::debug, MTB200_HALO_DEBUG) is compiled OUT of the
// product library: build with `make EXTRA=-DMTB_HALO_DEBUG=1` to get it back for a perf experiment.
#ifndef MTB_HALO_DEBUG
#define MTB_HALO_DEBUG 0
#endif
#define HALO_DBG(p) (MTB_HALO_DEBUG ? (p).debug : 0)

namespace mtb {

namespace {

constexpr int kThreads = kConvThreads;
constexpr int kTW = 8, kTH = 16;                    // output tile (pixels)
constexpr int kHW = kTW + 2, kHH = kTH + 2;         // halo tile
constexpr int kHaloBytes = kHW * kHH * 128;         // 23040
constexpr int kSlotBytes = 23552;                   // 1024-aligned slot
constexpr int kSlots = 3;
constexpr int kTapBytes = 128 * 128;                // [W_hi(64 rows) ; W_lo(64 rows)] x 128 B
constexpr int kWBytes = 9 * kTapBytes;              // 147456

// PLANES == 2: bf16x3 (fp32-grade);  PLANES == 1: plain bf16
template <int PLANES, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_c64_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                         // resident weights
  uint8_t* sA = smem + kWBytes;               // ring of halo slots
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + kSlots * kSlotBytes);
  uint64_t* empty_bar = full_bar + kSlots;
  uint64_t* tfull_bar = empty_bar + kSlots;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int ACC_COLS = (PLANES == 2) ? 128 : 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * ACC_COLS);
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

  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total_tiles = p.N * tiles_per_img;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: per tap, W_hi rows then W_lo rows
      mbar_expect_tx(w_bar, static_cast<uint32_t>(9 * PLANES * 64 * 128));
      for (int tap = 0; tap < 9; ++tap)
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_2d(sW + tap * kTapBytes + pl * 64 * 128, &tmB, w_bar, 0, (pl * 9 + tap) * 64);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
        // pull the halos of the tile three iterations ahead into L2 while this one streams into shared memory: the
        // 3-slot ring only hides ~1.5 tiles of latency, so the eventual TMA load should be an L2 hit
        {
          const int pt = tile + 3 * gridDim.x;
          if (pt < total_tiles) {
            const int pn = pt / tiles_per_img;
            const int prem = pt - pn * tiles_per_img;
            const int pty = prem / p.tiles_x, ptx = prem - pty * p.tiles_x;
            for (int pl = 0; pl < PLANES; ++pl)
              tma_prefetch_l2_4d(&tmA, 0, ptx * kTW - 1, pty * kTH - 1, pl * p.N + pn);
          }
        }
        for (int pl = 0; pl < PLANES; ++pl) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          if (HALO_DBG(p) & 4) {
            mbar_arrive(&full_bar[slot]);
          } else {
            mbar_expect_tx(&full_bar[slot], kHaloBytes);
            tma_load_4d(sA + slot * kSlotBytes, &tmA, &full_bar[slot], 0, txi * kTW - 1, tyi * kTH - 1, pl * p.N + n);
          }
          if (++slot == kSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_wide = make_idesc_bf16(128, ACC_COLS);  // A_hi x [W_hi;W_lo]  (or A x W for bf16)
    const uint32_t idesc_64 = make_idesc_bf16(128, 64);          // A_lo x W_hi
    mbar_wait(w_bar, 0);
    tc_fence_after();
    const uint32_t sw = smem_u32(sW);
    int slot = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    long long dbg_wfull = 0, dbg_wtempty = 0, dbg_tiles = 0;
    const long long dbg_t0 = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const long long ta = (HALO_DBG(p) & 32) ? clock64() : 0;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      if (HALO_DBG(p) & 32) {
        dbg_wtempty += clock64() - ta;
        ++dbg_tiles;
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * ACC_COLS);
      for (int pl = 0; pl < PLANES; ++pl) {
        const long long tf = (HALO_DBG(p) & 32) ? clock64() : 0;
        mbar_wait(&full_bar[slot], phase);
        if (HALO_DBG(p) & 32) dbg_wfull += clock64() - tf;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(sA + slot * kSlotBytes);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if ((HALO_DBG(p) & 2) && tap > 0) continue;
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t a0 = sa + (ky * kHW + kx) * 128;
            const uint32_t b0 = sw + tap * kTapBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_sdesc_sw128(a0 + k * 32, kHW * 128, 0);
              const uint64_t db = make_sdesc_sw128(b0 + k * 32, 1024, 0);
              if (pl == 0) umma_bf16(d_tmem, da, db, idesc_wide, (tap > 0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, da, db, idesc_64, 1u);
            }
          }
          umma_commit(&empty_bar[slot]);
        }
        __syncwarp();
        if (++slot == kSlots) {
          slot = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull_bar[as]);
      __syncwarp();
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if ((HALO_DBG(p) & 32) && p.dbg_out && lane == 0) {
      p.dbg_out[blockIdx.x * 16 + 4] = dbg_wfull;
      p.dbg_out[blockIdx.x * 16 + 5] = dbg_wtempty;
      p.dbg_out[blockIdx.x * 16 + 6] = dbg_tiles;
      p.dbg_out[blockIdx.x * 16 + 7] = clock64() - dbg_t0;
    }
  } else {
    // 16 epilogue warps: lane quarter q = warp % 4, channel chunk cg = (warp - 2) / 4 (16 channels each)
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    const int c0 = cg * 16;
    int as = 0;
    uint32_t aphase = 0;
    float cta_sum = 0.f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      const int oy = tyi * kTH + ty, ox = txi * kTW + tx;
      const bool valid = (oy < p.Ho) && (ox < p.Wo);
      const long long pix = (static_cast<long long>(n) * p.Ho + oy) * p.Wo + ox;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * ACC_COLS);
      uint32_t acc[16];
      float v[16];
      tmem_ld16(taddr + c0, acc);
      if (PLANES == 2) {
        uint32_t acc2[16];
        tmem_ld16(taddr + 64 + c0, acc2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) + __uint_as_float(acc2[j]);
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
      }
      // the accumulator values are in registers: release the TMEM stage before the (long) store path
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      epilogue_chunk16<ACT>(p, v, valid && !(HALO_DBG(p) & 1), pix, c0, n, oy, ox, lane, q, tile, cta_sum);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (p.tile_sums && p.sums_per_cta && lane < 16)
      p.tile_sums[(static_cast<long long>(blockIdx.x) * 4 + q) * 64 + c0 + lane] = cta_sum;
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

}  // namespace

bool conv_halo_eligible(const ConvParams& p, int cin) {
  return p.KH == 3 && p.KW == 3 && p.stride == 1 && p.pad == 1 && cin == 64 && p.Cout == 64 && p.out != nullptr &&
         p.Ho > 1 && p.in_coff == 0 && p.out_coff == 0 && p.out_cstride == 64 &&
         (p.residual == nullptr || (p.res_coff == 0 && p.res_cstride == 64 && !p.res_bcast)) && !p.pixel_shuffle;
}

int launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p_in, int nsplit,
                     cudaStream_t stream) {
  ConvParams p = p_in;
  if (const char* e = getenv("MTB200_HALO_DEBUG")) p.debug = atoi(e);
  if (const char* e = getenv("MTB200_HALO_DEBUG_PTR")) p.dbg_out = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  const size_t smem = 1024 + kWBytes + kSlots * kSlotBytes + 16 * 8 + 16;
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = static_cast<long long>(p.N) * p.tiles_y * p.tiles_x;
  const int grid = static_cast<int>(total < sms ? total : sms);
  if (grid <= 0) return 0;
#define MTB_LAUNCH_HALO(PL, ACT)                                                                                    \
  do {                                                                                                              \
    MTB_CUDA_OK(cudaFuncSetAttribute(conv3x3_c64_halo_kernel<PL, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     static_cast<int>(smem)));                                                      \
    conv3x3_c64_halo_kernel<PL, ACT><<<grid, kThreads, smem, stream>>>(tmA, tmB, p);                                \
  } while (0)
  if (nsplit == 3) {
    switch (p.act) {
      case ACT_NONE: MTB_LAUNCH_HALO(2, ACT_NONE); break;
      case ACT_RELU: MTB_LAUNCH_HALO(2, ACT_RELU); break;
      default: MTB_LAUNCH_HALO(2, -1); break;
    }
  } else {
    switch (p.act) {
      case ACT_NONE: MTB_LAUNCH_HALO(1, ACT_NONE); break;
      case ACT_RELU: MTB_LAUNCH_HALO(1, ACT_RELU); break;
      default: MTB_LAUNCH_HALO(1, -1); break;
    }
  }
#undef MTB_LAUNCH_HALO
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mtb
