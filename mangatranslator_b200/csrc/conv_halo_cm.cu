// 3x3 / stride 1 / pad 1 / 64 -> 64 channel convolution in bf16x3, CHANNEL-MAJOR accumulators — the RCAN body layer
// (reference: spandrel RCAN behind core/image/image_utils.py:369-374; ~400 launches per page).
//
// Why a second halo kernel.  `mtb_exp_mma_rate` (experiments.cu, profiles/r01_mma_rate.json) measured the cost of one
// tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, SW128 smem operands) on B200 as
//        N=64: 98 clk      N=128: 117 clk      N=256: 128 clk (= the 8192 FLOP/clk/SM math rate)
// i.e. an MMA costs ~100 clk of operand fetch whatever its N, so the pixel-major kernel in conv_halo.cu (M = 128
// pixels, N = 128 [W_hi;W_lo] + N = 64 W_hi per tap and k16: 215-244 clk per 128 pixels) sits at its issue bound at
// ~45 % of the tensor pipe.  Here the roles are swapped:
//
//        D[128 x 240]  +=  A[128 x 16] (weights, resident)  x  B[240 x 16] (pixels, tap-shifted halo view)
//
//   * M = 128 TMEM lanes = 64 output channels x {W_hi, W_lo} rows; N = 240 pixels (an 8 x 30 output tile).  One
//     ~126-clk MMA now covers 240 pixels for BOTH weight planes; the activation's lo plane reuses the same A rows
//     (the extra W_lo*X_lo term is ~2^-18 relative and only makes the result more exact).  2 x 126 clk per 240 pixels
//     per (tap, k16) = 1.05 clk/pixel against 1.7-1.9 before.
//   * rows are interleaved in groups of 16 — lanes 32q..32q+15 = W_hi rows of channels 16q..16q+15, lanes
//     32q+16..32q+31 = the W_lo rows of the same channels — so the hi/lo partial products of a channel sit in ONE
//     warp's TMEM lane quarter and are summed with a shuffle; no cross-warp exchange.
//   * the halo tile (10 x 32 pixels x 128 B, 128B-swizzled by TMA) is the B operand: tap (ky,kx) is the descriptor
//     start address + (ky*10+kx)*128 B with SBO = 1280 B, exactly the shifted-view trick of conv_halo.cu.
//   * epilogue: a lane ends up with 4 consecutive pixels x 2 adjacent channels, so global traffic is 4-byte
//     accesses that fill whole 32-byte sectors; per-channel sums for the global average pool are plain per-lane
//     accumulations.
//   * shared memory: 144 KB resident weights + 2 x 40 KB plane slots (hi and lo planes of a tile alternate through
//     the ring, so the next tile's hi plane streams in under the current tile's lo-plane MMAs) = 225 KB.
//     TMEM: 2 accumulator stages x 256 columns = all 512 columns.
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
constexpr int kTW = 8;                              // output tile width (pixels); the height is a template parameter:
                                                    //   30 rows (N = 240) with direct 4-byte global stores (default), or
                                                    //   22 rows (N = 176), which frees 16 KB of shared memory to stage the
                                                    //   output and write it with TMA (experiment, see conv_halo_cm_tile)
constexpr int kHW = kTW + 2;                        // halo tile width
constexpr int kSlots = 2;
constexpr int kStageBytes = 4096;                   // per epilogue group: [hi|lo][16 pixels][128 B], 128B-swizzled
constexpr int kTapBytes = 128 * 128;                // 128 interleaved weight rows x 128 B
constexpr int kWBytes = 9 * kTapBytes;              // 147456
constexpr int kAccCols = 256;                       // TMEM columns per accumulator stage

// ACT: activation (-1 = from ConvParams); HAS_RES: a two-plane residual is added after the activation;
// TH: tile height (30: direct stores, 22: staged TMA stores)
template <int ACT, bool HAS_RES, int TH>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_c64_cm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmO, const ConvParams p) {
  constexpr int kTH = TH;
  constexpr bool kStaged = (TH == 22);
  constexpr int kNPix = kTW * kTH;                  // N of the MMA
  constexpr int kHH = kTH + 2;
  constexpr int kSlotBytes = kHW * kHH * 128;       // a multiple of 1024 for both tile heights
  constexpr int kChunks = kNPix / 16;               // column chunks of 16 pixels (= 2 tile rows)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                         // resident weights (A operand)
  uint8_t* sX = smem + kWBytes;               // ring of halo plane slots (B operand)
  uint8_t* sStage = sX + kSlots * kSlotBytes; // staged output (kStaged only): 4 groups x kStageBytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + (kStaged ? 4 * kStageBytes : 0));
  uint64_t* empty_bar = full_bar + kSlots;
  uint64_t* tfull_bar = empty_bar + kSlots;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
    tmem_alloc(tmem_slot, 2 * kAccCols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if (kStaged) tma_prefetch_desc(&tmO);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 TMEM columns are ours, so the allocation starts at column 0.  Using the literal keeps every operand of
  // the MMA instruction in the uniform datapath (an address read back from shared memory is a per-lane value and costs
  // an ELECT/R2UR round trip per MMA, which made the single issuing thread — not the tensor pipe — the bottleneck).
  if (*tmem_slot != 0) __trap();
  constexpr uint32_t tmem_base = 0;

  const int tiles_per_img = p.tiles_y * p.tiles_x;
  const int total_tiles = p.N * tiles_per_img;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: per tap 128 rows; rows [32g, 32g+16) = W_hi of channels 16g.., rows [32g+16, 32g+32) = W_lo
      mbar_expect_tx(w_bar, static_cast<uint32_t>(kWBytes));
      for (int tap = 0; tap < 9; ++tap)
        for (int g = 0; g < 4; ++g)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d(sW + tap * kTapBytes + (g * 32 + pl * 16) * 128, &tmW, w_bar, 0, (pl * 9 + tap) * 64 + g * 16);
      int slot = 0;
      uint32_t phase = 0;
      long long dbg_tma = 0, dbg_empty = 0, dbg_n = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
        // pull the tile two iterations ahead into L2: with only two plane slots a TMA load has one plane's worth of
        // MMAs (~2.6 us) to land, which HBM latency under store traffic does not always meet
        if (!(HALO_DBG(p) & 8)) {
          const int pt = tile + 2 * gridDim.x;
          if (pt < total_tiles) {
            const int pn = pt / tiles_per_img;
            const int prem = pt - pn * tiles_per_img;
            const int pty = prem / p.tiles_x, ptx = prem - pty * p.tiles_x;
            for (int pl = 0; pl < 2; ++pl) tma_prefetch_l2_4d(&tmX, 0, ptx * kTW - 1, pty * kTH - 1, pl * p.N + pn);
          }
        }
        for (int pl = 0; pl < 2; ++pl) {
          const long long te = (HALO_DBG(p) & 16) ? clock64() : 0;
          mbar_wait(&empty_bar[slot], phase ^ 1);
          const long long ti = (HALO_DBG(p) & 16) ? clock64() : 0;
          if (HALO_DBG(p) & 4) {
            mbar_arrive(&full_bar[slot]);
          } else {
            mbar_expect_tx(&full_bar[slot], kSlotBytes);
            tma_load_4d(sX + slot * kSlotBytes, &tmX, &full_bar[slot], 0, txi * kTW - 1, tyi * kTH - 1, pl * p.N + n);
          }
          if (HALO_DBG(p) & 16) {   // experiment: time from TMA issue to landing, and the wait for a free slot
            mbar_wait(&full_bar[slot], phase);
            dbg_tma += clock64() - ti;
            dbg_empty += ti - te;
            ++dbg_n;
          }
          if (++slot == kSlots) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
      if ((HALO_DBG(p) & 16) && p.dbg_out) {
        p.dbg_out[blockIdx.x * 16 + 0] = dbg_tma;
        p.dbg_out[blockIdx.x * 16 + 1] = dbg_empty;
        p.dbg_out[blockIdx.x * 16 + 2] = dbg_n;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, kNPix);
    mbar_wait(w_bar, 0);
    tc_fence_after();
    // shared-window addresses from 32-bit arithmetic on the (uniform) window offset of the dynamic segment
    const uint32_t sw = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sx0 = sw + kWBytes;
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
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kAccCols);
      for (int pl = 0; pl < 2; ++pl) {
        const long long tf = (HALO_DBG(p) & 32) ? clock64() : 0;
        mbar_wait(&full_bar[slot], phase);
        if (HALO_DBG(p) & 32) dbg_wfull += clock64() - tf;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sx = sx0 + slot * kSlotBytes;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if ((HALO_DBG(p) & 2) && tap > 0) continue;
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t a0 = sw + tap * kTapBytes;
            const uint32_t b0 = sx + (ky * kHW + kx) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_sdesc_sw128(a0 + k * 32, 1024, 0);
              const uint64_t db = make_sdesc_sw128(b0 + k * 32, kHW * 128, 0);
              umma_bf16(d_tmem, da, db, idesc, (pl > 0 || tap > 0 || k > 0) ? 1u : 0u);
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
    // 16 epilogue warps.  q = TMEM lane quarter = channels 16q..16q+15 (hi rows in lanes 0-15, lo rows in 16-31);
    // wj = which column chunks (ci = wj, wj+4, ...) of the 15 this warp drains.
    const int q = warp & 3;
    const int wj = (warp - 2) >> 2;
    const bool upper = lane >= 16;
    const bool odd = (lane & 1) != 0;
    const int cpair = 16 * q + (lane & 14);          // the two adjacent channels this lane stores
    const float b0 = p.bias ? __ldg(p.bias + cpair) : 0.0f;
    const float b1 = p.bias ? __ldg(p.bias + cpair + 1) : 0.0f;
    // the channel scale may be written by the kernel just before this one: plain loads, not the read-only path
    const float s0 = p.chan_scale ? p.chan_scale[cpair] : 1.0f;
    const float s1 = p.chan_scale ? p.chan_scale[cpair + 1] : 1.0f;
    const bool want_sums = p.tile_sums != nullptr;
    const long long row2 = 2ll * p.Wo * 64;          // elements between the tile rows of consecutive chunks
    float sum0 = 0.0f, sum1 = 0.0f;
    const bool want_border = want_sums && p.sums_per_cta && p.border_sums != nullptr;
    float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // (top, bottom, left, right) x 2 channels
    int as = 0;
    uint32_t aphase = 0;
    long long dbg_wtfull = 0, dbg_work = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
      // this lane's pixels of chunk ci: tile row 2*ci + upper, columns ox0 .. ox0+3, channels cpair, cpair+1
      const int oy0 = tyi * kTH + (upper ? 1 : 0);
      const int ox0 = txi * kTW + (odd ? 4 : 0);
      const int nvalid = (HALO_DBG(p) & 1) ? 0 : min(4, p.Wo - ox0);            // valid columns (<= 0: none)
      const long long off0 = ((static_cast<long long>(n) * p.Ho + oy0) * p.Wo + ox0) * 64 + cpair;
      // residual words of the first chunk are requested before the accumulator is waited for
      uint32_t rh[4] = {0u, 0u, 0u, 0u}, rl[4] = {0u, 0u, 0u, 0u};
      if (HAS_RES) {
        const int ci = wj;
        if (oy0 + 2 * ci < p.Ho) {
          const uint16_t* rp = p.residual + off0 + ci * row2;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nvalid) {
              rh[j] = *reinterpret_cast<const uint32_t*>(rp + j * 64);
              rl[j] = *reinterpret_cast<const uint32_t*>(rp + p.res_plane_stride + j * 64);
            }
        }
      }
      const long long tw = (HALO_DBG(p) & 32) ? clock64() : 0;
      mbar_wait(&tfull_bar[as], aphase);
      const long long tw1 = (HALO_DBG(p) & 32) ? clock64() : 0;
      dbg_wtfull += tw1 - tw;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * kAccCols);
      float t0 = 0.0f, t1 = 0.0f;
#pragma unroll 1
      for (int ci = wj; ci < kChunks; ci += 4) {
        uint32_t acc[16];
        tmem_ld16(taddr + ci * 16, acc);
        // next chunk's residual words go out while the TMEM read is in flight
        uint32_t nh[4] = {0u, 0u, 0u, 0u}, nl[4] = {0u, 0u, 0u, 0u};
        if (HAS_RES) {
          const int cn = ci + 4;
          if (cn < kChunks && oy0 + 2 * cn < p.Ho) {
            const uint16_t* rp = p.residual + off0 + cn * row2;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < nvalid) {
                nh[j] = *reinterpret_cast<const uint32_t*>(rp + j * 64);
                nl[j] = *reinterpret_cast<const uint32_t*>(rp + p.res_plane_stride + j * 64);
              }
          }
        }
        tmem_ld_wait();
        if (ci + 4 >= kChunks) {
          // last TMEM read of this warp for the tile: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        // hi + lo rows of a channel: lanes < 16 keep the chunk's first tile row (8 pixels), lanes >= 16 the second
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float send = __uint_as_float(upper ? acc[j] : acc[j + 8]);
          const float mine = __uint_as_float(upper ? acc[j + 8] : acc[j]);
          v[j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        // adjacent channels: even lanes keep pixels 0-3 of (c, c+1), odd lanes pixels 4-7 of (c-1, c)
        float x0[4], x1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float send = odd ? v[j] : v[j + 4];
          const float mine = odd ? v[j + 4] : v[j];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
          x0[j] = odd ? recv : mine;
          x1[j] = odd ? mine : recv;
        }
        const int oy = oy0 + 2 * ci;
        const int nv = (oy < p.Ho) ? nvalid : 0;
        uint16_t* op = p.out + off0 + ci * row2;
        float c0 = 0.f, c1 = 0.f;                    // this chunk's contribution to the channel sums
        uint32_t hiw[4], low[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = epi_act<ACT>((x0[j] + b0) * s0, p.act);
          float b = epi_act<ACT>((x1[j] + b1) * s1, p.act);
          if (HAS_RES) {
            const float2 r0 = unpack_bf16x2(rh[j]), r1 = unpack_bf16x2(rl[j]);
            a += r0.x + r1.x;
            b += r0.y + r1.y;
          }
          if (j < nv) {
            c0 += a;
            c1 += b;
            if (want_border) {
              const int ox = ox0 + j;
              if (ox == 0) {
                bsum[4] += a;
                bsum[5] += b;
              }
              if (ox == p.Wo - 1) {
                bsum[6] += a;
                bsum[7] += b;
              }
            }
          }
          {
            const uint32_t hi = pack_bf16x2(a, b);
            const float2 h = unpack_bf16x2(hi);
            hiw[j] = hi;
            low[j] = pack_bf16x2(a - h.x, b - h.y);
          }
          if (!kStaged && j < nv) {
            *reinterpret_cast<uint32_t*>(op + j * 64) = hiw[j];
            *reinterpret_cast<uint32_t*>(op + p.out_plane_stride + j * 64) = low[j];
          }
        }
        if (kStaged) {
          // stage the chunk (16 pixels x 64 channels x hi/lo) in shared memory, 128B-swizzled, and let TMA write full
          // lines; the four warps of this group (one per TMEM lane quarter) each contribute 32 B of every pixel row
          uint8_t* stg = sStage + wj * kStageBytes;
          if (q == 0 && lane == 0) bulk_wait_read0();          // the previous chunk's store has read the buffer
          named_bar_sync(1 + wj, 128);
          const int cp = (lane & 15) >> 1;                     // channel pair inside the quarter
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // upper lanes walk their pixels rotated by two so the four pixel rows of one instruction differ in more
            // than address bit 7: with the XOR swizzle that makes the 32 lanes hit 32 distinct banks
            const int jj = upper ? (k ^ 2) : k;
            const uint32_t wh = upper ? hiw[k ^ 2] : hiw[k];
            const uint32_t wl = upper ? low[k ^ 2] : low[k];
            const int px = (upper ? 8 : 0) + (odd ? 4 : 0) + jj;
            const int boff = px * 128 + ((((q << 1) | (cp >> 2)) ^ (px & 7)) << 4) + ((cp & 3) << 2);
            *reinterpret_cast<uint32_t*>(stg + boff) = wh;
            *reinterpret_cast<uint32_t*>(stg + 2048 + boff) = wl;
          }
          fence_proxy_async();
          named_bar_sync(1 + wj, 128);
          if (q == 0 && lane == 0 && !(HALO_DBG(p) & 1)) {
            tma_store_4d(&tmO, stg, 0, txi * kTW, tyi * kTH + 2 * ci, n);
            tma_store_4d(&tmO, stg + 2048, 0, txi * kTW, tyi * kTH + 2 * ci, p.N + n);
            bulk_commit_group();
          }
        }
        t0 += c0;
        t1 += c1;
        if (want_border) {
          if (oy == 0) {
            bsum[0] += c0;
            bsum[1] += c1;
          }
          if (oy == p.Ho - 1) {
            bsum[2] += c0;
            bsum[3] += c1;
          }
        }
        if (HAS_RES) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rh[j] = nh[j];
            rl[j] = nl[j];
          }
        }
      }
      if (want_sums) {
        if (p.sums_per_cta) {
          sum0 += t0;
          sum1 += t1;
        } else {
          // the four lanes {l, l^1, l^16, l^17} hold the same channel pair (different pixels)
          t0 += __shfl_xor_sync(0xffffffffu, t0, 1);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
          t0 += __shfl_xor_sync(0xffffffffu, t0, 16);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
          if (!upper && !odd) {
            float* row = p.tile_sums + (static_cast<long long>(tile) * 4 + wj) * 64 + cpair;
            row[0] = t0;
            row[1] = t1;
          }
        }
      }
      if (HALO_DBG(p) & 32) dbg_work += clock64() - tw1;
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (kStaged && q == 0 && lane == 0) bulk_wait_all0();      // all staged stores have landed
    if ((HALO_DBG(p) & 32) && p.dbg_out && lane == 0 && (warp == 2 || warp == 17)) {
      p.dbg_out[blockIdx.x * 16 + (warp == 2 ? 8 : 10)] = dbg_wtfull;
      p.dbg_out[blockIdx.x * 16 + (warp == 2 ? 9 : 11)] = dbg_work;
    }
    if (want_border) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        bsum[k] += __shfl_xor_sync(0xffffffffu, bsum[k], 1);
        bsum[k] += __shfl_xor_sync(0xffffffffu, bsum[k], 16);
      }
      if (!upper && !odd) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* row = p.border_sums + ((static_cast<long long>(blockIdx.x) * 4 + wj) * 4 + k) * 64 + cpair;
          row[0] = bsum[2 * k];
          row[1] = bsum[2 * k + 1];
        }
      }
    }
    if (want_sums && p.sums_per_cta) {
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 16);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 16);
      if (!upper && !odd) {
        float* row = p.tile_sums + (static_cast<long long>(blockIdx.x) * 4 + wj) * 64 + cpair;
        row[0] = sum0;
        row[1] = sum1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kAccCols);
}

}  // namespace

// what the channel-major kernel handles beyond conv_halo_eligible(): two output planes, a residual (if any) with two
// planes, activation before the residual
bool conv_halo_cm_eligible(const ConvParams& p) {
  return p.planes_out == 2 && p.out_f32 == nullptr && (p.residual == nullptr || p.res_planes == 2) && !p.act_after_res;
}

// tile geometry of the two variants: direct stores (8 x 30, the default) or staged TMA stores (8 x 22,
// MTB200_CM_STAGED_STORES=1).  Measured on B200 (profiles/r01_cm_staged_stores.json): staging removes the store
// wavefronts but the smaller tile pays more per-tile pipeline bubbles and two named barriers per chunk — 436K vs 413K
// cycles per CTA, 162 vs 156 ms per RCAN page — so it stays an experiment.
void conv_halo_cm_tile(int* tw, int* th) {
  const char* e = getenv("MTB200_CM_STAGED_STORES");
  *tw = kTW;
  *th = (e && atoi(e) != 0) ? 22 : 30;
}

template <int TH>
static int launch_cm_th(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, const ConvParams& p,
                        cudaStream_t stream) {
  constexpr int slot = kHW * (TH + 2) * 128;
  const size_t smem = 1024 + kWBytes + kSlots * slot + (TH == 22 ? 4 * kStageBytes : 0) + 16 * 8 + 16;
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = static_cast<long long>(p.N) * p.tiles_y * p.tiles_x;
  const int grid = static_cast<int>(total < sms ? total : sms);
  if (grid <= 0) return 0;
#define MTB_LAUNCH_CM(ACT, RES)                                                                                       \
  do {                                                                                                                \
    MTB_CUDA_OK(cudaFuncSetAttribute(conv3x3_c64_cm_kernel<ACT, RES, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     static_cast<int>(smem)));                                                        \
    conv3x3_c64_cm_kernel<ACT, RES, TH><<<grid, kThreads, smem, stream>>>(tmX, tmW, tmO, p);                          \
  } while (0)
  const bool res = p.residual != nullptr;
  switch (p.act) {
    case ACT_NONE: if (res) MTB_LAUNCH_CM(ACT_NONE, true); else MTB_LAUNCH_CM(ACT_NONE, false); break;
    case ACT_RELU: if (res) MTB_LAUNCH_CM(ACT_RELU, true); else MTB_LAUNCH_CM(ACT_RELU, false); break;
    default: if (res) MTB_LAUNCH_CM(-1, true); else MTB_LAUNCH_CM(-1, false); break;
  }
#undef MTB_LAUNCH_CM
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv_halo_cm(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, const ConvParams& p_in,
                        cudaStream_t stream) {
  ConvParams p = p_in;
  if (const char* e = getenv("MTB200_HALO_DEBUG")) p.debug = atoi(e);
  if (const char* e = getenv("MTB200_HALO_DEBUG_PTR")) p.dbg_out = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  return p.TH == 22 ? launch_cm_th<22>(tmX, tmW, tmO, p, stream) : launch_cm_th<30>(tmX, tmW, tmO, p, stream);
}

}  // namespace mtb
