// Hardware-behaviour probes (test-only; exported so tests/ can run them on the GPU box).
// mtb_exp_shifted_desc: does a UMMA shared-memory descriptor whose start address is offset by a whole number of
// 128-byte rows inside a 128B-swizzled TMA tile read the rows one expects?  (Needed for halo-tile convolution,
// where the nine 3x3 taps are nine shifted views of ONE shared-memory tile.)
#include "../../include/mtb200.h"
#include "common.cuh"

using namespace mtb;

namespace {

__global__ void __launch_bounds__(128, 1)
shifted_desc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D,
                    int shift_rows, int sbo_bytes, int base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 512 rows x 128 B = 64 KB
  uint8_t* sB = smem + 512 * 128;     // 64 rows x 128 B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 512 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sA + 256 * 128, &tmA, bar, 0, 256);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  if (warp == 1) {
    mbar_wait(bar, 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64);
      const uint32_t a0 = smem_u32(sA) + shift_rows * 128;
      const uint32_t b0 = smem_u32(sB);
      for (int k = 0; k < 4; ++k) {
        umma_bf16(tmem_base, make_sdesc_sw128(a0 + k * 32, sbo_bytes, base_offset),
                  make_sdesc_sw128(b0 + k * 32, 1024, 0), idesc, k > 0);
      }
      umma_commit(mma_bar);
    }
    __syncwarp();
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[r * 64 + c0 + j] = __uint_as_float(acc[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// Shifted-view probe for 8-bit operands with 64-byte rows (the e5m2 activation plane of DESIGN.md §8.0: 64 channels x 1 B
// per pixel): same question as shifted_desc_kernel, for SWIZZLE_64B and kind::f8f6f4.  A = 512 rows x 64 B written by TMA
// with the 64B swizzle; the descriptor starts `shift_rows` rows into the tile and walks 8-row groups `sbo_bytes` apart;
// D[128][64] = A_view * B^T over K = 64 (two K = 32 instructions).
__global__ void __launch_bounds__(128, 1)
shifted_desc8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D,
                     int shift_rows, int sbo_bytes, int base_offset, int kind /* 0: e5m2 K=32, 1: fp16 K=16 */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 512 rows x 64 B = 32 KB
  uint8_t* sB = smem + 512 * 64;      // 64 rows x 64 B = 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 64);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 512 * 64 + 64 * 64);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sA + 256 * 64, &tmA, bar, 0, 256);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  if (warp == 1) {
    mbar_wait(bar, 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t a0 = smem_u32(sA) + shift_rows * 64;
      const uint32_t b0 = smem_u32(sB);
      for (int k = 0; k < 2; ++k) {
        // a 64-byte row is K = 64 e5m2 values or K = 32 halves: either way two instructions, 32 B apart
        if (kind == 0)
          umma_f8f6f4(tmem_base, make_sdesc_sw64(a0 + k * 32, sbo_bytes, base_offset),
                      make_sdesc_sw64(b0 + k * 32, 512, 0), make_idesc_e5m2(128, 64), k > 0);
        else
          umma_bf16(tmem_base, make_sdesc_sw64(a0 + k * 32, sbo_bytes, base_offset),
                    make_sdesc_sw64(b0 + k * 32, 512, 0), make_idesc_f16(128, 64), k > 0);
      }
      umma_commit(mma_bar);
    }
    __syncwarp();
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[r * 64 + c0 + j] = __uint_as_float(acc[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// Mixed-kind accumulation probe: D[128][64] = A16 * B16^T (fp16, K = 64: four kind::f16 MMAs) +
// A8 * B8^T (e5m2, K = 128: four kind::f8f6f4 MMAs of M = m8 rows) accumulated in ONE fp32 TMEM tile.  All operands are
// K-major 128-byte rows written by TMA with the 128B swizzle (a row of 64 halves or 128 bytes).  The host compares D with
// the float64 product; with m8 = 64 it also shows in which TMEM lanes an M = 64 instruction puts its rows.
__global__ void __launch_bounds__(128, 1)
mixed_kind_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmB16,
                  const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8, float* D, int m8,
                  int which) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA16 = smem;                       // 128 rows x 128 B
  uint8_t* sB16 = smem + 16 * 1024;           //  64 rows x 128 B
  uint8_t* sA8 = smem + 24 * 1024;            // 128 rows x 128 B
  uint8_t* sB8 = smem + 40 * 1024;            //  64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (128 + 64 + 128 + 64) * 128);
    tma_load_2d(sA16, &tmA16, bar, 0, 0);
    tma_load_2d(sB16, &tmB16, bar, 0, 0);
    tma_load_2d(sA8, &tmA8, bar, 0, 0);
    tma_load_2d(sB8, &tmB8, bar, 0, 0);
  }
  if (warp == 1) {
    mbar_wait(bar, 0);
    tc_fence_after();
    if (lane == 0) {
      bool acc = false;
      if (which & 1) {
        const uint32_t idesc = make_idesc_f16(128, 64);
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base, make_sdesc_sw128(smem_u32(sA16) + k * 32, 1024, 0),
                    make_sdesc_sw128(smem_u32(sB16) + k * 32, 1024, 0), idesc, acc);
          acc = true;
        }
      }
      if (which & 2) {
        const uint32_t idesc = make_idesc_e5m2(m8, 64);
        for (int k = 0; k < 4; ++k) {
          umma_f8f6f4(tmem_base, make_sdesc_sw128(smem_u32(sA8) + k * 32, 1024, 0),
                      make_sdesc_sw128(smem_u32(sB8) + k * 32, 1024, 0), idesc, acc);
          acc = true;
        }
      }
      umma_commit(mma_bar);
    }
    __syncwarp();
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, acc);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[r * 64 + c0 + j] = __uint_as_float(acc[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// MMA rate probe: one elected lane per CTA issues `iters` rounds of 36 tcgen05.mma (nine row-shifted A views x four
// k16 slices, like one output tile of the halo kernels) on arbitrary shared-memory contents and the CTA reports the
// clock64 span from first issue to commit arrival.  The shape is a template parameter and the issue loop is branch-free
// so that the tensor pipe, not the issuing thread, is what is measured.  M2 > 0 appends a second MMA of shape M2 x N2
// after each one (the mixed shapes of the bf16x3 schemes).
// FP8 = 1: the FIRST MMA is kind::f8f6f4 (e5m2, K = 32); FP8 = 2: the SECOND one is, and it accumulates into the same TMEM
// columns as the first (the fp16 + e5m2-correction scheme of DESIGN.md §8.0: mixed kinds into one fp32 accumulator).
// FP8 = 3 (the slot schedule of the fp16 + e5m2 body conv): two kind::f16 M1 x N1 MMAs, then one e5m2 M2 x N2 MMA, all
// into one accumulator.  SW64: operands described as 64-byte-swizzled 64-byte rows (B groups 512 B apart).
template <int M1, int N1, int M2, int N2, int FP8 = 0, bool SW64 = false>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(long long* cycles, int iters, int a_sbo, int a_shift) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (48 + 144) * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (48 + 144) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    long long t0 = clock64();
    if (elect_one()) {
      constexpr uint32_t id1 = make_idesc_bf16(M1, N1);
      constexpr uint32_t id2 = make_idesc_bf16(M2 > 0 ? M2 : 128, N2 > 0 ? N2 : 64);
      const uint32_t a_base = sbase, b_base = sbase + 48 * 1024;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a0 = a_base + tap * a_shift, b0 = b_base + (tap & 7) * 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = SW64 ? make_sdesc_sw64(a0 + (k & 1) * 32 + (k >> 1) * 8192, a_sbo, 0)
                                     : make_sdesc_sw128(a0 + k * 32, a_sbo, 0);
            const uint64_t db = SW64 ? make_sdesc_sw64(b0 + (k & 1) * 32 + (k >> 1) * 8192, 512, 0)
                                     : make_sdesc_sw128(b0 + k * 32, 1024, 0);
            if (FP8 == 3) {
              umma_bf16(0, da, db, id1, (it | tap | k) ? 1u : 0u);
              umma_bf16(0, da, db, id1, 1u);
              umma_f8f6f4(0, da, db, id2, 1u);
              continue;
            }
            if (FP8 == 1) umma_f8f6f4(0, da, db, id1, (it | tap | k) ? 1u : 0u);
            else umma_bf16(0, da, db, id1, (it | tap | k) ? 1u : 0u);
            if (M2 > 0) {
              if (FP8 == 2) umma_f8f6f4(0, da, db, id2, 1u);
              else umma_bf16(256, da, db, id2, (it | tap | k) ? 1u : 0u);
            }
          }
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(0, 512);
}

// tcgen05.ld fragment-layout probe: every thread writes its own TMEM lane with tcgen05.st.32x32b (value = lane * 1000 +
// column), then each warp reads 16 lanes x 16 columns with tcgen05.ld.16x256b.x2 starting at lane 32*warp + lane_off and
// dumps its 8 registers: out[warp][thread][8].
__global__ void __launch_bounds__(128, 1) ldtm_layout_kernel(float* out, int lane_off) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot;
  uint32_t v[16];
  for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(static_cast<float>((warp * 32 + lane) * 1000 + j));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(base + (static_cast<uint32_t>(warp * 32) << 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
      "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(base + (static_cast<uint32_t>(warp * 32 + lane_off) << 16))
               : "memory");
  tmem_ld_wait();
  for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 8 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 32);
}

}  // namespace

extern "C" int mtb_exp_ldtm_layout(float* out /* [4][32][8] */, int lane_off, void* stream) {
  ldtm_layout_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(out, lane_off);
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

// pattern: 0 M128N64, 1 M128N128, 2 M128N256, 3 M128N128+M128N64, 4 M128N240, 5 M64N240, 6 M64N256, 7 M64N128,
//          8 M128N240+M64N240, 9 M64N64, 10 M128N240 e5m2, 11 M64N240 e5m2, 12 M128N240 f16 + M64N240 e5m2 into one
//          accumulator, 13 M128N240 f16 + M128N240 e5m2 into one accumulator, 14 = 2 x M128N240 f16 + 1 x M64N240 e5m2
//          (one accumulator, 128B swizzle), 15 = 14 with 64-byte-swizzled 64-byte rows, 16 = M128N240 f16 alone, 64B swizzle,
//          17 = M64N240 e5m2 alone, 64B swizzle
extern "C" int mtb_exp_mma_rate(long long* cycles, int ctas, int pattern, int iters, int a_sbo, int a_shift, int a_off,
                                void* stream) {
  (void)a_off;
  const size_t smem = 1024 + (48 + 144) * 1024 + 64;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MTB_RATE9(M1, N1, M2, N2, F8, S64)                                                                        \
  do {                                                                                                            \
    MTB_CUDA_OK(cudaFuncSetAttribute(mma_rate_kernel<M1, N1, M2, N2, F8, S64>,                                    \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));       \
    mma_rate_kernel<M1, N1, M2, N2, F8, S64><<<ctas, 64, smem, st>>>(cycles, iters, a_sbo, a_shift);              \
  } while (0)
#define MTB_RATE8(M1, N1, M2, N2, F8) MTB_RATE9(M1, N1, M2, N2, F8, false)
#define MTB_RATE(M1, N1, M2, N2) MTB_RATE8(M1, N1, M2, N2, 0)
  switch (pattern) {
    case 0: MTB_RATE(128, 64, 0, 0); break;
    case 1: MTB_RATE(128, 128, 0, 0); break;
    case 2: MTB_RATE(128, 256, 0, 0); break;
    case 3: MTB_RATE(128, 128, 128, 64); break;
    case 4: MTB_RATE(128, 240, 0, 0); break;
    case 5: MTB_RATE(64, 240, 0, 0); break;
    case 6: MTB_RATE(64, 256, 0, 0); break;
    case 7: MTB_RATE(64, 128, 0, 0); break;
    case 8: MTB_RATE(128, 240, 64, 240); break;
    case 10: MTB_RATE8(128, 240, 0, 0, 1); break;
    case 11: MTB_RATE8(64, 240, 0, 0, 1); break;
    case 12: MTB_RATE8(128, 240, 64, 240, 2); break;
    case 13: MTB_RATE8(128, 240, 128, 240, 2); break;
    case 14: MTB_RATE9(128, 240, 64, 240, 3, false); break;
    case 15: MTB_RATE9(128, 240, 64, 240, 3, true); break;
    case 16: MTB_RATE9(128, 240, 0, 0, 0, true); break;
    case 17: MTB_RATE9(64, 240, 0, 0, 1, true); break;
    default: MTB_RATE(64, 64, 0, 0); break;
  }
#undef MTB_RATE
#undef MTB_RATE8
#undef MTB_RATE9
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mtb_exp_shifted_desc(const void* A /* bf16 [512][64] */, const void* B /* bf16 [64][64] */,
                                    float* D /* [128][64] */, int shift_rows, int sbo_bytes, int base_offset,
                                    void* stream) {
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {64, 512};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 256};
    if (encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    if (encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  const size_t smem = 1024 + 512 * 128 + 64 * 128 + 64;
  MTB_CUDA_OK(cudaFuncSetAttribute(shifted_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  shifted_desc_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, D, shift_rows, sbo_bytes,
                                                                          base_offset);
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}


// which: bit 0 = the four fp16 MMAs, bit 1 = the four e5m2 MMAs (m8 = 64 or 128 rows).  With bit 0 clear and m8 = 64 the
// accumulator is not initialised in the lanes an M = 64 instruction does not write: zero D first and read where rows land.
extern "C" int mtb_exp_mixed_kind(const void* A16 /* fp16 [128][64] */, const void* B16 /* fp16 [64][64] */,
                                  const void* A8 /* e5m2 [128][128] */, const void* B8 /* e5m2 [64][128] */,
                                  float* D /* [128][64] */, int m8, int which, void* stream) {
  MTB_REQUIRE(m8 == 64 || m8 == 128, "mtb_exp_mixed_kind: m8 must be 64 or 128");
  CUtensorMap t[4];
  const void* ptrs[4] = {A16, B16, A8, B8};
  const uint64_t rows[4] = {128, 64, 128, 64};
  for (int i = 0; i < 4; ++i) {
    const bool half = i < 2;
    const uint64_t dims[2] = {half ? 64u : 128u, rows[i]};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {half ? 64u : 128u, static_cast<uint32_t>(rows[i])};
    if (encode_tmap(&t[i], half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptrs[i], dims,
                    strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B))
      return -3;
  }
  const size_t smem = 1024 + 48 * 1024 + 64;
  MTB_CUDA_OK(cudaFuncSetAttribute(mixed_kind_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  mixed_kind_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(t[0], t[1], t[2], t[3], D, m8, which);
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mtb_exp_shifted_desc8(const void* A /* e5m2 [512][64] or fp16 [512][32] */,
                                     const void* B /* e5m2 [64][64] or fp16 [64][32] */, float* D /* [128][64] */,
                                     int shift_rows, int sbo_bytes, int base_offset, int kind, void* stream) {
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {64, 512};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, 256};
    if (encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B))
      return -3;
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, 64};
    if (encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B))
      return -3;
  }
  const size_t smem = 1024 + 512 * 64 + 64 * 64 + 64;
  MTB_CUDA_OK(cudaFuncSetAttribute(shifted_desc8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  shifted_desc8_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, D, shift_rows, sbo_bytes,
                                                                           base_offset, kind);
  MTB_CUDA_OK(cudaGetLastError());
  return 0;
}
