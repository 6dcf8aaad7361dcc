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

}  // namespace

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
