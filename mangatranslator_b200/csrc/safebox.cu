// C-ABI entry points of the safe-text-box path (see safebox_core.cuh for the algorithm).
#include <atomic>
#include <cmath>

#include "../../include/mtb200.h"
#include "common.cuh"
#include "safebox_core.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtbsafe;

static_assert(sizeof(mtb_safebox_job) == sizeof(Job), "mtb_safebox_job / mtbsafe::Job layout mismatch");
static_assert(sizeof(mtb_safebox_result) == sizeof(Result), "mtb_safebox_result / mtbsafe::Result layout mismatch");

namespace {

constexpr int kSafeThreads = 512;
constexpr int kBoundsThreads = 256;
constexpr int kBoundsBlocksPerJob = 64;

// Tight bounds of the nonzero pixels of every job's mask: one warp per row, 16-byte loads when the rows allow it.
// Streams n x H x W bytes once (HBM-bound); results[j].mask_bbox holds running maxima of {-x0, -y0, x1, y1}
// (initialised to a very negative value by the caller's memset).
__global__ void __launch_bounds__(kBoundsThreads) safebox_bounds_kernel(const Job* jobs, Result* results) {
  const int j = blockIdx.y;
  const Job J = jobs[j];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kBoundsThreads / 32;
  const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * warps_per_block;
  constexpr int kNone = -0x7fffffff;
  int nx0 = kNone, ny0 = kNone, x1 = kNone, y1 = kNone;
  const bool vec = ((reinterpret_cast<uintptr_t>(J.mask) | static_cast<uintptr_t>(J.pitch)) & 15) == 0;
  for (int Y = warp; Y < J.H; Y += n_warps) {
    const uint8_t* row = J.mask + static_cast<long long>(Y) * J.pitch;
    bool any = false;
    int X0 = 0;
    if (vec) {
      const int nq = J.W >> 4;
      for (int q = lane; q < nq; q += 32) {
        const uint4 v = reinterpret_cast<const uint4*>(row)[q];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!w[k]) continue;
          const int base = 16 * q + 4 * k;
          const int lo = base + ((__ffs(static_cast<int>(w[k])) - 1) >> 3), hi = base + ((31 - __clz(static_cast<int>(w[k]))) >> 3);
          nx0 = max(nx0, -lo);
          x1 = max(x1, hi);
          any = true;
        }
      }
      X0 = nq << 4;
    }
    for (int X = X0 + lane; X < J.W; X += 32) {
      if (!row[X]) continue;
      nx0 = max(nx0, -X);
      x1 = max(x1, X);
      any = true;
    }
    if (any) {
      ny0 = max(ny0, -Y);
      y1 = max(y1, Y);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    nx0 = max(nx0, __shfl_xor_sync(0xffffffffu, nx0, o));
    ny0 = max(ny0, __shfl_xor_sync(0xffffffffu, ny0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0 && x1 != kNone) {
    atomicMax(&results[j].mask_bbox[0], nx0);
    atomicMax(&results[j].mask_bbox[1], ny0);
    atomicMax(&results[j].mask_bbox[2], x1);
    atomicMax(&results[j].mask_bbox[3], y1);
  }
}

__global__ void __launch_bounds__(kSafeThreads) safebox_kernel(const Job* jobs, Result* results, int n_jobs) {
  __shared__ Shared sh;
  for (int j = blockIdx.x; j < n_jobs; j += gridDim.x) {
    __syncthreads();
    safe_job(jobs[j], results[j], &sh);
    __syncthreads();
  }
}

}  // namespace

extern "C" {

unsigned int mtb_safebox_threshold_sq(double padding_pixels) {
  const float p = static_cast<float>(padding_pixels);   // `dist >= padding` compares in float32 (NumPy 2 weak scalars)
  if (!(p > 0.0f)) return 0u;
  double n = std::ceil(static_cast<double>(p) * static_cast<double>(p));
  if (n > 4.0e9) return 0xffffffffu;
  unsigned int t = static_cast<unsigned int>(n);
  while (t > 0 && std::sqrt(static_cast<float>(t - 1)) >= p) --t;
  while (t < 0xffffffffu && std::sqrt(static_cast<float>(t)) < p) ++t;
  return t;
}

int mtb_safe_boxes(const mtb_safebox_job* jobs_dev, mtb_safebox_result* results_dev, int n_jobs, void* stream) {
  MTB_REQUIRE(jobs_dev && results_dev, "mtb_safe_boxes: null argument");
  if (n_jobs <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // every int of the results becomes 0x80808080 (very negative): the neutral element of the bounds pass' maxima
  MTB_CUDA_OK(cudaMemsetAsync(results_dev, 0x80, sizeof(Result) * static_cast<size_t>(n_jobs), st));
  safebox_bounds_kernel<<<dim3(kBoundsBlocksPerJob, n_jobs), kBoundsThreads, 0, st>>>(
      reinterpret_cast<const Job*>(jobs_dev), reinterpret_cast<Result*>(results_dev));
  MTB_CUDA_OK(cudaGetLastError());
  mtb::g_launches.fetch_add(1);
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_jobs < 4 * sms ? n_jobs : 4 * sms;
  safebox_kernel<<<grid, kSafeThreads, 0, st>>>(reinterpret_cast<const Job*>(jobs_dev),
                                                reinterpret_cast<Result*>(results_dev), n_jobs);
  MTB_CUDA_OK(cudaGetLastError());
  mtb::g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
