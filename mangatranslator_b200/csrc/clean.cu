// C-ABI entry points of the bubble-cleaning path (see clean_core.cuh for the algorithm).
#include <atomic>

#include "../../include/mtb200.h"
#include "clean_core.cuh"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtbclean;

static_assert(sizeof(mtb_clean_params) == sizeof(Params), "mtb_clean_params / mtbclean::Params layout mismatch");
static_assert(sizeof(mtb_clean_job) == sizeof(Job), "mtb_clean_job / mtbclean::Job layout mismatch");
static_assert(sizeof(mtb_clean_result) == sizeof(Result), "mtb_clean_result / mtbclean::Result layout mismatch");
static_assert(MTB_CLEAN_MAX_SE == kMaxSE && MTB_CLEAN_MAX_BALL == kMaxBall && MTB_CLEAN_MAX_NEIGHBORS == kMaxNeighbors,
              "constant mismatch");

namespace {

constexpr int kCleanThreads = 512;
__constant__ Params c_params;

__global__ void __launch_bounds__(kCleanThreads) clean_kernel(const Job* jobs, Result* results, int n_jobs) {
  __shared__ Shared sh;
  for (int j = blockIdx.x; j < n_jobs; j += gridDim.x) {
    __syncthreads();
    clean_job(c_params, jobs[j], results[j], &sh);
    __syncthreads();
  }
}

// rank of each job's fill colour in first-seen order within its page (dict insertion order, cleaning.py:1022-1029)
__global__ void rank_kernel(const Job* jobs, const Result* results, int n_jobs, int* rank) {
  // one thread per job; O(n^2) over the jobs of the same page (n <= a few hundred)
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  if (results[j].status != ST_OK) {
    rank[j] = -1;
    return;
  }
  const int page = jobs[j].page_index;
  int r = 0;
  // distinct colours first seen before this job's colour
  for (int a = 0; a < n_jobs; ++a) {
    if (jobs[a].page_index != page || results[a].status != ST_OK) continue;
    const int* ca = results[a].fill_bgr;
    const int* cj = results[j].fill_bgr;
    if (ca[0] == cj[0] && ca[1] == cj[1] && ca[2] == cj[2]) break;  // first occurrence of my colour reached
    // is `a` the first occurrence of its own colour?
    bool first = true;
    for (int b = 0; b < a; ++b) {
      if (jobs[b].page_index != page || results[b].status != ST_OK) continue;
      const int* cb = results[b].fill_bgr;
      if (cb[0] == ca[0] && cb[1] == ca[1] && cb[2] == ca[2]) {
        first = false;
        break;
      }
    }
    if (first) ++r;
  }
  rank[j] = r;
}

__global__ void paint_kernel(const Job* jobs, const Result* results, int n_jobs, uint8_t* const* pages_out,
                             const int* rank, int want_rank) {
  for (int j = blockIdx.x; j < n_jobs; j += gridDim.x) {
    if (rank[j] != want_rank) continue;
    const Job& J = jobs[j];
    const Result& R = results[j];
    const int cwords = (J.cw + 31) / 32;
    const uint32_t* fin = J.work + static_cast<size_t>(PL_FINAL) * cwords * J.ch;
    uint8_t* page = pages_out[J.page_index];
    const int n = cwords * J.ch * 32;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / (cwords * 32);
      const int x = i - y * cwords * 32;
      if (x >= J.cw) continue;
      if (!((fin[static_cast<size_t>(y) * cwords + (x >> 5)] >> (x & 31)) & 1u)) continue;
      uint8_t* px = page + static_cast<long long>(J.wy0 + y) * J.img_pitch + static_cast<long long>(J.wx0 + x) * J.img_c;
      px[0] = static_cast<uint8_t>(R.fill_bgr[0]);
      px[1] = static_cast<uint8_t>(R.fill_bgr[1]);
      px[2] = static_cast<uint8_t>(R.fill_bgr[2]);
    }
  }
}

__global__ void export_mask_kernel(const Job* jobs, int job_index, uint8_t* out, long long out_pitch, int pl) {
  const Job& J = jobs[job_index];
  const int cwords = (J.cw + 31) / 32;
  const uint32_t* src = J.work + static_cast<size_t>(pl) * cwords * J.ch;
  const long long n = static_cast<long long>(J.img_h) * J.img_w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int Y = static_cast<int>(i / J.img_w), X = static_cast<int>(i - static_cast<long long>(Y) * J.img_w);
    const int x = X - J.wx0, y = Y - J.wy0;
    uint8_t v = 0;
    if (x >= 0 && y >= 0 && x < J.cw && y < J.ch)
      v = ((src[static_cast<size_t>(y) * cwords + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
    out[static_cast<long long>(Y) * out_pitch + X] = v;
  }
}

}  // namespace

extern "C" {

unsigned long long mtb_clean_workspace_words(int cw, int ch, int max_runs) {
  return static_cast<unsigned long long>(workspace_words(cw, ch, max_runs));
}

int mtb_clean_bubbles(const mtb_clean_params* params, const mtb_clean_job* jobs_dev, mtb_clean_result* results_dev,
                      int n_jobs, void* stream) {
  MTB_REQUIRE(params && jobs_dev && results_dev, "mtb_clean_bubbles: null argument");
  if (n_jobs <= 0) return 0;
  MTB_REQUIRE(params->kd >= 1 && params->kd <= kMaxSE && (params->kd & 1), "clean: bad dilation kernel %d", params->kd);
  MTB_REQUIRE(params->ke >= 1 && params->ke <= kMaxSE && (params->ke & 1), "clean: bad erosion kernel %d", params->ke);
  MTB_REQUIRE(params->ball_r >= 0 && params->ball_r <= kMaxBall && params->jball_r >= 0 && params->jball_r <= kMaxBall,
              "clean: chamfer ball radius out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTB_CUDA_OK(cudaMemcpyToSymbolAsync(c_params, params, sizeof(Params), 0, cudaMemcpyHostToDevice, st));
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_jobs < 4 * sms ? n_jobs : 4 * sms;
  clean_kernel<<<grid, kCleanThreads, 0, st>>>(reinterpret_cast<const Job*>(jobs_dev),
                                               reinterpret_cast<Result*>(results_dev), n_jobs);
  MTB_CUDA_OK(cudaGetLastError());
  mtb::g_launches.fetch_add(1);
  return 0;
}

int mtb_clean_paint(const mtb_clean_job* jobs_dev, const mtb_clean_result* results_dev, int n_jobs,
                    uint8_t* const* pages_out_dev, int* rank_scratch_dev, int max_ranks, void* stream) {
  MTB_REQUIRE(jobs_dev && results_dev && pages_out_dev && rank_scratch_dev, "mtb_clean_paint: null argument");
  if (n_jobs <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rank_kernel<<<(n_jobs + 127) / 128, 128, 0, st>>>(reinterpret_cast<const Job*>(jobs_dev),
                                                    reinterpret_cast<const Result*>(results_dev), n_jobs,
                                                    rank_scratch_dev);
  MTB_CUDA_OK(cudaGetLastError());
  mtb::g_launches.fetch_add(1);
  int dev = 0, sms = 0;
  MTB_CUDA_OK(cudaGetDevice(&dev));
  MTB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = n_jobs < 8 * sms ? n_jobs : 8 * sms;
  for (int r = 0; r < max_ranks; ++r) {
    paint_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const Job*>(jobs_dev),
                                       reinterpret_cast<const Result*>(results_dev), n_jobs, pages_out_dev,
                                       rank_scratch_dev, r);
    MTB_CUDA_OK(cudaGetLastError());
    mtb::g_launches.fetch_add(1);
  }
  return 0;
}

int mtb_clean_export_mask(const mtb_clean_job* jobs_dev, int job_index, uint8_t* out, long long out_pitch, int plane,
                          void* stream) {
  MTB_REQUIRE(jobs_dev && out, "mtb_clean_export_mask: null argument");
  MTB_REQUIRE(plane >= 0 && plane < PL_COUNT, "mtb_clean_export_mask: bad plane %d", plane);
  export_mask_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const Job*>(jobs_dev),
                                                                         job_index, out, out_pitch, plane);
  MTB_CUDA_OK(cudaGetLastError());
  mtb::g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
