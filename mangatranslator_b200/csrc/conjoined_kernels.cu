// Conjoined-bubble mask splitting on the device — the pixel work of the reference's `_split_conjoined_mask`
// (core/image/detection.py:971-1035) and its helpers `_seed_mask_from_box` (:646-672), `_split_overlap_zone_with_line`
// (:675-800) and `_expand_resolved_masks_within_parent` (:932-968).
//
// The reference walks the full frame a dozen times per group in NumPy; here one parent mask goes in and K child masks
// come out of six small launches, all bit-exact:
//   * seed_k   = parent ∧ rect_k (rect_k = floor/ceil clip of child box k); an empty seed falls back to the parent
//     pixel nearest to the box centre (float64 squared distance, first in row-major order)
//   * every overlap zone parent ∧ rect_i ∧ rect_j is re-divided between i and j by the sign of one linear function
//     v(x, y) = (x - cx)*ax + (y - cy)*ay evaluated in float64 WITHOUT contraction (NumPy evaluates it as separate
//     IEEE ops); which line, which side and which comparison is geometry the host decides per pair (conjoined.py)
//   * parent pixels that no child owns go to the child with the smallest `cv2.distanceTransform(~seed, DIST_L2, 5)`,
//     first child on ties (np.argmin).  OpenCV's own implementation of that transform is a 16.16 fixed-point 5x5
//     chamfer (a = 1, b = 1.4, c = 2.1969); its value at p is min over seed pixels s of the chamfer norm of p - s
//     (closed form below; checked against cv2 with IPP disabled in tests/test_conjoined.py), and the norm is monotone
//     in |dx|, so per (pixel, child) only the nearest seed pixel of each ROW matters: a per-row nearest-seed table
//     turns the sequential two-pass raster scan into an embarrassingly parallel search that stops as soon as the row
//     distance alone exceeds the best value.  The float32 the reference compares is float(int) * 2^-16.
#include <math.h>
#include <string.h>

#include <atomic>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

constexpr int kMaxK = MTB_SPLIT_MAX_CHILDREN;
constexpr int kMaxPairs = kMaxK * (kMaxK - 1) / 2;
constexpr unsigned int kChamferA = 65536u, kChamferB = 91750u, kChamferC = 143976u;   // CV_FLT_TO_FIX(1 / 1.4 / 2.1969, 16)
constexpr uint16_t kRemaining = 0x8000u;
constexpr uint16_t kNoSeedInRow = 0xffffu;

struct SplitParams {
  int H, W, K, n_pairs;
  int wx0, wy0, ww, wh;           // window that contains every parent pixel and every child rectangle
  int rect[kMaxK][4];
  double center[kMaxK][2];
  mtb_split_pair pair[kMaxPairs];
};

struct SplitFlags {
  int rect_seed_any[kMaxK];       // parent ∧ rect_k has a pixel
  int seed_any[kMaxK];            // child k owns a pixel after the overlap zones were divided
  int any_remaining;
  int pad;
  unsigned long long dmin[kMaxK]; // empty-seed fallback: bits of the smallest squared distance / its first pixel
  unsigned long long imin[kMaxK];
};

__device__ __forceinline__ bool in_rect(const int* r, int x, int y) { return x >= r[0] && x < r[2] && y >= r[1] && y < r[3]; }

// set-once flag: after the first few writers everybody sees it set and skips the atomic
__device__ __forceinline__ void warp_flag(int* flag, bool v) {
  if (v && *reinterpret_cast<volatile int*>(flag) == 0) atomicOr(flag, 1);
}

__global__ void split_scan_parent_kernel(const uint8_t* __restrict__ parent, const SplitParams* __restrict__ P,
                                         SplitFlags* __restrict__ F) {
  const int ww = P->ww, wh = P->wh, K = P->K;
  const long long total = static_cast<long long>(ww) * wh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = P->wy0 + static_cast<int>(i / ww), x = P->wx0 + static_cast<int>(i % ww);
    if (!parent[static_cast<long long>(y) * P->W + x]) continue;
    for (int k = 0; k < K; ++k)
      if (in_rect(P->rect[k], x, y)) warp_flag(&F->rect_seed_any[k], true);
  }
}

// pass 0: smallest squared distance of a parent pixel to the centre of every child box whose seed is empty
// pass 1: the first (row-major) pixel that attains it
__global__ void split_nearest_parent_kernel(const uint8_t* __restrict__ parent, const SplitParams* __restrict__ P,
                                            SplitFlags* __restrict__ F, int pass) {
  const int K = P->K;
  bool need = false;
  for (int k = 0; k < K; ++k) need |= (F->rect_seed_any[k] == 0);
  if (!need) return;
  const int ww = P->ww, wh = P->wh;
  const long long total = static_cast<long long>(ww) * wh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = P->wy0 + static_cast<int>(i / ww), x = P->wx0 + static_cast<int>(i % ww);
    if (!parent[static_cast<long long>(y) * P->W + x]) continue;
    for (int k = 0; k < K; ++k) {
      if (F->rect_seed_any[k]) continue;
      const double ex = __dsub_rn(static_cast<double>(x), P->center[k][0]);
      const double ey = __dsub_rn(static_cast<double>(y), P->center[k][1]);
      const double d = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
      const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(d));   // d >= 0: order-preserving
      if (pass == 0) {
        atomicMin(&F->dmin[k], bits);
      } else if (bits == F->dmin[k]) {
        atomicMin(&F->imin[k], static_cast<unsigned long long>(y) * P->W + x);
      }
    }
  }
}

__global__ void split_resolve_kernel(const uint8_t* __restrict__ parent, const SplitParams* __restrict__ P,
                                     SplitFlags* __restrict__ F, uint16_t* __restrict__ bits_out) {
  const int ww = P->ww, wh = P->wh, K = P->K, np = P->n_pairs;
  const long long total = static_cast<long long>(ww) * wh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = P->wy0 + static_cast<int>(i / ww), x = P->wx0 + static_cast<int>(i % ww);
    const bool base = parent[static_cast<long long>(y) * P->W + x] != 0;
    unsigned int owned = 0, inr = 0;
    if (base) {
      for (int k = 0; k < K; ++k) {
        const bool r = in_rect(P->rect[k], x, y);
        inr |= (r ? 1u : 0u) << k;
        const bool seed = r || (F->rect_seed_any[k] == 0 && F->imin[k] == static_cast<unsigned long long>(y) * P->W + x);
        owned |= (seed ? 1u : 0u) << k;
      }
      for (int q = 0; q < np; ++q) {
        const mtb_split_pair& pr = P->pair[q];
        const unsigned int bi = 1u << pr.i, bj = 1u << pr.j;
        if ((inr & bi) && (inr & bj)) {                 // pixel of the overlap zone parent ∧ rect_i ∧ rect_j
          owned &= ~(bi | bj);
          if (pr.mode != 0) {
            // signed distance to the dividing line minus the text-safe offset of the pair (0 without OSB text boxes),
            // operation for operation like the reference's `pixel_dist - split_offset` (detection.py:684-687, :761)
            const double v = __dsub_rn(__dadd_rn(__dmul_rn(__dsub_rn(static_cast<double>(x), pr.cx), pr.ax),
                                                 __dmul_rn(__dsub_rn(static_cast<double>(y), pr.cy), pr.ay)),
                                       pr.off);
            const bool to_i = pr.mode == 1 ? (v <= 0.0) : (v >= 0.0);
            const bool to_j = pr.mode == 1 ? (v > 0.0) : (v < 0.0);
            if (to_i) owned |= bi;
            if (to_j) owned |= bj;
          }
        }
      }
    }
    const bool remaining = base && owned == 0;
    for (int k = 0; k < K; ++k) warp_flag(&F->seed_any[k], ((owned >> k) & 1u) != 0);
    warp_flag(&F->any_remaining, remaining);
    bits_out[i] = static_cast<uint16_t>(owned | (remaining ? kRemaining : 0u));
  }
}

// one WARP per (child, window row): distance to the nearest pixel of that child in the row, for every column.  Each
// 32-column chunk is one coalesced load and one ballot; the nearest seed at-or-left of a lane is the highest set bit at
// or below it (or the carry from earlier chunks), the nearest at-or-right the lowest set bit at or above it.
__global__ void split_rowdist_kernel(const SplitParams* __restrict__ P, const SplitFlags* __restrict__ F,
                                     const uint16_t* __restrict__ bits, uint16_t* __restrict__ rowdist) {
  if (!F->any_remaining) return;
  const int ww = P->ww, wh = P->wh, K = P->K;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= K * wh) return;
  const int k = t / wh, row = t - k * wh;
  if (!F->seed_any[k]) return;
  const uint16_t* b = bits + static_cast<long long>(row) * ww;
  uint16_t* rd = rowdist + (static_cast<long long>(k) * wh + row) * ww;
  int last = -1;
  for (int x0 = 0; x0 < ww; x0 += 32) {
    const int x = x0 + lane;
    const bool s = x < ww && ((b[x] >> k) & 1u);
    const unsigned m = __ballot_sync(0xffffffffu, s);
    const unsigned below = m & (0xffffffffu >> (31 - lane));
    const int l = below ? x0 + 31 - __clz(below) : last;
    if (x < ww) rd[x] = l < 0 ? kNoSeedInRow : static_cast<uint16_t>(min(x - l, 65534));
    if (m) last = x0 + 31 - __clz(m);
  }
  int next = -1;
  for (int x0 = ((ww - 1) / 32) * 32; x0 >= 0; x0 -= 32) {
    const int x = x0 + lane;
    const bool s = x < ww && ((b[x] >> k) & 1u);
    const unsigned m = __ballot_sync(0xffffffffu, s);
    const unsigned above = m & (0xffffffffu << lane);
    const int r = above ? x0 + __ffs(above) - 1 : next;
    if (x < ww && r >= 0) {
      const int d = min(r - x, 65534);
      if (d < rd[x]) rd[x] = static_cast<uint16_t>(d);      // same lane wrote rd[x] in the forward sweep
    }
    if (m) next = x0 + __ffs(m) - 1;
  }
}

__device__ __forceinline__ unsigned int chamfer_norm(unsigned int dx, unsigned int dy) {
  const unsigned int M = max(dx, dy), m = min(dx, dy);
  return (M >= 2u * m) ? m * kChamferC + (M - 2u * m) * kChamferA : (M - m) * kChamferC + (2u * m - M) * kChamferB;
}

// remaining parent pixels go to the nearest child; then every child's full-frame uint8 mask is written
__global__ void split_expand_kernel(const SplitParams* __restrict__ P, const SplitFlags* __restrict__ F,
                                    const uint16_t* __restrict__ bits, const uint16_t* __restrict__ rowdist,
                                    uint8_t* __restrict__ out) {
  const int H = P->H, W = P->W, K = P->K, ww = P->ww, wh = P->wh;
  const long long total = static_cast<long long>(H) * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
    const int wy = y - P->wy0, wx = x - P->wx0;
    unsigned int owned = 0;
    if (wy >= 0 && wy < wh && wx >= 0 && wx < ww) {
      const uint16_t b = bits[static_cast<long long>(wy) * ww + wx];
      owned = b & 0x7fffu;
      if (b & kRemaining) {
        int best_k = 0;
        float best_f = INFINITY;
        for (int k = 0; k < K; ++k) {
          float fk = INFINITY;                          // the reference's distance map of an empty seed is +inf
          if (F->seed_any[k]) {
            const uint16_t* rd = rowdist + static_cast<long long>(k) * wh * ww + wx;
            unsigned int best = 0xffffffffu;
            for (int t = 0; t < wh; ++t) {
              if (static_cast<unsigned long long>(t) * kChamferA >= best) break;   // rows further away cannot win
              const int r0 = wy - t, r1 = wy + t;
              if (r0 < 0 && r1 >= wh) break;
              if (r0 >= 0) {
                const uint16_t d = rd[static_cast<long long>(r0) * ww];
                if (d != kNoSeedInRow) best = min(best, chamfer_norm(d, static_cast<unsigned int>(t)));
              }
              if (t > 0 && r1 < wh) {
                const uint16_t d = rd[static_cast<long long>(r1) * ww];
                if (d != kNoSeedInRow) best = min(best, chamfer_norm(d, static_cast<unsigned int>(t)));
              }
            }
            fk = __uint2float_rn(best);                 // (float)(t0 * 2^-16): the power-of-two scale changes no order
          }
          if (fk < best_f) {                            // np.argmin keeps the first minimum
            best_f = fk;
            best_k = k;
          }
        }
        owned |= 1u << best_k;
      }
    }
    for (int k = 0; k < K; ++k) out[static_cast<long long>(k) * total + i] = ((owned >> k) & 1u) ? 255 : 0;
  }
}

int sm_count_c() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" {

long long mtb_split_conjoined_workspace_bytes(int win_h, int win_w, int K) {
  if (win_h <= 0 || win_w <= 0 || K <= 0 || K > kMaxK) return -1;
  const size_t px = static_cast<size_t>(win_h) * win_w;
  return static_cast<long long>(align_up(sizeof(SplitParams), 256) + align_up(sizeof(SplitFlags), 256) +
                                align_up(px * sizeof(uint16_t), 256) + align_up(px * K * sizeof(uint16_t), 256));
}

int mtb_split_conjoined(const uint8_t* parent, int H, int W, int K, const int* rects, const double* centers,
                        const int* window, int n_pairs, const mtb_split_pair* pairs, uint8_t* out, void* workspace,
                        long long workspace_bytes, void* stream) {
  MTB_REQUIRE(parent && out && workspace && rects && centers && window, "mtb_split_conjoined: null argument");
  MTB_REQUIRE(K >= 1 && K <= kMaxK && n_pairs >= 0 && n_pairs <= kMaxPairs && (n_pairs == 0 || pairs) && H > 0 && W > 0,
              "mtb_split_conjoined: bad sizes (K=%d, pairs=%d)", K, n_pairs);
  const int wx0 = window[0], wy0 = window[1], ww = window[2] - window[0], wh = window[3] - window[1];
  MTB_REQUIRE(wx0 >= 0 && wy0 >= 0 && ww > 0 && wh > 0 && wx0 + ww <= W && wy0 + wh <= H && ww + wh < 29000,
              "mtb_split_conjoined: bad window (32-bit 16.16 chamfer sums hold up to 29000 px)");
  MTB_REQUIRE(workspace_bytes >= mtb_split_conjoined_workspace_bytes(wh, ww, K), "mtb_split_conjoined: workspace too small");
  SplitParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.H = H;
  hp.W = W;
  hp.K = K;
  hp.n_pairs = n_pairs;
  hp.wx0 = wx0;
  hp.wy0 = wy0;
  hp.ww = ww;
  hp.wh = wh;
  for (int k = 0; k < K; ++k) {
    for (int c = 0; c < 4; ++c) hp.rect[k][c] = rects[k * 4 + c];
    const bool empty = hp.rect[k][2] <= hp.rect[k][0] || hp.rect[k][3] <= hp.rect[k][1];
    const bool inside = hp.rect[k][0] >= wx0 && hp.rect[k][1] >= wy0 && hp.rect[k][2] <= wx0 + ww && hp.rect[k][3] <= wy0 + wh;
    MTB_REQUIRE(empty || inside, "mtb_split_conjoined: child rectangle %d outside the window", k);
    hp.center[k][0] = centers[k * 2];
    hp.center[k][1] = centers[k * 2 + 1];
  }
  for (int q = 0; q < n_pairs; ++q) {
    MTB_REQUIRE(pairs[q].i >= 0 && pairs[q].i < pairs[q].j && pairs[q].j < K && pairs[q].mode >= 0 && pairs[q].mode <= 2,
                "mtb_split_conjoined: bad pair %d", q);
    hp.pair[q] = pairs[q];
  }
  SplitFlags hf;
  memset(&hf, 0, sizeof(hf));
  for (int k = 0; k < kMaxK; ++k) hf.dmin[k] = hf.imin[k] = ~0ull;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  SplitParams* dP = reinterpret_cast<SplitParams*>(ws);
  ws += align_up(sizeof(SplitParams), 256);
  SplitFlags* dF = reinterpret_cast<SplitFlags*>(ws);
  ws += align_up(sizeof(SplitFlags), 256);
  uint16_t* dBits = reinterpret_cast<uint16_t*>(ws);
  ws += align_up(static_cast<size_t>(wh) * ww * sizeof(uint16_t), 256);
  uint16_t* dRow = reinterpret_cast<uint16_t*>(ws);
  // pageable sources: staged before the calls return
  MTB_CUDA_OK(cudaMemcpyAsync(dP, &hp, sizeof(hp), cudaMemcpyHostToDevice, st));
  MTB_CUDA_OK(cudaMemcpyAsync(dF, &hf, sizeof(hf), cudaMemcpyHostToDevice, st));
  const long long wpx = static_cast<long long>(wh) * ww;
  const int sms = sm_count_c();
  auto grid = [&](long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(sms) * 16)); };
  split_scan_parent_kernel<<<grid(wpx), 256, 0, st>>>(parent, dP, dF);
  MTB_CUDA_OK(cudaGetLastError());
  split_nearest_parent_kernel<<<grid(wpx), 256, 0, st>>>(parent, dP, dF, 0);
  MTB_CUDA_OK(cudaGetLastError());
  split_nearest_parent_kernel<<<grid(wpx), 256, 0, st>>>(parent, dP, dF, 1);
  MTB_CUDA_OK(cudaGetLastError());
  split_resolve_kernel<<<grid(wpx), 256, 0, st>>>(parent, dP, dF, dBits);
  MTB_CUDA_OK(cudaGetLastError());
  split_rowdist_kernel<<<(K * wh + 7) / 8, 256, 0, st>>>(dP, dF, dBits, dRow);      // one warp per (child, row)
  MTB_CUDA_OK(cudaGetLastError());
  split_expand_kernel<<<grid(static_cast<long long>(H) * W), 256, 0, st>>>(dP, dF, dBits, dRow, out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(6);
  return 0;
}

}  // extern "C"
