// Lossless PNG encoding of a finished page ON THE DEVICE ("page I/O ... PNG encode becomes the
// bottleneck once compute is fast").  The reference writes every page with PIL's PNG encoder (+ oxipng) on a host thread
// (core/image/image_utils.py:59-170 save_image_with_compression, called from core/pipeline.py:1996-2018); a 3072x2048
// RGBA page costs ~1 s of one host core (measured), i.e. 8 cores per GPU at this build's page rate.  Here the page never
// leaves the device uncompressed:
//
//   png_filter_kernel     one CTA per scanline: the five PNG filters (None / Sub / Up / Average / Paeth) are scored with
//                         libpng's minimum-sum-of-absolute-differences heuristic, the best one is applied, and the
//                         filter-type byte + filtered bytes go to the "filtered stream" (what zlib would be fed).  RGB in,
//                         RGB or RGBA (opaque alpha) out.
//   png_histogram_kernel  the stream is cut into 16 KB segments (one deflate block each) and every segment into 256 spans
//                         of 64 bytes; a span is tokenised greedily into literals and distance-1 matches (runs of a
//                         repeated byte, length 3..63: the flat regions of a page).  Symbol histogram for ONE dynamic
//                         Huffman table per image + the Adler-32 partial sums of every segment.
//   (host, ~0.3 ms)       length-limited Huffman code from the 286-bin histogram, block header bits
//   png_deflate_kernel    one CTA per segment: same tokenisation, bit lengths -> block-wide prefix sum -> every thread ORs
//                         its codes into the block's bit buffer in shared memory -> header + tokens + end-of-block + an
//                         empty stored block that byte-aligns the stream (what Z_SYNC_FLUSH emits, as pigz does), so the
//                         blocks concatenate as bytes
//   png_compact_kernel    segments -> one contiguous deflate stream (offsets from a prefix sum of the block sizes)
//
// The host adds the 8-byte signature, IHDR, the zlib header / Adler-32 trailer, chunk CRCs and IEND around the stream it
// copies back (~0.4 x the raw bytes), ~10 ms of one core instead of ~1000.  Integer / byte work, HBM-bound: the page is
// read twice (filter scoring + application, L2-resident second time), the filtered stream written once and read twice.
#include <atomic>

#include "../../include/mtb200.h"
#include "common.cuh"

namespace mtb {
extern std::atomic<long long> g_launches;
}
using namespace mtb;

namespace {

constexpr int kSeg = MTB_PNG_SEGMENT;            // bytes of filtered stream per deflate block
constexpr int kSpan = 64;                        // bytes per thread
constexpr int kThreads = kSeg / kSpan;           // 256
constexpr int kMaxHeaderWords = 96;              // dynamic-block header: <= 3072 bits
constexpr int kBufWords = (kSeg * 15 + 4096) / 32 + kMaxHeaderWords + 8;   // worst case: every literal 15 bits

__device__ __forceinline__ int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// pixel byte k of output row y (oc channels; channel 3 = opaque alpha), 0 outside the image (PNG's convention)
__device__ __forceinline__ int px(const uint8_t* __restrict__ img, int W, int ic, int oc, int y, int k) {
  if (y < 0 || k < 0) return 0;
  const int x = k / oc, c = k - x * oc;
  return c < ic ? img[(static_cast<long long>(y) * W + x) * ic + c] : 255;
}

__global__ void __launch_bounds__(256)
png_filter_kernel(const uint8_t* __restrict__ img, int H, int W, int ic, int oc, uint8_t* __restrict__ stream) {
  const int y = blockIdx.x;
  const int bpr = W * oc;
  __shared__ unsigned int cost[5];
  __shared__ int best_s;
  if (threadIdx.x < 5) cost[threadIdx.x] = 0;
  __syncthreads();
  unsigned int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  for (int k = threadIdx.x; k < bpr; k += blockDim.x) {
    const int x = px(img, W, ic, oc, y, k), a = px(img, W, ic, oc, y, k - oc), b = px(img, W, ic, oc, y - 1, k),
              c = px(img, W, ic, oc, y - 1, k - oc);
    // libpng's heuristic: sum of the residuals read as signed bytes, in absolute value
    auto mag = [](int r) { r &= 255; return static_cast<unsigned int>(r < 128 ? r : 256 - r); };
    c0 += mag(x);
    c1 += mag(x - a);
    c2 += mag(x - b);
    c3 += mag(x - ((a + b) >> 1));
    c4 += mag(x - paeth(a, b, c));
  }
  atomicAdd(&cost[0], c0);
  atomicAdd(&cost[1], c1);
  atomicAdd(&cost[2], c2);
  atomicAdd(&cost[3], c3);
  atomicAdd(&cost[4], c4);
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int f = 1; f < 5; ++f)
      if (cost[f] < cost[best]) best = f;          // first minimum in filter order, like libpng
    best_s = best;
  }
  __syncthreads();
  const int f = best_s;
  uint8_t* row = stream + static_cast<long long>(y) * (bpr + 1);
  if (threadIdx.x == 0) row[0] = static_cast<uint8_t>(f);
  for (int k = threadIdx.x; k < bpr; k += blockDim.x) {
    const int x = px(img, W, ic, oc, y, k);
    int pred = 0;
    if (f == 1) pred = px(img, W, ic, oc, y, k - oc);
    else if (f == 2) pred = px(img, W, ic, oc, y - 1, k);
    else if (f == 3) pred = (px(img, W, ic, oc, y, k - oc) + px(img, W, ic, oc, y - 1, k)) >> 1;
    else if (f == 4) pred = paeth(px(img, W, ic, oc, y, k - oc), px(img, W, ic, oc, y - 1, k), px(img, W, ic, oc, y - 1, k - oc));
    row[1 + k] = static_cast<uint8_t>(x - pred);
  }
}

// length 3..258 -> (length symbol, number of extra bits, extra value)
__device__ __forceinline__ void length_code(int len, int& sym, int& ebits, int& extra) {
  if (len <= 10) {
    sym = 254 + len;
    ebits = 0;
    extra = 0;
  } else if (len == 258) {
    sym = 285;
    ebits = 0;
    extra = 0;
  } else {
    const int l = len - 3;
    const int e = 29 - __clz(l);                   // floor(log2(l)) - 2
    sym = 261 + 4 * e + ((l >> e) & 3);
    ebits = e;
    extra = l & ((1 << e) - 1);
  }
}

// Greedy tokenisation of one span: the first byte is a literal; afterwards a run of >= 3 bytes equal to the byte before
// it becomes one match (length = run, distance 1), anything else a literal.  `lit(byte)` / `match(len)` are called in
// stream order.  Identical in the histogram and the encoding pass by construction.
template <class L, class M>
__device__ __forceinline__ void tokenize_span(const uint8_t* __restrict__ p, int n, L&& lit, M&& match) {
  if (n <= 0) return;
  lit(p[0]);
  int i = 1;
  while (i < n) {
    const uint8_t prev = p[i - 1];
    int r = 0;
    while (i + r < n && p[i + r] == prev) ++r;
    if (r >= 3) {
      match(r);
      i += r;
    } else {
      lit(p[i]);
      ++i;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
png_histogram_kernel(const uint8_t* __restrict__ stream, long long total, unsigned int* __restrict__ hist /* [288] */,
                     unsigned long long* __restrict__ adler /* [segments][2] */) {
  __shared__ unsigned int h[288];
  __shared__ unsigned long long s1s, s2s;
  for (int i = threadIdx.x; i < 288; i += blockDim.x) h[i] = 0;
  if (threadIdx.x == 0) s1s = s2s = 0;
  __syncthreads();
  const long long seg0 = static_cast<long long>(blockIdx.x) * kSeg;
  const long long seg_n = min(static_cast<long long>(kSeg), total - seg0);
  const long long off = seg0 + static_cast<long long>(threadIdx.x) * kSpan;
  const int n = static_cast<int>(max(0ll, min(static_cast<long long>(kSpan), total - off)));
  uint8_t buf[kSpan];
  if (n == kSpan) {
    const uint4* v = reinterpret_cast<const uint4*>(stream + off);
#pragma unroll
    for (int q = 0; q < kSpan / 16; ++q) reinterpret_cast<uint4*>(buf)[q] = v[q];
  } else {
    for (int i = 0; i < n; ++i) buf[i] = stream[off + i];
  }
  tokenize_span(buf, n, [&](uint8_t b) { atomicAdd(&h[b], 1u); },
                [&](int len) {
                  int sym, eb, ex;
                  length_code(len, sym, eb, ex);
                  atomicAdd(&h[sym], 1u);
                });
  // Adler-32 partials of the segment: s1 = sum of bytes, s2 = sum of (seg_n - index) * byte
  unsigned long long s1 = 0, s2 = 0;
  const long long base = static_cast<long long>(threadIdx.x) * kSpan;
  for (int i = 0; i < n; ++i) {
    s1 += buf[i];
    s2 += static_cast<unsigned long long>(seg_n - (base + i)) * buf[i];
  }
  atomicAdd(&s1s, s1);
  atomicAdd(&s2s, s2);
  __syncthreads();
  for (int i = threadIdx.x; i < 288; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
  if (threadIdx.x == 0) {
    atomicAdd(&hist[256], 1u);                     // one end-of-block per segment
    adler[2 * blockIdx.x] = s1s;
    adler[2 * blockIdx.x + 1] = s2s;
  }
}

struct BitWriter {
  unsigned int* words;
  unsigned long long acc;
  int nacc;
  int wpos;
  __device__ __forceinline__ void init(unsigned int* w, unsigned int bit_offset) {
    words = w;
    wpos = static_cast<int>(bit_offset >> 5);
    nacc = static_cast<int>(bit_offset & 31);
    acc = 0;
  }
  __device__ __forceinline__ void put(unsigned int value, int nbits) {     // LSB-first, nbits <= 24
    acc |= static_cast<unsigned long long>(value) << nacc;
    nacc += nbits;
    if (nacc >= 32) {
      atomicOr(&words[wpos], static_cast<unsigned int>(acc));
      ++wpos;
      acc >>= 32;
      nacc -= 32;
    }
  }
  __device__ __forceinline__ void flush() {
    if (nacc > 0) atomicOr(&words[wpos], static_cast<unsigned int>(acc));
  }
};

__global__ void __launch_bounds__(kThreads)
png_deflate_kernel(const uint8_t* __restrict__ stream, long long total, const unsigned short* __restrict__ code /* [288] bit-reversed */,
                   const unsigned char* __restrict__ clen /* [288] */, const unsigned int* __restrict__ header, int header_bits,
                   uint8_t* __restrict__ staged, int stride, unsigned int* __restrict__ sizes) {
  __shared__ unsigned int words[kBufWords];
  __shared__ unsigned short scode[288];
  __shared__ unsigned char slen[288];
  __shared__ unsigned int warp_sum[kThreads / 32];
  for (int i = threadIdx.x; i < kBufWords; i += blockDim.x) words[i] = 0;
  for (int i = threadIdx.x; i < 288; i += blockDim.x) {
    scode[i] = code[i];
    slen[i] = clen[i];
  }
  __syncthreads();
  const long long seg0 = static_cast<long long>(blockIdx.x) * kSeg;
  const bool last = seg0 + kSeg >= total;
  const long long off = seg0 + static_cast<long long>(threadIdx.x) * kSpan;
  const int n = static_cast<int>(max(0ll, min(static_cast<long long>(kSpan), total - off)));
  uint8_t buf[kSpan];
  if (n == kSpan) {
    const uint4* v = reinterpret_cast<const uint4*>(stream + off);
#pragma unroll
    for (int q = 0; q < kSpan / 16; ++q) reinterpret_cast<uint4*>(buf)[q] = v[q];
  } else {
    for (int i = 0; i < n; ++i) buf[i] = stream[off + i];
  }
  // pass 1: bits of this span
  unsigned int bits = 0;
  tokenize_span(buf, n, [&](uint8_t b) { bits += slen[b]; },
                [&](int len) {
                  int sym, eb, ex;
                  length_code(len, sym, eb, ex);
                  bits += slen[sym] + eb + 1;      // + the one-bit distance code (distance 1)
                });
  // block-wide exclusive prefix sum
  unsigned int incl = bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  unsigned int before = 0, all = 0;
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < warp) before += warp_sum[w];
    all += warp_sum[w];
  }
  const unsigned int start = static_cast<unsigned int>(header_bits) + before + incl - bits;
  // header (BFINAL = 0 in the stored copy; the last block sets bit 0)
  for (int i = threadIdx.x; i < (header_bits + 31) / 32; i += blockDim.x) atomicOr(&words[i], header[i]);
  if (last && threadIdx.x == 0) atomicOr(&words[0], 1u);
  // pass 2: codes
  BitWriter bw;
  bw.init(words, start);
  tokenize_span(buf, n, [&](uint8_t b) { bw.put(scode[b], slen[b]); },
                [&](int len) {
                  int sym, eb, ex;
                  length_code(len, sym, eb, ex);
                  bw.put(scode[sym], slen[sym]);
                  bw.put(static_cast<unsigned int>(ex), eb + 1);           // extra bits, then distance code '0'
                });
  bw.flush();
  unsigned int end_bits = static_cast<unsigned int>(header_bits) + all;
  if (threadIdx.x == 0) {
    BitWriter tail;
    tail.init(words, end_bits);
    tail.put(scode[256], slen[256]);                                      // end of block
    tail.flush();
  }
  end_bits += slen[256];
  unsigned int nbytes;
  if (last) {
    nbytes = (end_bits + 7) >> 3;
  } else {
    // empty stored block: BFINAL = 0, BTYPE = 00 (three zero bits, already zero), pad to a byte, LEN = 0, NLEN = 0xFFFF
    const unsigned int b0 = (end_bits + 3 + 7) >> 3;
    nbytes = b0 + 4;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (unsigned int k = 0; k < 2; ++k) {
        const unsigned int byte = b0 + 2 + k;
        atomicOr(&words[byte >> 2], 0xFFu << ((byte & 3) * 8));
      }
    }
  }
  __syncthreads();
  uint8_t* dst = staged + static_cast<long long>(blockIdx.x) * stride;
  for (unsigned int i = threadIdx.x; i < (nbytes + 3) / 4; i += blockDim.x) reinterpret_cast<unsigned int*>(dst)[i] = words[i];
  if (threadIdx.x == 0) sizes[blockIdx.x] = nbytes;
}

__global__ void __launch_bounds__(256)
png_compact_kernel(const uint8_t* __restrict__ staged, int stride, const unsigned int* __restrict__ sizes,
                   const long long* __restrict__ offsets /* exclusive prefix sum of sizes */, uint8_t* __restrict__ out) {
  const uint8_t* src = staged + static_cast<long long>(blockIdx.x) * stride;
  uint8_t* dst = out + offsets[blockIdx.x];
  const unsigned int n = sizes[blockIdx.x];
  for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" {

int mtb_png_filter(const uint8_t* img, int H, int W, int in_channels, int out_channels, uint8_t* stream, void* stream_) {
  MTB_REQUIRE(img && stream && H > 0 && W > 0, "mtb_png_filter: bad arguments");
  MTB_REQUIRE((in_channels == 3 || in_channels == 4) && (out_channels == 3 || out_channels == 4),
              "mtb_png_filter: 3 or 4 channels");
  png_filter_kernel<<<H, 256, 0, static_cast<cudaStream_t>(stream_)>>>(img, H, W, in_channels, out_channels, stream);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_png_histogram(const uint8_t* stream, long long total, unsigned int* hist, unsigned long long* adler_parts,
                      void* stream_) {
  MTB_REQUIRE(stream && hist && adler_parts && total > 0, "mtb_png_histogram: bad arguments");
  MTB_REQUIRE((reinterpret_cast<uintptr_t>(stream) & 15) == 0, "mtb_png_histogram: the stream must be 16-byte aligned");
  const int segs = static_cast<int>((total + kSeg - 1) / kSeg);
  png_histogram_kernel<<<segs, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(stream, total, hist, adler_parts);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_png_deflate(const uint8_t* stream, long long total, const unsigned short* code, const unsigned char* code_len,
                    const unsigned int* header, int header_bits, uint8_t* staged, int stride, unsigned int* sizes,
                    void* stream_) {
  MTB_REQUIRE(stream && code && code_len && header && staged && sizes && total > 0, "mtb_png_deflate: bad arguments");
  MTB_REQUIRE(header_bits > 0 && header_bits <= kMaxHeaderWords * 32, "mtb_png_deflate: header of %d bits", header_bits);
  MTB_REQUIRE(stride >= MTB_PNG_SEGMENT_STRIDE && stride % 4 == 0, "mtb_png_deflate: stride too small");
  const int segs = static_cast<int>((total + kSeg - 1) / kSeg);
  png_deflate_kernel<<<segs, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(stream, total, code, code_len, header,
                                                                             header_bits, staged, stride, sizes);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mtb_png_compact(const uint8_t* staged, int stride, const unsigned int* sizes, const long long* offsets, int segments,
                    uint8_t* out, void* stream_) {
  MTB_REQUIRE(staged && sizes && offsets && out && segments > 0, "mtb_png_compact: bad arguments");
  png_compact_kernel<<<segments, 256, 0, static_cast<cudaStream_t>(stream_)>>>(staged, stride, sizes, offsets, out);
  MTB_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
