// C-ABI: library plumbing + conv plan entry points (include/mtb200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>

#include "../../include/mtb200.h"
#include "common.cuh"
#include "conv_gemm.cuh"

namespace mtb {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  MTB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
  return 0;
}

}  // namespace mtb

using namespace mtb;

namespace mtb {
bool conv_halo_eligible(const ConvParams& p, int cin);
void conv_halo_cm_tile(int* tw, int* th);
bool conv_halo_cm_eligible(const ConvParams& p);
int launch_conv_halo_cm(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, const ConvParams& p,
                        cudaStream_t stream);
int launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int nsplit,
                     cudaStream_t stream);
void conv_halo_fp16c_tile(int* tw, int* th);
int launch_conv_halo_fp16c(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmR, const ConvParams& p,
                           cudaStream_t stream);
int launch_bf16x2_to_fp16c(const void* in, long long npix, void* out, float lo_scale, cudaStream_t stream);
int launch_fp16c_to_bf16x2(const void* in, long long npix, void* out, float lo_inv, cudaStream_t stream);
}  // namespace mtb

struct mtb_conv_plan {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmO;   // channel-major halo kernel with staged stores: the output planes, box = 64 ch x 8 px x 2 rows
  ConvParams p;
  int nsplit;
  int halo;
};

extern "C" {

const char* mtb_last_error(void) { return mtb::last_error(); }
int mtb_version(void) { return 100; }
long long mtb_launch_count(void) { return mtb::g_launches.load(); }

int mtb_conv_plan_create(const mtb_conv_desc* d, const void* x, const void* w, const float* bias, void* out,
                         const void* residual, float* tile_sums, mtb_conv_plan** plan_out) {
  MTB_REQUIRE(d && x && w && out && plan_out, "mtb_conv_plan_create: null argument");
  MTB_REQUIRE(d->Cin % 64 == 0 && d->Cin > 0, "conv: Cin (%d) must be padded to a multiple of 64", d->Cin);
  MTB_REQUIRE(d->Cout % 16 == 0 && d->Cout > 0, "conv: Cout (%d) must be padded to a multiple of 16", d->Cout);
  MTB_REQUIRE(d->planes_in == 1 || d->planes_in == 2, "conv: planes_in must be 1 or 2");
  MTB_REQUIRE(d->planes_out == 1 || d->planes_out == 2 || d->planes_out == 4, "conv: planes_out must be 1, 2 or 4");
  MTB_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  MTB_REQUIRE(d->res_planes >= 0 && d->res_planes <= 2, "conv: res_planes must be 0..2");
  MTB_REQUIRE((d->res_planes == 0) == (residual == nullptr), "conv: residual pointer / res_planes mismatch");
  MTB_REQUIRE(!(d->planes_out == 4 && residual), "conv: fp32 output with residual is not supported");
  MTB_REQUIRE((d->x_ctotal == 0 || d->x_ctotal % 8 == 0) && d->x_coff % 8 == 0 && d->out_coff % 8 == 0 &&
                  d->res_coff % 8 == 0 && (d->out_ctotal == 0 || d->out_ctotal % 8 == 0) &&
                  (d->res_ctotal == 0 || d->res_ctotal % 8 == 0),
              "conv: channel totals / offsets must be multiples of 8 (16-byte rows)");

  mtb_conv_plan* pl = new (std::nothrow) mtb_conv_plan();
  MTB_REQUIRE(pl != nullptr, "conv: out of host memory");
  ConvParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.N = d->N;
  p.H = d->H;
  p.W = d->W;
  p.KH = d->KH;
  p.KW = d->KW;
  p.pad = d->pad;
  p.stride = d->stride;
  p.Ho = (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  p.Wo = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  p.cin_chunks = d->Cin / 64;
  p.Cout = d->Cout;
  if (d->tile_w > 0 && d->tile_h > 0) {
    p.TW = d->tile_w;
    p.TH = d->tile_h;
  } else if (p.Ho == 1) {
    p.TW = 128;
    p.TH = 1;
  } else if (p.Wo <= 8) {
    p.TW = 8;
    p.TH = 16;
  } else {
    p.TW = 16;
    p.TH = 8;
  }
  if (p.TW * p.TH != 128) {
    const int tw = p.TW, th = p.TH;
    delete pl;
    MTB_REQUIRE(false, "conv: tile %dx%d must hold 128 pixels", tw, th);
  }
  p.planes_out = d->planes_out == 4 ? 1 : d->planes_out;
  p.act = d->act;
  p.in_coff = d->x_coff;
  p.out_cstride = d->out_ctotal > 0 ? d->out_ctotal : d->Cout;
  p.out_coff = d->out_coff;
  p.res_cstride = d->res_ctotal > 0 ? d->res_ctotal : d->Cout;
  p.res_coff = d->res_coff;
  p.out_plane_stride = static_cast<long long>(p.N) * p.Ho * p.Wo * p.out_cstride;
  p.res_plane_stride = static_cast<long long>(d->res_bcast ? 1 : p.N) * p.Ho * p.Wo * p.res_cstride *
                       (d->pixel_shuffle ? 4 : 1);
  p.res_planes = d->res_planes;
  p.bias = bias;
  if (d->planes_out == 4) {
    p.out = nullptr;
    p.out_f32 = static_cast<float*>(out);
  } else {
    p.out = static_cast<uint16_t*>(out);
    p.out_f32 = nullptr;
  }
  p.residual = static_cast<const uint16_t*>(residual);
  p.tile_sums = tile_sums;
  p.sums_per_cta = 0;
  p.pixel_shuffle = d->pixel_shuffle ? 1 : 0;
  p.res_bcast = d->res_bcast ? 1 : 0;
  p.act_after_res = d->act_after_res ? 1 : 0;
  if (p.pixel_shuffle && (d->Cout % 64 != 0 || d->planes_out == 4)) {
    delete pl;
    MTB_REQUIRE(false, "conv: pixel_shuffle needs Cout %% 64 == 0 and bf16 plane output");
  }
  // halo-tile kernel for the RCAN body layer (3x3, stride 1, 64 -> 64 channels); mode 1 forces the per-tap kernel
  pl->halo = (d->mode != 1 && conv_halo_eligible(p, d->Cin)) ? 1 : 0;
  if (d->mode == 2 && !pl->halo) {
    delete pl;
    MTB_REQUIRE(false, "conv: halo mode requested but layer is not eligible (needs 3x3 s1 p1, 64->64 channels)");
  }
  if (pl->halo) {
    p.TW = 8;
    p.TH = 16;
    // bf16x3 layers take the channel-major kernel (conv_halo_cm.cu) unless mode 3 pins the pixel-major one
    if (d->planes_in == 2 && d->mode != 3 && conv_halo_cm_eligible(p)) {
      pl->halo = 2;
      conv_halo_cm_tile(&p.TW, &p.TH);
    }
  }
  p.tiles_x = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_y = (p.Ho + p.TH - 1) / p.TH;
  {
    // channel tile (N of the MMA): a multiple-of-16 divisor of Cout.  Start from the largest one that still leaves
    // three pipeline stages (two-plane layers: 128), then shrink towards 64 while the layer has fewer tiles than SMs —
    // deep, spatially small layers are latency bound and gain more from extra CTAs and ring depth than from a wide N.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long m_tiles = static_cast<long long>(p.N) * p.tiles_y * p.tiles_x;
    const int cap = d->planes_in == 2 ? 128 : 256;
    int bn = 0;
    for (int c = cap; c >= 16; c -= 16)
      if (d->Cout % c == 0) {
        bn = c;
        break;
      }
    if (bn == 0)
      for (int c = 256; c > cap; c -= 16)     // no divisor below the cap (e.g. Cout = 176): take what divides
        if (d->Cout % c == 0) bn = c;
    while (bn > 64 && m_tiles * (d->Cout / bn) < sms) {
      int nb = 0;
      for (int c = bn - 16; c >= 64; c -= 16)
        if (d->Cout % c == 0) {
          nb = c;
          break;
        }
      if (nb == 0) break;
      bn = nb;
    }
    p.BN = bn;
    p.n_tiles_n = d->Cout / bn;
  }
  // global-average-pool partials: one row per (CTA, lane quarter) when a launch covers a single image
  p.sums_per_cta = (p.N == 1 && p.n_tiles_n == 1) ? 1 : 0;
  pl->nsplit = d->planes_in == 2 ? 3 : 1;
  int cols = 32;
  while (cols < 2 * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage = static_cast<size_t>(d->planes_in) * (128 * 128 + p.BN * 128);
  int stages = static_cast<int>((200 * 1024) / stage);
  if (stages > 8) stages = 8;
  p.num_stages = stages;
  if (stages < 2) {
    delete pl;
    MTB_REQUIRE(false, "conv: stage of %zu bytes leaves < 2 pipeline stages", stage);
  }

  // activations: [planes*N][H][W][Cin] bf16
  {
    const uint64_t xc = static_cast<uint64_t>(d->x_ctotal > 0 ? d->x_ctotal : d->Cin);
    const uint64_t dims[4] = {xc, static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N) * d->planes_in};
    const uint64_t strides[3] = {xc * 2, static_cast<uint64_t>(d->W) * xc * 2, static_cast<uint64_t>(d->H) * d->W * xc * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(p.TW * d->stride), static_cast<uint32_t>(p.TH * d->stride), 1};
    if (pl->halo) {
      box[1] = static_cast<uint32_t>(p.TW + 2);
      box[2] = static_cast<uint32_t>(p.TH + 2);
    }
    const uint32_t es[4] = {1, static_cast<uint32_t>(d->stride), static_cast<uint32_t>(d->stride), 1};
    if (encode_tmap(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es,
                    CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
      delete pl;
      return -3;
    }
  }
  memset(&pl->tmO, 0, sizeof(pl->tmO));
  if (pl->halo == 2) {
    // output planes [2*N][Ho][Wo][64] for the staged TMA store
    const uint64_t dims[4] = {64, static_cast<uint64_t>(p.Wo), static_cast<uint64_t>(p.Ho), static_cast<uint64_t>(p.N) * 2};
    const uint64_t strides[3] = {128, static_cast<uint64_t>(p.Wo) * 128, static_cast<uint64_t>(p.Ho) * p.Wo * 128};
    const uint32_t box[4] = {64, 8, 2, 1};
    if (encode_tmap(&pl->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
      delete pl;
      return -3;
    }
  }
  // weights: [planes*taps*Cout][Cin] bf16
  {
    const uint64_t rows = static_cast<uint64_t>(d->planes_in) * d->KH * d->KW * d->Cout;
    const uint64_t dims[2] = {static_cast<uint64_t>(d->Cin), rows};
    const uint64_t strides[1] = {static_cast<uint64_t>(d->Cin) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(pl->halo == 2 ? 16 : p.BN)};
    if (encode_tmap(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
      delete pl;
      return -3;
    }
  }
  *plan_out = pl;
  return 0;
}

/* RCAN body layer (3x3, stride 1, pad 1, 64 -> 64) in the fp16 + e5m2-correction format (conv_halo_fp16c.cu). */
int mtb_rcan_conv_plan_create(int N, int H, int W, const void* x, const void* w_packed, const float* bias, void* out,
                              const void* residual, float* tile_sums, int act, int lo_shift, mtb_conv_plan** plan_out) {
  MTB_REQUIRE(x && w_packed && out && plan_out, "mtb_rcan_conv_plan_create: null argument");
  MTB_REQUIRE(N > 0 && H > 0 && W > 0, "mtb_rcan_conv_plan_create: empty tensor");
  MTB_REQUIRE(lo_shift >= -14 && lo_shift <= 14, "mtb_rcan_conv_plan_create: lo_shift out of range");
  MTB_REQUIRE(static_cast<long long>(N) * 3 < 65536, "mtb_rcan_conv_plan_create: batch too large for the plane index");
  mtb_conv_plan* pl = new (std::nothrow) mtb_conv_plan();
  MTB_REQUIRE(pl != nullptr, "conv: out of host memory");
  ConvParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.H = p.Ho = H;
  p.W = p.Wo = W;
  p.KH = p.KW = 3;
  p.pad = 1;
  p.stride = 1;
  p.cin_chunks = 1;
  p.Cout = 64;
  p.BN = 64;
  p.n_tiles_n = 1;
  conv_halo_fp16c_tile(&p.TW, &p.TH);
  p.tiles_x = (W + p.TW - 1) / p.TW;
  p.tiles_y = (H + p.TH - 1) / p.TH;
  p.planes_out = 3;
  p.act = act;
  p.out_cstride = p.res_cstride = 64;
  p.out_plane_stride = p.res_plane_stride = static_cast<long long>(N) * H * W * 64;     /* bytes */
  p.res_planes = residual ? 3 : 0;
  p.bias = bias;
  p.out = static_cast<uint16_t*>(out);
  p.residual = static_cast<const uint16_t*>(residual);
  p.tile_sums = tile_sums;
  p.sums_per_cta = (N == 1) ? 1 : 0;
  p.lo_scale = ldexpf(1.0f, lo_shift);
  p.lo_inv_scale = ldexpf(1.0f, -lo_shift);
  pl->halo = 3;
  pl->nsplit = 1;
  {
    // activations: byte planes [3*N][H][W][64 B] (fp16 channels 0-31 | fp16 channels 32-63 | e5m2 channels 0-63)
    const uint64_t dims[4] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N) * 3};
    const uint64_t strides[3] = {64, static_cast<uint64_t>(W) * 64, static_cast<uint64_t>(H) * W * 64};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.TW + 2), static_cast<uint32_t>(p.TH + 2), 1};
    if (encode_tmap(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, x, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_64B) != 0) {
      delete pl;
      return -3;
    }
  }
  {
    // weights: [2880 rows][64 B], already in shared-memory order (planes.conv_weight_to_fp16c)
    const uint64_t dims[2] = {64, 2880};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, 64};
    if (encode_tmap(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, w_packed, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_64B) != 0) {
      delete pl;
      return -3;
    }
  }
  memset(&pl->tmO, 0, sizeof(pl->tmO));
  if (residual) {
    // the residual tile (no halo), for the producer's L2 prefetch
    const uint64_t dims[4] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N) * 3};
    const uint64_t strides[3] = {64, static_cast<uint64_t>(W) * 64, static_cast<uint64_t>(H) * W * 64};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.TW), static_cast<uint32_t>(p.TH), 1};
    if (encode_tmap(&pl->tmO, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, residual, dims, strides, box, nullptr,
                    CU_TENSOR_MAP_SWIZZLE_NONE) != 0) {
      delete pl;
      return -3;
    }
  } else {
    pl->tmO = pl->tmA;      /* never dereferenced without a residual; keeps the kernel parameter a valid descriptor */
  }
  *plan_out = pl;
  return 0;
}

int mtb_planes_bf16x2_to_fp16c(const void* in, long long npix, void* out, int lo_shift, void* stream) {
  MTB_REQUIRE(in && out && npix >= 0, "mtb_planes_bf16x2_to_fp16c: bad arguments");
  const int rc = launch_bf16x2_to_fp16c(in, npix, out, ldexpf(1.0f, lo_shift), static_cast<cudaStream_t>(stream));
  if (rc == 0) mtb::g_launches.fetch_add(1);
  return rc;
}

int mtb_planes_fp16c_to_bf16x2(const void* in, long long npix, void* out, int lo_shift, void* stream) {
  MTB_REQUIRE(in && out && npix >= 0, "mtb_planes_fp16c_to_bf16x2: bad arguments");
  const int rc = launch_fp16c_to_bf16x2(in, npix, out, ldexpf(1.0f, -lo_shift), static_cast<cudaStream_t>(stream));
  if (rc == 0) mtb::g_launches.fetch_add(1);
  return rc;
}

int mtb_conv_plan_run(mtb_conv_plan* plan, void* stream) {
  MTB_REQUIRE(plan != nullptr, "mtb_conv_plan_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = plan->halo == 3   ? launch_conv_halo_fp16c(plan->tmA, plan->tmB, plan->tmO, plan->p, st)
           : plan->halo == 2 ? launch_conv_halo_cm(plan->tmA, plan->tmB, plan->tmO, plan->p, st)
           : plan->halo == 1 ? launch_conv_halo(plan->tmA, plan->tmB, plan->p, plan->nsplit, st)
                             : launch_conv_gemm(plan->tmA, plan->tmB, plan->p, plan->nsplit, st);
  if (rc == 0) mtb::g_launches.fetch_add(1);
  return rc;
}

int mtb_conv_plan_set_channel_scale(mtb_conv_plan* plan, const float* scale) {
  MTB_REQUIRE(plan != nullptr, "mtb_conv_plan_set_channel_scale: null plan");
  MTB_REQUIRE(!plan->p.pixel_shuffle, "conv: channel scale with pixel_shuffle is not supported");
  plan->p.chan_scale = scale;
  return 0;
}

int mtb_conv_plan_set_border_sums(mtb_conv_plan* plan, float* border) {
  MTB_REQUIRE(plan != nullptr, "mtb_conv_plan_set_border_sums: null plan");
  MTB_REQUIRE((plan->halo == 2 || plan->halo == 3) && plan->p.sums_per_cta && plan->p.tile_sums != nullptr,
              "conv: border sums need the channel-major halo kernel with per-CTA tile sums (one image, bf16x3)");
  plan->p.border_sums = border;
  return 0;
}

int mtb_conv_plan_set_fixed_sums(mtb_conv_plan* plan, long long* fixed) {
  MTB_REQUIRE(plan != nullptr, "mtb_conv_plan_set_fixed_sums: null plan");
  MTB_REQUIRE(plan->halo == 3 && plan->p.sums_per_cta && plan->p.tile_sums != nullptr && plan->p.border_sums != nullptr,
              "conv: fixed-point sums need an fp16c plan (one image) with tile sums and border sums set");
  plan->p.sums_fixed = fixed;
  return 0;
}

int mtb_conv_plan_set_fused_gate(mtb_conv_plan* plan, const long long* fixed_in, long long* fixed_zero, const void* u,
                                 const float* conv_w, const float* conv_b, const float* w1, const float* b1,
                                 const float* w2, const float* b2, int R) {
  MTB_REQUIRE(plan != nullptr, "mtb_conv_plan_set_fused_gate: null plan");
  MTB_REQUIRE(plan->halo == 3 && plan->p.N == 1 && plan->p.residual != nullptr && plan->p.tile_sums == nullptr,
              "conv: a fused gate needs an fp16c plan (one image) with a residual and without sums (an RCAB's second conv)");
  MTB_REQUIRE(fixed_in && u && conv_w && w1 && w2 && R > 0 && R <= 16,
              "mtb_conv_plan_set_fused_gate: bad arguments (the fused gate handles up to 16 hidden units; use mtb_rcan_gate_fp16c)");
  ConvParams& p = plan->p;
  p.gate_fixed = fixed_in;
  p.gate_zero = fixed_zero;
  p.gate_u = static_cast<const uint8_t*>(u);
  p.gate_w = conv_w;
  p.gate_b = conv_b;
  p.gate_w1 = w1;
  p.gate_b1 = b1;
  p.gate_w2 = w2;
  p.gate_b2 = b2;
  p.gate_R = R;
  p.chan_scale = nullptr;
  return 0;
}

int mtb_conv_plan_num_mtiles(const mtb_conv_plan* plan) {
  return plan ? plan->p.N * plan->p.tiles_y * plan->p.tiles_x : 0;
}

int mtb_conv_plan_num_sum_rows(const mtb_conv_plan* plan) {
  if (!plan) return 0;
  const long long total = static_cast<long long>(plan->p.N) * plan->p.tiles_y * plan->p.tiles_x * plan->p.n_tiles_n;
  if (!plan->p.sums_per_cta) return static_cast<int>(static_cast<long long>(plan->p.N) * plan->p.tiles_y * plan->p.tiles_x * 4);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return static_cast<int>((total < sms ? total : sms) * 4);
}

void mtb_conv_plan_destroy(mtb_conv_plan* plan) { delete plan; }

}  // extern "C"
