/*
 * mtb200 — C ABI of the B200-native vision hot path of MangaTranslator
 * (detect -> segment -> clean -> upscale).
 *
 * The reference (meangrinch/MangaTranslator, Python) has no FFI of its own: its operator boundary is the set of
 * duck-typed objects returned by core/ml/model_manager.py and the stage functions of core/image
 * (SURVEY.md §8b).  The Python classes in mangatranslator_b200/core mirror those interfaces and call ONLY the
 * functions below (through ctypes).  Every entry point takes raw device pointers, sizes and a cudaStream_t
 * (passed as void*), allocates nothing behind the caller's back unless stated, and returns 0 on success;
 * on failure a message is available from mtb_last_error().
 *
 * Conventions
 *   - images are interleaved uint8 HxWxC in device memory (C = 3 or 4, BGR(A) like the reference's cv2 arrays);
 *   - masks are uint8 HxW with values {0,255} (reference: core/image/detection.py:1732-1750);
 *   - DNN activations are NHWC bf16 "planes" ([planes][N][H][W][C], C padded to a multiple of 64);
 *     two planes (hi, lo) carry an fp32-grade value, contracted with three bf16 tensor-core MMAs.
 */
#ifndef MTB200_H
#define MTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---- */
const char* mtb_last_error(void);
int mtb_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches counter) */
long long mtb_launch_count(void);

/* ---- conv / linear plans (tcgen05 implicit GEMM) -------------------------------------------------------
 * Replaces the cuDNN/cuBLAS calls made underneath:
 *   ultralytics Conv+SiLU blocks           (reference core/image/detection.py:1338-1345)
 *   spandrel RCAN conv stack               (reference core/image/image_utils.py:369-374)
 *   transformers Sam2Model linears / convs (reference core/image/detection.py:494-511)
 */
typedef struct mtb_conv_desc {
  int N, H, W, Cin; /* input geometry; Cin padded to a multiple of 64 */
  int Cout;         /* padded to a multiple of 16 */
  int KH, KW, stride, pad;
  int planes_in;  /* 1: bf16 GEMM, 2: hi/lo planes -> bf16x3 (fp32-grade) */
  int planes_out; /* 1 or 2 bf16 planes, or 4 = fp32 NHWC output */
  int act;        /* 0 none, 1 relu, 2 silu, 3 gelu(erf), 4 sigmoid */
  int res_planes; /* planes of the residual tensor (0 = none) */
  int tile_w, tile_h; /* pixel tile (product 128); 0 = choose */
} mtb_conv_desc;

typedef struct mtb_conv_plan mtb_conv_plan;

/* x: bf16 [planes_in][N][H][W][Cin]; w: bf16 [planes_in][KH*KW][Cout][Cin]; bias: fp32 [Cout] or NULL;
 * out: bf16 [planes_out][N][Ho][Wo][Cout] (or fp32 [N][Ho][Wo][Cout]); residual: like out (res_planes) or NULL;
 * tile_sums: fp32 [mtb_conv_plan_num_mtiles()*4][Cout] or NULL (per-warp channel sums of the output). */
int mtb_conv_plan_create(const mtb_conv_desc* d, const void* x, const void* w, const float* bias, void* out,
                         const void* residual, float* tile_sums, mtb_conv_plan** plan);
int mtb_conv_plan_run(mtb_conv_plan* plan, void* stream);
int mtb_conv_plan_num_mtiles(const mtb_conv_plan* plan);
void mtb_conv_plan_destroy(mtb_conv_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* MTB200_H */
