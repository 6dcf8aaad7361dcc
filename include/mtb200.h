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
  int x_ctotal, x_coff;     /* input tensor channel count (0 = Cin) and first channel read (concat slices) */
  int out_ctotal, out_coff; /* output tensor channel count (0 = Cout) and first channel written */
  int res_ctotal, res_coff; /* same for the residual tensor */
  int res_bcast;      /* 1: residual has batch 1, shared by all N images */
  int act_after_res;  /* 1: activation after the residual add (default: before) */
  int pixel_shuffle;  /* 1: fuse PixelShuffle(2) into the store (Cout = 4 blocks [dy*2+dx] of Cout/4 channels) */
  int mode;           /* 0 auto, 1 force the per-tap kernel, 2 require the halo-tile kernel (3x3 s1 p1, 64->64) */
} mtb_conv_desc;

typedef struct mtb_conv_plan mtb_conv_plan;

/* x: bf16 [planes_in][N][H][W][x_ctotal]; w: bf16 [planes_in][KH*KW][Cout][Cin]; bias: fp32 [Cout] or NULL;
 * out: bf16 [planes_out][N][Ho][Wo][Cout] (or fp32 [N][Ho][Wo][Cout]); residual: like out (res_planes) or NULL;
 * tile_sums: fp32 [mtb_conv_plan_num_sum_rows()][Cout] or NULL (partial channel sums of the output). */
int mtb_conv_plan_create(const mtb_conv_desc* d, const void* x, const void* w, const float* bias, void* out,
                         const void* residual, float* tile_sums, mtb_conv_plan** plan);

/* RCAN body layer (3x3, stride 1, pad 1, 64 -> 64 channels) in the "fp16c" format: one fp16 product plus an e5m2
 * correction product per tap instead of the four bf16 products of the bf16x3 path (same fp32-grade result within the
 * 1e-3 bound, 25 % fewer tensor-core cycles, 3 bytes per activation).  Replaces the same spandrel RCAN conv stack
 * (reference core/image/image_utils.py:369-374, loader core/ml/model_manager.py:640-654).
 * x / out / residual: byte planes [3][N][H][W][64]: plane 0 = fp16 of channels 0-31, plane 1 = fp16 of channels 32-63,
 * plane 2 = e5m2 of (v - fp16(v)) * 2^lo_shift for channels 0-63.  w_packed: [2880][64] bytes in the kernel's
 * shared-memory order (mangatranslator_b200.planes.conv_weight_to_fp16c).  act: 0 none, 1 relu.  The plan is run,
 * configured (channel scale, border sums, sum rows) and destroyed with the mtb_conv_plan_* functions. */
int mtb_rcan_conv_plan_create(int N, int H, int W, const void* x, const void* w_packed, const float* bias, void* out,
                              const void* residual, float* tile_sums, int act, int lo_shift, mtb_conv_plan** plan_out);
/* bf16 hi/lo planes [2][npix][64] <-> fp16c planes [3][npix][64 B] at the boundary of the RCAN body */
int mtb_planes_bf16x2_to_fp16c(const void* in, long long npix, void* out, int lo_shift, void* stream);
int mtb_planes_fp16c_to_bf16x2(const void* in, long long npix, void* out, int lo_shift, void* stream);
int mtb_conv_plan_run(mtb_conv_plan* plan, void* stream);
/* optional per-output-channel factor applied to (acc + bias) before activation / residual: out = act((acc+b)*s) + res.
 * `scale` ([Cout] floats, device) is read at run time, so a kernel earlier in the stream may produce it. */
int mtb_conv_plan_set_channel_scale(mtb_conv_plan* plan, const float* scale);
/* optional second reduction of the same launch: per-channel sums of the output over image row 0, row Ho-1, column 0
 * and column Wo-1, as [num_sum_rows][4][Cout] floats (row split as tile_sums).  Only the channel-major halo kernel
 * with per-CTA sums provides it (returns an error otherwise); consumed by mtb_rcan_gate. */
int mtb_conv_plan_set_border_sums(mtb_conv_plan* plan, float* border);
int mtb_conv_plan_num_mtiles(const mtb_conv_plan* plan);
/* rows of the tile_sums buffer this plan writes ([rows][Cout] fp32): per (CTA, lane quarter) when the launch covers a
 * single image, else per (pixel tile, lane quarter) */
int mtb_conv_plan_num_sum_rows(const mtb_conv_plan* plan);
void mtb_conv_plan_destroy(mtb_conv_plan* plan);

/* ---- bubble cleaning (bit-exact integer path) ------------------------------------------------------------
 * Replaces the per-bubble OpenCV pipeline of core/image/cleaning.py:210-521 (process_single_bubble) and the
 * colour-grouped fill of core/image/cleaning.py:1021-1039, for all bubbles of all pages of a batch in one launch.
 */
#define MTB_CLEAN_MAX_SE 63
#define MTB_CLEAN_MAX_BALL 65
#define MTB_CLEAN_MAX_NEIGHBORS 16

typedef struct mtb_clean_params {
  int thr_value;   /* CleaningConfig.thresholding_value (core/config.py:28) */
  int use_otsu;    /* CleaningConfig.use_otsu_threshold */
  int retry_otsu;  /* retry failed bubbles once with Otsu (core/image/cleaning.py:690-734) */
  int kd, ke;      /* scale_kernel(DILATION/EROSION_KERNEL_SIZE, processing_scale) (cleaning.py:637-642) */
  int sed_hw[MTB_CLEAN_MAX_SE]; /* half-width of each row of cv2.getStructuringElement(MORPH_ELLIPSE,(kd,kd)); -1 = empty */
  int see_hw[MTB_CLEAN_MAX_SE];
  int ball_r;                               /* rows -ball_r..ball_r of the chamfer ball {N(dx,dy) < roi_shrink} */
  int ball_hw[2 * MTB_CLEAN_MAX_BALL + 1];
  int jball_r;                              /* same for JUNCTION_MIN_SHRINK (cleaning.py:38-39,167-168) */
  int jball_hw[2 * MTB_CLEAN_MAX_BALL + 1];
  int junction_margin;
  double min_area; /* scale_area(MIN_CONTOUR_AREA, ...) (cleaning.py:643-648) */
  int margin;      /* window margin applied by the host around each detection bbox */
} mtb_clean_params;

typedef struct mtb_clean_job {
  const uint8_t* img; /* device: page, interleaved BGR or BGRA */
  long long img_pitch;
  int img_h, img_w, img_c;
  const uint8_t* mask; /* device: mask bytes (>0 = set) of a rectangle placed at (mask_x0, mask_y0) */
  long long mask_pitch;
  int mask_x0, mask_y0, mask_w, mask_h;
  int wx0, wy0, cw, ch; /* crop window, inside the page */
  int bbox[4];          /* detection bbox (x0,y0,x1,y1) */
  int n_neighbors;      /* conjoined_neighbor_bboxes (core/image/detection.py:1197-1205) */
  int neighbors[MTB_CLEAN_MAX_NEIGHBORS][4];
  uint32_t* work;       /* device workspace of mtb_clean_workspace_words(cw, ch, max_runs) 32-bit words */
  int max_runs;
  int page_index;
} mtb_clean_job;

typedef struct mtb_clean_result {
  int status; /* 0 ok, 1 empty mask, 2 no valid contour, 3 window too small, 4 workspace overflow */
  int used_otsu, otsu_thr;
  int is_black;
  int fill_bgr[3];
  int text_bbox[4];
  int has_text_color;
  int text_color[4];
  int n_components, n_valid;
  int final_start;
  long long final_pixels;
  double final_area;
  unsigned long long gray_sum;
  unsigned int gray_cnt;
} mtb_clean_result;

unsigned long long mtb_clean_workspace_words(int cw, int ch, int max_runs);
/* jobs/results: device arrays of n_jobs entries; params: host pointer (copied). One CTA per job. */
int mtb_clean_bubbles(const mtb_clean_params* params, const mtb_clean_job* jobs_dev, mtb_clean_result* results_dev,
                      int n_jobs, void* stream);
/* paint every ok job's final mask with its fill colour into pages_out[page_index] (same geometry as the job's
 * page). Colour groups are applied in first-seen order per page, later groups overwrite earlier ones
 * (core/image/cleaning.py:1021-1039); max_ranks bounds the number of distinct colours per page (2 unless
 * coloured-bubble classification is on). The alpha channel of BGRA pages is preserved. */
int mtb_clean_paint(const mtb_clean_job* jobs_dev, const mtb_clean_result* results_dev, int n_jobs,
                    uint8_t* const* pages_out_dev, int* rank_scratch_dev /* n_jobs ints */, int max_ranks,
                    void* stream);
/* expand one job's final mask (bit-plane, crop window) into a full-frame uint8 {0,255} mask */
int mtb_clean_export_mask(const mtb_clean_job* jobs_dev, int job_index, uint8_t* out /* img_h x img_w */,
                          long long out_pitch, int plane /* 9 = final mask */, void* stream);

/* ---- HBM-bound glue around the convolutions ---------------------------------------------------------------
 * image_to_planes : core/image/image_utils.py:351-358 image_to_tensor (u8 -> float/255 NCHW), here u8 HWC -> NHWC planes
 * ca_scale        : RCAN CALayer (adaptive avg-pool finish + 1x1 -> ReLU -> 1x1 -> sigmoid), spandrel RCAN
 * scale_residual  : RCAB tail  y = x + t * gate[c]
 * f32_to_u8       : core/image/image_utils.py:361-366 tensor_to_image (clamp, *255, truncate)
 */
int mtb_image_to_planes(const uint8_t* img, int H, int W, int cimg, int swap_rb, float mul, const float* sub3 /* host */,
                        void* planes_out /* bf16 [planes][H][W][cpad] */, int cpad, int planes, void* stream);
int mtb_ca_scale(const float* sums, int n_images, int parts_per_image, int C, float inv_hw, const float* w1,
                 const float* b1, const float* w2, const float* b2, int R, float* scale_out, void* stream);
int mtb_scale_residual(const void* t, const void* x, const float* scale, void* y, long long pix_per_image, int N, int C,
                       int planes, void* stream);
/* rcan_gate: the CALayer gate of an RCAB computed from the INPUT of the block's second 3x3 conv (64 channels, one
 * image): sums = per-channel partial sums of u (conv1's tile_sums rows), u = conv1's output planes; conv_w/conv_b =
 * the second conv's fp32 weights [64][64][3][3] / bias; w1,b1,w2,b2 = conv_du.  By linearity of the zero-padded conv
 * mean(conv2(u)) follows from the total and the four border lines of u, so conv2 can apply the gate in its own
 * epilogue (mtb_conv_plan_set_channel_scale) and RCAB's `x + gate * t` needs no pass of its own. */
int mtb_rcan_gate(const float* sums, int parts, const float* border_sums /* [parts][4][64] or NULL: read u's lines */,
                  const void* u, int planes, int H, int W, const float* conv_w, const float* conv_b, const float* w1,
                  const float* b1, const float* w2, const float* b2, int R, float* scale_out, void* stream);
/* same for an RCAB whose body runs in the fp16c format (mtb_rcan_conv_plan_create): u = conv1's byte planes
 * [3][H][W][64], lo_shift as given to the plans; the border sums always come from conv1's epilogue. */
int mtb_rcan_gate_fp16c(const float* sums, int parts, const float* border_sums /* [parts][4][64] */,
                        long long* fixed_sums /* or: [5][64] from mtb_conv_plan_set_fixed_sums, consumed and zeroed */,
                        const void* u, int lo_shift, int H, int W, const float* conv_w, const float* conv_b,
                        const float* w1, const float* b1, const float* w2, const float* b2, int R, float* scale_out,
                        void* stream);
/* fp16c plans: accumulate the channel sums of the output (whole image, row 0, row Ho-1, column 0, column Wo-1) into
 * `fixed` ([5][64] int64, units of 2^-20, zero before the launch) with integer atomics instead of writing per-CTA
 * rows: the order of the additions cannot change the result, and the gate reads 2.5 KB instead of ~750 KB. */
int mtb_conv_plan_set_fixed_sums(mtb_conv_plan* plan, long long* fixed);
/* fp16c plans, an RCAB's second conv: compute the block's CALayer gate in this launch's prologue instead of a launch of
 * its own (mtb_rcan_gate_fp16c).  Every CTA derives the 64 gates from `fixed_in` (what the first conv accumulated through
 * mtb_conv_plan_set_fixed_sums) and the four corner pixels of `u` (this layer's input planes) while its weights stream
 * in, and applies them as the channel scale; CTA 0 zeroes `fixed_zero` (the accumulators the NEXT block's first conv
 * uses: alternate two buffers) — `fixed_in` itself is left as it is.  conv_w/conv_b: this layer's fp32 weights
 * [64][64][3][3] / bias; w1,b1,w2,b2,R: conv_du as in mtb_rcan_gate. */
int mtb_conv_plan_set_fused_gate(mtb_conv_plan* plan, const long long* fixed_in, long long* fixed_zero, const void* u,
                                 const float* conv_w, const float* conv_b, const float* w1, const float* b1,
                                 const float* w2, const float* b2, int R);
int mtb_f32_to_u8(const float* in, long long npix, int cpad, const float* add3 /* host */, float mul, uint8_t* out,
                  float* out_f /* optional float copy [npix][3] */, void* stream);
/* "_PU" RCAN variants (ModelManager.load_upscale_lite, core/ml/model_manager.py:660-700 -> spandrel RCAN with
 * unshuffle_mod): PixelUnshuffle(d) of the reflect-padded page folded into the u8 -> planes conversion
 * (planes_out: bf16 [planes][ceil(H/d)][ceil(W/d)][cpad], channel c*d*d + dy*d + dx), and the crop back to the
 * un-padded frame folded into the float -> u8 conversion. */
int mtb_image_to_planes_unshuffle(const uint8_t* img, int H, int W, int cimg, int swap_rb, float mul,
                                  const float* sub3 /* host */, int d, void* planes_out, int cpad, int planes, void* stream);
int mtb_f32_to_u8_crop(const float* in, int Hin, int Win, int cpad, int Hout, int Wout, const float* add3 /* host */,
                       float mul, uint8_t* out, float* out_f, void* stream);

/* ---- detection glue ---------------------------------------------------------------------------------------
 * maxpool / upsample2x : SPPF and FPN ops of the YOLO graph on channel slices of NHWC plane tensors
 * yolo_decode          : ultralytics Segment head decode (DFL + sigmoid) + confidence filter
 *                        (reference call: core/image/detection.py:1338-1345)
 * nms                  : ultralytics non_max_suppression + scale_boxes, then the reference's own
 *                        _deduplicate_primary_boxes / _remove_contained_boxes (core/image/detection.py:219-295)
 */
int mtb_maxpool(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int k,
                int planes, void* stream);
int mtb_upsample2x(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int planes,
                   void* stream);
/* Depthwise k x k convolution (groups = channels; stride 1, pad k/2) + bias (+ SiLU when act != 0) on a channel slice of an
 * NHWC plane tensor: ultralytics DWConv (YOLO11 class branch, detection.py:1817 panel model) and the positional convolution
 * of the C2PSA / A2C2f attention blocks (YOLO12x OSB-text model, detection.py:120-201).  w: device fp32 [k*k][c]. */
int mtb_dwconv(const void* x, void* y, int N, int H, int W, int ct_in, int ci, int ct_out, int co, int c, int k,
               const float* w, const float* bias, int act, int planes, void* stream);

typedef struct mtb_yolo_level {
  const float* box; /* device fp32 [N][H][W][64] DFL logits */
  const float* cls; /* device fp32 [N][H][W][ncp] class logits */
  int H, W, stride, reserved;
} mtb_yolo_level;

/* cand: [N][max_cand][6] (x1,y1,x2,y2 in letterboxed px, score, class); count: [N] */
int mtb_yolo_decode(const mtb_yolo_level* levels /* host */, int n_levels, int N, int nc, int ncp, float conf,
                    int max_cand, float* cand, int* cand_anchor, int* count, void* stream);

/* retina masks (ultralytics process_mask_native, detection.py:1338-1345 retina_masks=True): out u8 [n][H][W] in {0,1};
 * rows (optional) selects rows of det; (top,left,ch,cw) = prototype crop that strips the letterbox padding */
int mtb_yolo_masks(const float* proto, int mh, int mw, int nm, const float* const* mc_levels /* host array of 3 device ptrs */,
                   const int* level_hw /* host 3x2 */, const float* det, const int* rows, int n, int top, int left, int ch,
                   int cw, int H, int W, uint8_t* out, void* stream);

typedef struct mtb_nms_params {
  int N, max_cand, max_det;
  float iou_thr, max_wh;
  float gain; /* letterbox gain (scale_boxes) */
  int pad_x, pad_y, img_w, img_h;
  double dedup_iou, contain_ioa;
  int apply_dedup;
} mtb_nms_params;

/* out_det: [N][max_det][8] (x1,y1,x2,y2 original px, score, class, anchor, kept-by-reference-logic);
 * out_count: [N][2] (after NMS, after dedup+containment); final_idx: [N][max_det] rows of out_det kept, in the
 * reference's order */
int mtb_nms(const mtb_nms_params* p /* host */, const float* cand, const int* cand_anchor, const int* count,
            int* order_ws /* [N][max_cand] */, unsigned char* dead_ws /* [N][max_cand] */, float* out_det,
            int* out_count, int* final_idx, void* stream);

/* ---- conjoined-bubble mask splitting (core/image/detection.py:971-1035 `_split_conjoined_mask` with
 * `_seed_mask_from_box` :646-672, `_split_overlap_zone_with_line` :675-800, `_expand_resolved_masks_within_parent`
 * :932-968) --------------------------------------------------------------------------------------------------
 * parent: uint8 HxW (non-zero = inside), device.  rects [K][4] (x0,y0,x1,y1: floor/ceil clip of the child boxes,
 * `_build_rect_mask_from_box` :568-579), centers [K][2] (float box centres, for the empty-seed fallback), window
 * (x0,y0,x1,y1: must contain every parent pixel and every child rectangle) and pairs (one per overlapping (i<j) in
 * the reference's loop order) are HOST arrays prepared by the caller.  A pair's zone parent∧rect_i∧rect_j is cleared
 * from both children and re-divided by v = (x-cx)*ax + (y-cy)*ay in float64: mode 1 gives i the pixels with v <= 0
 * and j those with v > 0, mode 2 gives i v >= 0 and j v < 0, mode 0 leaves the zone to the nearest-child rule.
 * out: uint8 [K][H][W] {0,255}.  Bit-exact with the reference under OpenCV's own (non-IPP) distanceTransform. */
#define MTB_SPLIT_MAX_CHILDREN 15
typedef struct mtb_split_pair {
  int i, j, mode, reserved;
  double cx, cy, ax, ay;
  double off; /* v = (x - cx) * ax + (y - cy) * ay - off: the text-safe offset of the cut (detection.py:700-761), 0 otherwise */
} mtb_split_pair;
long long mtb_split_conjoined_workspace_bytes(int win_h, int win_w, int K);
int mtb_split_conjoined(const uint8_t* parent, int H, int W, int K, const int* rects, const double* centers,
                        const int* window, int n_pairs, const mtb_split_pair* pairs, uint8_t* out, void* workspace,
                        long long workspace_bytes, void* stream);

/* ---- RT-DETRv2 glue (core/ml/rtdetr_adapter.py:61-113 -> transformers RTDetrV2ForObjectDetection; secondary detector of
 * core/image/detection.py:1392-1548) --------------------------------------------------------------------------
 * maxpool2d   : ResNet stem MaxPool2d(k, stride, pad) on NHWC planes (padding never wins)
 * deform_attn : multi_scale_deformable_attention_v2 (method "default") for one image: value bf16 planes [tokens][ctotal]
 *               (levels concatenated; level l = H_l x W_l rows from row start_l), offsets fp32 [Q][off_stride] laid out
 *               (head, level*point, xy), logits fp32 [Q][logit_stride] laid out (head, level*point), ref fp32 [Q][4]
 *               (cx, cy, w, h in [0,1]); out bf16 planes [Q][ctotal].  Head dim must be 32. */
int mtb_maxpool2d(const void* x, void* y, int N, int H, int W, int C, int k, int stride, int pad, int planes, void* stream);
int mtb_deform_attn(const void* value, long long value_plane_stride, int planes, int ctotal, int heads, int hd, int n_levels,
                    const int* level_h_w_start /* host [n_levels][3] */, int n_points, const float* offsets, int off_stride,
                    const float* logits, int logit_stride, const float* ref_cxcywh, int Q, float offset_scale, void* out,
                    long long out_plane_stride, void* stream);

/* ---- input pre-processing, bit-exact with the CPU libraries the reference calls ----------------------------
 * letterbox_u8 : ultralytics LetterBox (cv2.resize INTER_LINEAR on uint8 + 114 border + BGR->RGB), detection.py:1338-1345
 * resize_aa_u8 : Sam2ImageProcessorFast resize (torchvision bilinear, antialias=True, uint8), detection.py:494-495
 */
int mtb_letterbox_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* dst /* dh x dw x 3 */, int dh, int dw, int top,
                     int left, int nh, int nw, int pad_value, int swap_rb, int* tables_dev /* >= 3*(nw+nh) ints */,
                     void* stream);
/* host-only: int16 antialias weight table of one axis (what resize_aa_u8 uploads); used by the CPU tests */
int mtb_aa_weights_host(int in_size, int out_size, int* start, int* len, short* weights, int kmax_cap, int* kmax,
                        int* prec);
int mtb_resize_aa_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* tmp /* sh x ow x 3 */,
                     uint8_t* dst /* oh x ow x 3 */, int oh, int ow, int* tables_dev, long long tables_ints, void* stream);
/* resize_lanczos_u8 : PIL `Image.resize((ow, oh), Image.LANCZOS)` on uint8 RGB — the exact-size resample after the RCAN
 * passes (core/image/image_utils.py:545) and resize_to_min_side / resize_to_max_side (:551-595) behind
 * process_bubble_image_cached (:678-746).  Pillow's separable resampler: double-precision windowed-sinc coefficients
 * rounded to 22 bits, uint8 intermediate after the horizontal pass.  `tables_dev`: caller-owned scratch of
 * mtb_resize_lanczos_table_ints(...) ints; `tmp`: sh x ow x 3 (may be NULL when only one axis changes). */
long long mtb_resize_lanczos_table_ints(int sh, int sw, int oh, int ow);
/* fills `tables_dev` for one geometry (host-side coefficient generation + one staged H2D copy) */
int mtb_resize_lanczos_tables(int sh, int sw, int oh, int ow, int* tables_dev, long long tables_ints, void* stream);
/* tables_ready = 1: `tables_dev` already holds this geometry's tables (callers keep one scratch per geometry) */
int mtb_resize_lanczos_u8(const uint8_t* src, int sh, int sw, int sc, uint8_t* tmp, uint8_t* dst /* oh x ow x 3 */,
                          int oh, int ow, int* tables_dev, long long tables_ints, int tables_ready, void* stream);
/* flatten_alpha_u8 : `background.paste(image, mask=alpha)` of convert_image_to_target_mode (image_utils.py:598-675) */
int mtb_flatten_alpha_u8(const uint8_t* src, int H, int W, const int* bg3 /* host */, uint8_t* dst, void* stream);
/* host-only: Pillow's 22-bit LANCZOS coefficient table of one axis (what resize_lanczos_u8 uploads); CPU tests */
int mtb_lanczos_weights_host(int in_size, int out_size, int* start, int* len, int* weights, int ksize_cap, int* ksize);

/* ---- safe text box of a cleaned bubble mask (bit-exact integer path) ----------------------------------------------
 * Replaces core/image/image_utils.py:173-348 calculate_centroid_expansion_box(cleaned_mask, padding_pixels), which the
 * reference's renderer runs per bubble on the masks clean_speech_bubbles returned (core/text/text_renderer.py:162):
 * exact Euclidean distance transform, safe area, centroid / pole of inaccessibility / nearest safe pixel, ray casts,
 * box.  One CTA per bubble, all bubbles in one launch; masks are read once by a bounds pass. */
typedef struct mtb_safebox_job {
  const uint8_t* mask; /* device, H x W uint8, nonzero = bubble interior (the reference passes 0 / 255) */
  long long pitch;     /* bytes between rows */
  int H, W;
  unsigned int t2;     /* mtb_safebox_threshold_sq(padding_pixels): safe <=> squared distance >= t2 */
  int cap;             /* capacity of g / safe in pixels; (mask bbox + 2)^2 must fit, (H+2)*(W+2) always does */
  uint16_t* g;         /* device workspace [cap] */
  uint8_t* safe;       /* device workspace [cap]; holds the window's safe-area map (0 / 255) afterwards */
} mtb_safebox_job;

typedef struct mtb_safebox_result {
  int status;       /* 0 ok; 1 empty mask ("Invalid or empty mask provided", :204-205); 2 no safe area, 3 no width or
                       height, 4 box leaves the image (all three: "Safe area calculation failed", :348); 5 workspace */
  int box[4];       /* x, y, width, height (:320) */
  int moved;        /* bit 0: anchor moved to the pole of inaccessibility (:247-253); bit 1: to the nearest safe pixel (:262-281) */
  int max_d2;       /* squared distance at the pole of inaccessibility */
  int anchor[2];    /* pixel the rays were cast from */
  int mask_bbox[4]; /* x0, y0, x1, y1 (inclusive) of the mask's nonzero pixels */
  int reserved;
  double cx, cy;    /* the centroid the reference returns (:255) */
} mtb_safebox_result;

/* host-only: smallest n with sqrtf((float)n) >= (float)padding_pixels, i.e. `distance_map >= padding_pixels` (:218) */
unsigned int mtb_safebox_threshold_sq(double padding_pixels);
/* jobs / results are device arrays; results need no initialisation.  Three stream operations: memset, bounds, boxes. */
int mtb_safe_boxes(const mtb_safebox_job* jobs_dev, mtb_safebox_result* results_dev, int n_jobs, void* stream);

/* ---- SAM 2.1 glue (transformers Sam2Model behind core/image/detection.py:475-511) -------------------------------- */
int mtb_layernorm(const void* x, long long rows, int C, int ct_in, int ci, int planes_in, const float* gamma,
                  const float* beta, float eps, void* y, int ct_out, int co, int planes_out, int gelu, void* stream);
int mtb_maxpool2x2(const void* x, void* y, int N, int H, int W, int C, int planes, void* stream);
int mtb_add_planes(const void* a, const void* b, void* out, long long rows, int C, long long b_rows, int planes,
                   void* stream);

typedef struct mtb_attn_desc {
  int B, heads, hd, nq, nk;
  float scale;
  const void *q, *k, *v; /* bf16 plane tensors, one row per token */
  void* out;
  int q_ct, q_off, k_ct, k_off, v_ct, v_off, o_ct, o_off; /* row width and first channel of q/k/v/out */
  long long q_ps, k_ps, v_ps, o_ps;                       /* plane strides (elements) */
  int planes;
  int mode; /* 0: token t of batch b is row b*n + t; 1: Hiera windows (ws x ws) on a grid_h x grid_w token grid,
               out-of-grid positions are padding whose q/k/v are pad_q/pad_k/pad_v (the qkv bias) */
  int grid_h, grid_w, ws, pool; /* pool: queries are 2x2 max-pooled inside each window (Hiera stage transition) */
  const float *pad_q, *pad_k, *pad_v;
  void* workspace;           /* optional device scratch: with >= mtb_attention_workspace_bytes() a mode-0 call with */
  long long workspace_bytes; /* two planes, head dim 64/96/128, nq >= 128, nk >= 256 runs on tcgen05 (attn_tc.cu)    */
} mtb_attn_desc;
int mtb_attention(const mtb_attn_desc* d /* host */, void* stream);
long long mtb_attention_workspace_bytes(int B, int heads, int hd, int nk);

/* Sam2PatchEmbeddings on the normalised image + positional embedding: u8 RGB HxWx3 -> planes [Ho][Wo][C] */
int mtb_sam_patch_embed(const uint8_t* img, int H, int W, const float* mean3, const float* std3, const float* w,
                        const float* b, const float* pos, int C, int k, int stride, int pad, void* out, int planes,
                        void* stream);
/* boxes in original pixels; sx = 1024/W, sy = 1024/H as the processor scales them */
int mtb_sam_prompt_boxes(const float* boxes, float sx, float sy, int P, const float* gauss, int half, const float* pe2, const float* pe3,
                         const float* not_a_point, float input_size, float* out, void* stream);
int mtb_sam_hyper_masks(const void* up, int planes, const float* hyper, int P, int K, int C, long long npix, float* out,
                        void* stream);
int mtb_sam_select_mask(const float* logits, const float* iou, int P, int K, long long npix, float delta, float thresh,
                        int* sel, void* stream);
/* masks[p] = (bilinear(logits[p][sel[p]], HxW) > 0) & rect(floor/ceil of boxes[p]) as uint8 {0,255}
 * (post_process_masks + core/image/detection.py:1732-1750); logit_out (optional) gets the interpolated logits */
int mtb_sam_mask_write(const float* logits, const int* sel, int K, int S, const float* boxes, int P, int H, int W,
                       uint8_t* masks, float* logit_out, void* stream);

/* ---- lossless PNG encoding of a finished page on the device --------------------------------------------------------
 * Replaces the host-side encoder of core/image/image_utils.py:59-170 save_image_with_compression (PIL's PNG writer, called
 * per page from core/pipeline.py:1996-2018) for the batch path: scanline filtering, one dynamic-Huffman deflate block per
 * 16 KB of filtered bytes (literals + distance-1 run matches), blocks byte-aligned with an empty stored block so they
 * concatenate.  The host builds the Huffman table from the histogram and wraps the stream in the PNG container.
 * Call order: mtb_png_filter -> mtb_png_histogram -> (host: table) -> mtb_png_deflate -> (prefix sum of sizes) ->
 * mtb_png_compact. */
#define MTB_PNG_SEGMENT 16384          /* filtered bytes per deflate block */
#define MTB_PNG_SEGMENT_STRIDE 32768   /* bytes reserved per encoded block in the staging buffer (worst case + header) */
/* img: uint8 [H][W][in_channels] (3 or 4) -> stream: [H][1 + W*out_channels] filter-type byte + filtered bytes; a fourth
 * output channel that the input lacks is opaque alpha (255).  Filter per row = libpng's minimum sum of absolute differences. */
int mtb_png_filter(const uint8_t* img, int H, int W, int in_channels, int out_channels, uint8_t* stream, void* cuda_stream);
/* hist: uint32 [288] (zero before the call) symbol counts of the literal/length alphabet over all blocks (incl. one
 * end-of-block per segment); adler_parts: uint64 [segments][2] = (sum of bytes, sum of (n - i) * byte_i) per segment. */
int mtb_png_histogram(const uint8_t* stream, long long total, unsigned int* hist, unsigned long long* adler_parts,
                      void* cuda_stream);
/* code: bit-reversed canonical Huffman codes, code_len: their lengths (0 = unused), header: the dynamic-block header with
 * BFINAL = 0 as little-endian 32-bit words (header_bits bits); staged: [segments][stride] bytes; sizes: bytes per block. */
int mtb_png_deflate(const uint8_t* stream, long long total, const unsigned short* code, const unsigned char* code_len,
                    const unsigned int* header, int header_bits, uint8_t* staged, int stride, unsigned int* sizes,
                    void* cuda_stream);
int mtb_png_compact(const uint8_t* staged, int stride, const unsigned int* sizes, const long long* offsets, int segments,
                    uint8_t* out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MTB200_H */
