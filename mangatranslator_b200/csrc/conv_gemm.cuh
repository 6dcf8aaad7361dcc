// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// One kernel serves every GEMM-shaped contraction of the hot path:
//   * YOLO Conv+SiLU blocks (3x3 / 1x1)                      (reference: core/image/detection.py:1338-1345 -> ultralytics)
//   * RCAN residual-channel-attention conv stack (3x3)       (reference: core/image/image_utils.py:369-374 -> spandrel RCAN)
//   * SAM 2.1 linears / 1x1 convs (as 1x1 "convs" over tokens) (reference: core/image/detection.py:475-511 -> transformers Sam2Model)
//
// Layout: activations are NHWC bf16 "planes".  A value that must keep fp32-grade accuracy is carried as two
// planes (hi = bf16_rn(v), lo = bf16_rn(v - hi)); the contraction then issues three bf16 MMAs
// (Ahi*Bhi + Ahi*Blo + Alo*Bhi) into one fp32 TMEM accumulator ("bf16x3", ~2^-16 relative error).
// With one plane it is a plain bf16 GEMM.
//
// CTA = 192 threads, persistent over output tiles:
//   warp 0     : TMA producer   (one lane)  global -> 128B-swizzled shared tiles, per (tap, 64-channel chunk)
//   warp 1     : MMA issuer     (one lane)  tcgen05.mma kind::f16, M=128 pixels x N=BN channels, K=16 per instr
//   warps 2..5 : epilogue       tcgen05.ld TMEM -> regs -> bias/activation/residual -> bf16 planes (or fp32) -> HBM
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace mtb {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_GELU = 3, ACT_SIGMOID = 4 };

struct ConvParams {
  int N, H, W, Ho, Wo;
  int KH, KW, pad, stride;
  int cin_chunks;   // padded Cin / 64
  int Cout;         // padded Cout (multiple of 16)
  int BN;           // channel tile (N of the MMA), multiple of 16, <= 256
  int n_tiles_n;    // Cout / BN
  int TW, TH;       // pixel tile: TW * TH == 128
  int tiles_x, tiles_y;
  int planes_out;   // 1 (bf16), 2 (bf16 hi/lo); ignored when out_f32 != nullptr
  int act;
  int num_stages;
  int tmem_cols;    // power of two >= 2*BN
  long long out_plane_stride;  // elements between the hi and lo planes of `out`
  long long res_plane_stride;  // elements between the hi and lo planes of `residual`
  int res_planes;              // 0, 1 or 2
  const float* bias;           // [Cout] or nullptr
  const float* chan_scale;     // [Cout] or nullptr: out = act((acc + bias) * chan_scale) (+ residual)
  uint16_t* out;               // bf16 planes [planes_out][N][Ho][Wo][Cout]
  float* out_f32;              // optional fp32 output [N][Ho][Wo][Cout]
  const uint16_t* residual;    // optional, same geometry as out, added after the activation
  float* tile_sums;            // optional per-warp channel sums of the output (global-average-pool partials):
                               //   sums_per_cta == 0: [N*tiles_y*tiles_x][4][Cout], one row per (tile, lane quarter)
                               //   sums_per_cta == 1: [gridDim.x][4][Cout], accumulated over the CTA's tiles (N == 1)
  int sums_per_cta;
  float* border_sums;          // optional (channel-major halo kernel, sums_per_cta): [gridDim.x][4][4][Cout] sums of the
                               //   output over image row 0, row Ho-1, column 0, column Wo-1 (same row split as tile_sums)
  int in_coff;                 // first input channel inside the (wider) input tensor
  int out_cstride, out_coff;   // channel count of the output tensor and first channel written (concat slices)
  int res_cstride, res_coff;   // same for the residual tensor
  int res_bcast;               // 1: the residual has batch 1 and is shared by all N images
  int act_after_res;           // 1: activation is applied after the residual add
  int debug;                   // perf experiments only (halo kernel): 1 skip stores, 2 skip MMA issue, 4 skip TMA loads
  long long* dbg_out;          // perf experiments only: per-CTA counters (halo kernels, debug bit 16)
  int pixel_shuffle;           // 1: Cout = 4 blocks of Cout/4 channels, block (dy*2+dx) is stored at pixel (2y+dy, 2x+dx)
  long long* sums_fixed;       // fp16c kernel, optional: [5][64] fixed-point (2^-20) accumulators of the output's channel sums
                               //   over the image, row 0, row Ho-1, column 0, column Wo-1, added with integer atomics
                               //   (order-independent, so bit-reproducible); replaces tile_sums / border_sums rows
  // fp16c kernel, optional: the RCAB gate computed in THIS launch's prologue (the layer is the block's second conv).  Every
  // CTA derives the 64 gates redundantly from `gate_fixed` (the [5][64] fixed-point sums of its input that the first conv
  // accumulated) while its weights stream in, and uses them as the channel scale; CTA 0 zeroes `gate_zero` (the other
  // block parity's accumulators) for the next block's first conv.
  const long long* gate_fixed;
  long long* gate_zero;
  const uint8_t* gate_u;       // this layer's input planes (fp16c), for the four corner pixels
  const float *gate_w, *gate_b, *gate_w1, *gate_b1, *gate_w2, *gate_b2;   // conv weights fp32 [64][64][3][3] / bias, conv_du
  int gate_R;
  int pf_x, pf_res;            // fp16c kernel: L2 prefetch distances (tiles ahead) of the activation / residual tiles
  float lo_scale, lo_inv_scale;  // fp16c kernel (conv_halo_fp16c.cu): the e5m2 residual plane holds (v - fp16(v)) * lo_scale;
                               //   there `out` / `residual` are byte tensors [3][N][H][W][64 B] and the plane strides are bytes
};

int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int nsplit,
                     cudaStream_t stream);
size_t conv_gemm_smem_bytes(const ConvParams& p, int nsplit);

}  // namespace mtb
