// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 activations.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel shifted by tap, cin] * W[cout, tap, cin]   (+bias, ReLU, +alpha*res)
//
// One CTA = one 128-pixel x BLOCK_N output tile.  The 128 "GEMM rows" are a th x tw rectangle of
// output pixels of one image; for every filter tap the matching input rectangle is fetched by ONE
// 4-D TMA box (channels innermost, 128B/64B swizzle) whose out-of-bounds part is zero-filled by the
// TMA unit - that is the convolution padding, and stride-2 layers use the tensor map's traversal
// stride.  No im2col buffer exists anywhere.  Weights are pre-packed K-major [Cout_pad][taps*Cin]
// and fetched by a 2-D TMA box.  Warp roles: warp0 = TMA producer, warp1 = TMEM owner + MMA issuer
// (single thread), warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU/residual -> global stores into
// a channel slice of the consumer's buffer, optionally pixel-shuffled for the 2x2 conv-transpose).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vgh {

struct ConvLaunch {
  CUtensorMap tmA;  // input  [C_total, W, H, B] bf16 (innermost first)
  CUtensorMap tmB;  // weights [K_total, N_pad]  bf16
  CUtensorMap tmOut;  // swapped kernel: output  [C_total, W, H, B] bf16, box [n_total, tw, th, 1] (TMA store)
  CUtensorMap tmRes;  // swapped kernel: residual [C_total, W, H, B] bf16, same box (TMA load)
  const float* bias;             // [N_pad]
  void* out;                     // output buffer base (bf16 or fp32)
  const __nv_bfloat16* res;      // residual buffer base or nullptr
  const int* dyn_rows;           // swapped kernel: device-side live height of a stacked-patch input (rows <= Ho), or nullptr
  int B, Ho, Wo;                 // GEMM-row space: output pixels (before any pixel shuffle)
  int tw, th, tiles_x, tiles_y;  // spatial tile (tw*th <= 128) and tile counts per image
  int cin_off, cin;              // channel coordinate offset / channels per tap (multiple of BK)
  int ntaps, kw, stride, pad;    // filter taps (kh*kw), kw, conv stride, padding
  int n_total, block_n;          // stored output channels (multiple of 16), UMMA N of this launch
  int out_cstride, out_coff, out_H, out_W;
  int up, up_cout;               // up=1: conv-transpose 2x2/s2, n-tile -> sub-pixel (n / up_cout)
  int relu, out_fp32;
  int res_cstride, res_coff;
  float res_alpha;
  int stages, tmem_cols;
  int mt;                        // M tiles (128-pixel accumulators) per work item sharing one weight stream
  int swap;                      // 1: operands swapped (conv_igemm_swap.cu): M = 128 cout rows, N = tw*th pixels
  int tail_rows, tail_imgs, n_tail_tiles;  // normal kernel: left-over rows of tail_imgs images share one tile
  int ks;                        // swapped kernel: k-blocks per pipeline stage (must divide taps*cin/BK)
  int ngroups, gw;               // swapped kernel: output-channel groups and channels per group (<= 128)
  int stg_bufs;                  // swapped kernel: epilogue staging tiles (2 = store of item i overlaps item i+1)
  int xr, xslots;                // swapped kernel, 3x3 stride 1: pixel tile + halo rows fetched once per column shift
                                 // and reused by the three row taps (xslots = pixel-tile ring depth; `stages` = weight ring)
  int cluster;                   // swapped kernel: 1, or 2 = CTA pairs share every weight k-block (each fetches half, TMA multicast)
  int pair;                      // swapped tap-reuse kernel: 1 = CTA pairs run one cta_group::2 MMA (M = 256 = two channel groups) per pixel tile
  int acc_stages, n_tiles, num_items;  // TMEM accumulator stages (1|2), N tiles, work items (persistent CTAs)
  int split;                     // normal kernel, parity mode: store y as bf16 terms h|m|l in six planes per 32-channel granule
  int f16;                       // 16-bit activations and weights are IEEE fp16 instead of bf16 (MMA operand format + epilogue conversions)
};

// Host side (conv_igemm.cu)
int conv_make_tensor_maps(ConvLaunch& L, const void* in_base, int in_C, int in_H, int in_W, const void* w_base,
                          int k_total, int n_pad, int bk);
// swapped kernel only: TMA maps of the output slice and (optional) residual slice
int conv_make_io_maps(ConvLaunch& L, void* out_base, const void* res_base);
int conv_launch(const ConvLaunch& L, int bk, cudaStream_t stream);
size_t conv_smem_bytes(const ConvLaunch& L, int bk);
int conv_pick_stages(int block_n, int bk, int mt);
int conv_default_mt(int block_n);
void conv_finalize(ConvLaunch& L);
bool conv_pdl_enabled();
size_t conv_swap_smem_bytes(const ConvLaunch& L, int bk);
int conv_swap_launch(const ConvLaunch& L, int bk, int sms, cudaStream_t stream, char* err, size_t errlen);
const char* conv_last_error();

}  // namespace vgh
