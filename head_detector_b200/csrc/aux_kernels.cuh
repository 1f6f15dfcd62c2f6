#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vgh {

// Per-level raw head outputs (fp32, pixel-major rows) and anchor bookkeeping (level-major anchors,
// row-major inside a level: yolo_head_ndfl_heads.py:214-231).
struct DecodeLevels {
  const float* reg[3];    // [B, hw, reg_cstride]: 68 DFL logits + 1 class logit (+pad)
  const float* flame[3];  // [B, hw, flame_cstride]: 205 raw flame channels (+pad)
  int a_off[4];           // anchor offset of each level, a_off[3] = A
  int W[3];
  int hw[3];
  float stride[3];
  int reg_cstride, flame_cstride;
  // sparse heads: flame[l] is the level's patch stack [patches*kPatch, kPatch, flame_cstride]; the row of head h
  // (packed image-major order) is the centre pixel of patch head_patch[h] of level head_level[h].  Null = dense maps.
  const int* head_level;
  const int* head_patch;
};

constexpr int kPatch = 8, kPatchC = 3;  // survivor patch size / offset of the anchor inside it (arch.py PATCH, PATCH_C)

// split != 0 (parity mode): the 32-wide row goes to planes 0, 2, 4 of a 192-wide granule (uint8 values are exact in bf16: m = l = 0)
int stem_pack_launch(const uint8_t* img, __nv_bfloat16* out, int B, int S, int split, cudaStream_t stream);
// fused stem (stem_conv.cu): uint8 image -> conv3x3 s2 (3 -> 48, K-major weights [>=48][32]) + bias + ReLU -> bf16 [B,S/2,S/2,out_cstride]
// f16 != 0: weights and output are IEEE fp16 instead of bf16 (same pointer types)
int stem_conv_launch(const uint8_t* img, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int B, int S, int out_cstride,
                     int relu, int f16, cudaStream_t stream);
int spp_pool_launch(__nv_bfloat16* buf, int B, int H, int W, int C, int f16, cudaStream_t stream);
// parity mode: the same pools on split activations (C logical channels = 6C physical per slice): max over y = h + m + l
int spp_pool_split_launch(__nv_bfloat16* buf, int B, int H, int W, int C, cudaStream_t stream);
int box_decode_launch(const DecodeLevels& lv, float* boxes, float* scores, int B, int A, cudaStream_t stream);
int flame_dense_launch(const DecodeLevels& lv, float* out, int B, int A, cudaStream_t stream);
int head_offsets_launch(const int* keep_cnt, int B, int* offsets, int* total, cudaStream_t stream);
int flame_gather_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, int B, int keep_k,
                        const float* img_xform, const int* offsets, float* params, float* head_xform,
                        int* head_img, cudaStream_t stream);
// sparse heads: survivors -> (level, patch) in image-major order; patch_src[l*cap + p] = image << 20 | y << 10 | x,
// level_rows[l] = patches of level l * kPatch (the live height of the level's patch stacks)
int patch_assign_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, const int* offsets, int B, int keep_k,
                        int cap, int* head_level, int* head_patch, int* patch_src, int* level_rows, cudaStream_t stream);
// feature map [B,H,W,C_total] channels [coff, coff+C) -> patch stack [cap*kPatch, kPatch, dst_C] channels [dst_coff, ..)
int patch_gather_launch(const __nv_bfloat16* feat, int H, int W, int C_total, int coff, int C, __nv_bfloat16* dst, int dst_C,
                        int dst_coff, const int* patch_src, const int* level_rows, int cap, cudaStream_t stream);
// zero the pixels of a patch stack that lie outside the H x W feature map (the dense graph's conv padding)
int patch_mask_launch(__nv_bfloat16* buf, int C_total, int coff, int C, int H, int W, const int* patch_src,
                      const int* level_rows, int cap, cudaStream_t stream);
int copy_rows_launch(const float* src, float* dst, const int* count_ptr, int row_floats, int max_rows, cudaStream_t stream);
}  // namespace vgh
