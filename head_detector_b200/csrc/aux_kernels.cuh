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
};

int stem_pack_launch(const uint8_t* img, __nv_bfloat16* out, int B, int S, cudaStream_t stream);
int spp_pool_launch(__nv_bfloat16* buf, int B, int H, int W, int C, cudaStream_t stream);
int box_decode_launch(const DecodeLevels& lv, float* boxes, float* scores, int B, int A, cudaStream_t stream);
int flame_dense_launch(const DecodeLevels& lv, float* out, int B, int A, cudaStream_t stream);
int flame_gather_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, int B, int keep_k,
                        const float* img_xform, int* offsets, int* total, float* params, float* head_xform,
                        int* head_img, cudaStream_t stream);
int copy_rows_launch(const float* src, float* dst, const int* count_ptr, int row_floats, int max_rows, cudaStream_t stream);
}  // namespace vgh
