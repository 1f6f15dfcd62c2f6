#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace vgh {
// boxes [B,A,4] xyxy fp32, scores [B,A] fp32 (>= 0).  keep_idx [B,keep_k] (anchor ids, -1 padded),
// keep_cnt [B]; keep_boxes [B,keep_k,4] / keep_scores [B,keep_k] optional.  top_k <= 1024.
int select_nms_launch(const float* boxes, const float* scores, int B, int A, float conf_thr, float iou_thr, int top_k,
                      int keep_k, int* keep_idx, int* keep_cnt, float* keep_boxes, float* keep_scores,
                      cudaStream_t stream, char* err, size_t errlen);
}  // namespace vgh
