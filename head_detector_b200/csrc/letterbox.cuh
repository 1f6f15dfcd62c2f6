#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace vgh {
// n RGB uint8 images packed in src_dev (image i at byte offsets[i], rows of widths[i]*3 bytes) ->
// out_dev [n,S,S,3]; xform_host (optional) [n,3] = (pad_x, pad_y, scale) of detector.py:46-52.
int letterbox_launch(const uint8_t* src_dev, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n, int S,
                     uint8_t* out_dev, float* xform_host, cudaStream_t stream, char* err, size_t errlen);
}  // namespace vgh
