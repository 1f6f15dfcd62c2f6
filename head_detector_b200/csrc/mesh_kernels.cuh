#pragma once
// Consumers of the decoded meshes that stay on the device: PNCC rasteriser and refined head boxes.
#include <cuda_runtime.h>
#include <cstdint>

namespace vgh {
// PNCCProcessor.__call__ (head_detector/pncc_processor.py:66-73) + Sim3DR `_rasterize` (Sim3DR/lib/rasterize_kernel.cpp:219-293):
// verts [n,5023,3] (image-space x, y; depth = -z as the reference flips z before rasterising), tris [ntri,3], colors [5023,3] in [0,1].
// image [H,W,3] uint8 is painted in place (zero it first for the reference's result); keys = H*W uint64 workspace.
int pncc_render_launch(const float* verts, int n, int nverts, const int32_t* tris, int ntri, const float* colors, int H, int W,
                       uint8_t* image, unsigned long long* keys, cudaStream_t stream);
// refined_head_bbox (head_detector/utils.py:26-35): per head int-truncated min/max of x, y over idx -> (x, y, w, h)
int head_bbox_launch(const float* verts, int n, int nverts, const int32_t* idx, int n_idx, int32_t* out_xywh, cudaStream_t stream);
}  // namespace vgh
