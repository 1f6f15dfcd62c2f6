// Mesh consumers on the device.
//
// PNCC: the reference paints every head with a CPU z-buffer rasteriser (Sim3DR, 3DDFA_V2's C++ core), one triangle after
// the other and one head after the other, each head starting from a fresh depth buffer and overwriting the heads before
// it wherever it is drawn (pncc_processor.py:66-73).  That order is turned into ONE 64-bit key per pixel,
//     key = (head + 1) << 52 | monotone(depth) << 20 | (0xFFFFF - triangle),
// resolved with atomicMax: a later head beats an earlier one, inside a head the larger depth wins, equal depths go to the
// lower triangle index - exactly what the sequential `p_depth > depth_buffer` test leaves behind.  A second pass recomputes
// the winner's barycentric weights and writes (unsigned char)(255 * colour).  All fp32 arithmetic is spelled with
// non-contracting intrinsics in the reference's operation order (get_point_weight, rasterize_kernel.cpp:54-83), so the
// image is bit-identical to the CPU rasteriser's.  (One theoretical difference: a pixel whose interpolated colour
// truncates to (0,0,0) is skipped by the reference's `sum != 0` composite and shows the head below; NCC colours are only
// black at the far corner of the template's bounding box, which no triangle reaches.)
#include "mesh_kernels.cuh"

#include <cmath>

namespace vgh {

struct Bary {
  float w0, w1, w2;
};

// get_point_weight (rasterize_kernel.cpp:54-83), operation for operation
__device__ __forceinline__ Bary point_weight(float px, float py, float x0, float y0, float x1, float y1, float x2, float y2) {
  const float v0x = __fsub_rn(x2, x0), v0y = __fsub_rn(y2, y0);
  const float v1x = __fsub_rn(x1, x0), v1y = __fsub_rn(y1, y0);
  const float v2x = __fsub_rn(px, x0), v2y = __fsub_rn(py, y0);
  const float dot00 = __fadd_rn(__fmul_rn(v0x, v0x), __fmul_rn(v0y, v0y));
  const float dot01 = __fadd_rn(__fmul_rn(v0x, v1x), __fmul_rn(v0y, v1y));
  const float dot02 = __fadd_rn(__fmul_rn(v0x, v2x), __fmul_rn(v0y, v2y));
  const float dot11 = __fadd_rn(__fmul_rn(v1x, v1x), __fmul_rn(v1y, v1y));
  const float dot12 = __fadd_rn(__fmul_rn(v1x, v2x), __fmul_rn(v1y, v2y));
  const float den = __fsub_rn(__fmul_rn(dot00, dot11), __fmul_rn(dot01, dot01));
  const float inv = den == 0.f ? 0.f : __fdiv_rn(1.f, den);
  const float u = __fmul_rn(__fsub_rn(__fmul_rn(dot11, dot02), __fmul_rn(dot01, dot12)), inv);
  const float v = __fmul_rn(__fsub_rn(__fmul_rn(dot00, dot12), __fmul_rn(dot01, dot02)), inv);
  Bary b;
  b.w0 = __fsub_rn(__fsub_rn(1.f, u), v);
  b.w1 = v;
  b.w2 = u;
  return b;
}

__device__ __forceinline__ uint32_t ordered_bits(float f) {  // monotone float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(128) pncc_raster_kernel(const float* __restrict__ verts, int n, int nverts, const int32_t* __restrict__ tris,
                                                          int ntri, int H, int W, unsigned long long* __restrict__ keys) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(n) * ntri) return;
  const int head = static_cast<int>(t / ntri), tri = static_cast<int>(t - static_cast<long long>(head) * ntri);
  const float* v = verts + static_cast<size_t>(head) * nverts * 3;
  const int i0 = tris[3 * tri], i1 = tris[3 * tri + 1], i2 = tris[3 * tri + 2];
  const float x0 = v[3 * i0], y0 = v[3 * i0 + 1], d0 = -v[3 * i0 + 2];
  const float x1 = v[3 * i1], y1 = v[3 * i1 + 1], d1 = -v[3 * i1 + 2];
  const float x2 = v[3 * i2], y2 = v[3 * i2 + 1], d2 = -v[3 * i2 + 2];
  const int x_min = max(static_cast<int>(ceilf(fminf(x0, fminf(x1, x2)))), 0);
  const int x_max = min(static_cast<int>(floorf(fmaxf(x0, fmaxf(x1, x2)))), W - 1);
  const int y_min = max(static_cast<int>(ceilf(fminf(y0, fminf(y1, y2)))), 0);
  const int y_max = min(static_cast<int>(floorf(fmaxf(y0, fmaxf(y1, y2)))), H - 1);
  if (x_max < x_min || y_max < y_min) return;
  const unsigned long long hi = static_cast<unsigned long long>(head + 1) << 52;
  const unsigned long long lo = static_cast<unsigned long long>(0xFFFFF - tri);
  for (int y = y_min; y <= y_max; ++y)
    for (int x = x_min; x <= x_max; ++x) {
      const Bary b = point_weight(static_cast<float>(x), static_cast<float>(y), x0, y0, x1, y1, x2, y2);
      if (!(b.w2 > 0.f && b.w1 > 0.f && b.w0 > 0.f)) continue;
      const float depth = __fadd_rn(__fadd_rn(__fmul_rn(b.w0, d0), __fmul_rn(b.w1, d1)), __fmul_rn(b.w2, d2));
      if (!(depth > -1e8f)) continue;  // the reference's depth buffer starts at -1e8
      atomicMax(keys + static_cast<size_t>(y) * W + x, hi | (static_cast<unsigned long long>(ordered_bits(depth)) << 20) | lo);
    }
}

__global__ void __launch_bounds__(256) pncc_resolve_kernel(const float* __restrict__ verts, int nverts, const int32_t* __restrict__ tris,
                                                           const float* __restrict__ colors, int H, int W,
                                                           const unsigned long long* __restrict__ keys, uint8_t* __restrict__ image) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const unsigned long long k = keys[p];
  if (k == 0ull) return;
  const int head = static_cast<int>(k >> 52) - 1, tri = 0xFFFFF - static_cast<int>(k & 0xFFFFFull);
  const float* v = verts + static_cast<size_t>(head) * nverts * 3;
  const int i0 = tris[3 * tri], i1 = tris[3 * tri + 1], i2 = tris[3 * tri + 2];
  const int y = p / W, x = p - y * W;
  const Bary b = point_weight(static_cast<float>(x), static_cast<float>(y), v[3 * i0], v[3 * i0 + 1], v[3 * i1], v[3 * i1 + 1], v[3 * i2], v[3 * i2 + 1]);
  uint8_t c[3];
  int sum = 0;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float pc = __fadd_rn(__fadd_rn(__fmul_rn(b.w0, colors[3 * i0 + ch]), __fmul_rn(b.w1, colors[3 * i1 + ch])), __fmul_rn(b.w2, colors[3 * i2 + ch]));
    // (unsigned char)((1 - alpha) * image + alpha * 255 * p_color) with alpha = 1
    c[ch] = static_cast<uint8_t>(static_cast<int>(__fadd_rn(0.f, __fmul_rn(255.f, pc))));
    sum += c[ch];
  }
  if (sum == 0) return;  // the reference composites only pixels whose colour is not black
  image[3 * p] = c[0];
  image[3 * p + 1] = c[1];
  image[3 * p + 2] = c[2];
}

int pncc_render_launch(const float* verts, int n, int nverts, const int32_t* tris, int ntri, const float* colors, int H, int W,
                       uint8_t* image, unsigned long long* keys, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n > 4094 || ntri > 0xFFFFF) return 2;
  if (cudaMemsetAsync(keys, 0, static_cast<size_t>(H) * W * 8, stream) != cudaSuccess) return 1;
  const long long work = static_cast<long long>(n) * ntri;
  pncc_raster_kernel<<<static_cast<int>((work + 127) / 128), 128, 0, stream>>>(verts, n, nverts, tris, ntri, H, W, keys);
  pncc_resolve_kernel<<<(H * W + 255) / 256, 256, 0, stream>>>(verts, nverts, tris, colors, H, W, keys, image);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// refined_head_bbox (utils.py:26-35): one warp per head
__global__ void __launch_bounds__(128) head_bbox_kernel(const float* __restrict__ verts, int n, int nverts, const int32_t* __restrict__ idx, int n_idx,
                                                        int32_t* __restrict__ out) {
  const int head = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (head >= n) return;
  const float* v = verts + static_cast<size_t>(head) * nverts * 3;
  float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
  for (int i = lane; i < n_idx; i += 32) {
    const float x = v[3 * idx[i]], y = v[3 * idx[i] + 1];
    x0 = fminf(x0, x); x1 = fmaxf(x1, x);
    y0 = fminf(y0, y); y1 = fmaxf(y1, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) {  // int() truncates towards zero
    const int ix = static_cast<int>(x0), iy = static_cast<int>(y0), ix1 = static_cast<int>(x1), iy1 = static_cast<int>(y1);
    out[4 * head] = ix; out[4 * head + 1] = iy; out[4 * head + 2] = ix1 - ix; out[4 * head + 3] = iy1 - iy;
  }
}

int head_bbox_launch(const float* verts, int n, int nverts, const int32_t* idx, int n_idx, int32_t* out_xywh, cudaStream_t stream) {
  if (n <= 0) return 0;
  head_bbox_kernel<<<(n + 3) / 4, 128, 0, stream>>>(verts, n, nverts, idx, n_idx, out_xywh);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
