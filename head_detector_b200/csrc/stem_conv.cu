// Fused stem: uint8 image -> 3x3 stride-2 conv (3 -> 48, /255 folded into the weights) + bias + ReLU -> bf16 NHWC with the
// channel dimension padded to 64 (the K-block of the next layer), in ONE pass over HBM.
//
// Reference: YoloNASStem = QARepVGG(3 -> 48, stride 2) of the backbone (yolo_heads_l_arch_params.yaml:4-10) fed with
// `image / 255` (head_detector/detector.py:51), in deploy form.  The layer is 0.16 % of the network's MACs but its
// output is 13 MB per image: it is bound by the HBM write.  The first version expanded the image to a 32-wide im2col
// buffer (6.5 MB/image written and read back) and ran a Cin = 32 1x1 conv on tcgen05 - 2.2x the necessary traffic
// (profiles/r1_layer_rooflines.txt).  Here a CTA stages the 33x33x3-byte input window of a 16x16 output tile in shared
// memory, builds the im2col fragments in registers (uint8 -> bf16 is exact) and multiplies with warp-level
// mma.sync.m16n8k16 (K = 27 padded to 32, N = 48): 20 GFLOP per 64 images - a tcgen05 / TMEM pipeline would buy
// nothing for a kernel that waits on its 128-byte output rows, which leave through a padded staging tile as full lines.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "aux_kernels.cuh"

namespace vgh {

constexpr int kStemTile = 16;                 // output pixels per tile side
constexpr int kStemIn = 2 * kStemTile + 1;    // input rows / columns of a tile (stride 2, 3x3, pad 1)
constexpr int kStemRow = 100;                 // bytes per staged input row (33 * 3 = 99, padded)
constexpr int kStemOutPitch = 72;             // staging row pitch in bf16 (64 + 8: conflict-free fragment stores)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// w: packed K-major weights [>= 48][32] bf16 (27 taps in (ky,kx,c) order + 5 zeros), bias [48] fp32
__global__ void __launch_bounds__(256, 4) stem_conv_kernel(const uint8_t* __restrict__ img, const __nv_bfloat16* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int S,
                                                        int out_cstride, int relu) {
  __shared__ __align__(16) uint8_t tile[kStemIn * kStemRow];
  __shared__ __align__(16) __nv_bfloat16 stage[kStemTile * kStemTile * kStemOutPitch];
  const int Ho = S >> 1;
  const int tiles_x = Ho / kStemTile;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int b = blockIdx.y;
  const int oy0 = ty * kStemTile, ox0 = tx * kStemTile;
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
  const uint8_t* src = img + static_cast<size_t>(b) * S * S * 3;
  {
    // all loads of a thread are issued before the first one is consumed (one memory latency per CTA, not thirteen)
    constexpr int kLoads = (kStemIn * kStemRow + 255) / 256;
    uint8_t v[kLoads];
#pragma unroll
    for (int q = 0; q < kLoads; ++q) {
      const int i = threadIdx.x + q * 256;
      const int r = i / kStemRow, c = i - r * kStemRow;
      const int iy = iy0 + r, ix = ix0 + c / 3;
      v[q] = 0;
      if (i < kStemIn * kStemRow && c < kStemIn * 3 && iy >= 0 && iy < S && ix >= 0 && ix < S)
        v[q] = __ldg(src + (static_cast<size_t>(iy) * S + ix0) * 3 + c);
    }
#pragma unroll
    for (int q = 0; q < kLoads; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < kStemIn * kStemRow) tile[i] = v[q];
    }
  }
  for (int i = threadIdx.x; i < kStemTile * kStemTile; i += blockDim.x) {   // channels 48..63 of the padded output
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 48]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 56]) = make_uint4(0, 0, 0, 0);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // K offsets of this thread's A-fragment elements inside the staged window: k = ky*9 + kx*3 + c -> ky*row + (k % 9)
  int koff[2][4];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * s + 2 * t + (q & 1) + 8 * (q >> 1);
      koff[s][q] = k < 27 ? (k / 9) * kStemRow + (k % 9) : -1;
    }
  // B fragments (weights) and bias of the six 8-channel n-tiles
  uint32_t bw[6][2][2];
  float bs[6][2];
  const uint32_t* w32 = reinterpret_cast<const uint32_t*>(w);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int n = 8 * j + g;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      bw[j][s][0] = __ldg(w32 + (n * 32 + 16 * s + 2 * t) / 2);
      bw[j][s][1] = __ldg(w32 + (n * 32 + 16 * s + 2 * t + 8) / 2);
    }
    bs[j][0] = __ldg(bias + 8 * j + 2 * t);
    bs[j][1] = __ldg(bias + 8 * j + 2 * t + 1);
  }
  __syncthreads();

#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int oy = 2 * warp + mi;                                   // one m-tile = one output row of the tile
    const uint8_t* r0 = tile + (2 * oy) * kStemRow + (2 * g) * 3;   // window of pixel (oy, g)
    const uint8_t* r1 = r0 + 16 * 3;                                // ... of pixel (oy, g + 8)
    uint32_t a[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      float f[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = koff[s][q];
        f[q] = o >= 0 ? static_cast<float>(r0[o]) : 0.f;
        f[4 + q] = o >= 0 ? static_cast<float>(r1[o]) : 0.f;
      }
      // a0a1 (row g, k 2t..), a2a3 (row g+8, k 2t..), a4a5 (row g, k 2t+8..), a6a7 (row g+8, k 2t+8..)
      __nv_bfloat162 p;
      p = __floats2bfloat162_rn(f[0], f[1]); a[s][0] = *reinterpret_cast<uint32_t*>(&p);
      p = __floats2bfloat162_rn(f[4], f[5]); a[s][1] = *reinterpret_cast<uint32_t*>(&p);
      p = __floats2bfloat162_rn(f[2], f[3]); a[s][2] = *reinterpret_cast<uint32_t*>(&p);
      p = __floats2bfloat162_rn(f[6], f[7]); a[s][3] = *reinterpret_cast<uint32_t*>(&p);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float d[4] = {bs[j][0], bs[j][1], bs[j][0], bs[j][1]};
      mma_bf16_16816(d, a[0], bw[j][0][0], bw[j][0][1]);
      mma_bf16_16816(d, a[1], bw[j][1][0], bw[j][1][1]);
      if (relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
      }
      __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
      *reinterpret_cast<__nv_bfloat162*>(&stage[(oy * kStemTile + g) * kStemOutPitch + 8 * j + 2 * t]) = lo;
      *reinterpret_cast<__nv_bfloat162*>(&stage[(oy * kStemTile + g + 8) * kStemOutPitch + 8 * j + 2 * t]) = hi;
    }
  }
  __syncthreads();
  // 256 pixels x 128 bytes, full lines
#pragma unroll
  for (int q = 0; q < kStemTile * kStemTile * 8 / 256; ++q) {
    const int i = threadIdx.x + q * 256;
    const int pix = i >> 3, c8 = i & 7;
    const int py = pix / kStemTile, px = pix - py * kStemTile;
    const uint4 v = *reinterpret_cast<const uint4*>(&stage[pix * kStemOutPitch + 8 * c8]);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * Ho + oy0 + py) * Ho + ox0 + px) * out_cstride + 8 * c8) = v;
  }
}

int stem_conv_launch(const uint8_t* img, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int B, int S, int out_cstride,
                     int relu, cudaStream_t stream) {
  const int Ho = S / 2;
  if (Ho % kStemTile || out_cstride < 64 || out_cstride % 8) return 1;
  dim3 grid((Ho / kStemTile) * (Ho / kStemTile), B);
  stem_conv_kernel<<<grid, 256, 0, stream>>>(img, w, bias, out, S, out_cstride, relu);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
