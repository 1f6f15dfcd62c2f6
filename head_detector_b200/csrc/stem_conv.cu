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
#include "device_attr.cuh"

namespace vgh {

constexpr int kStemTile = 16;                 // output pixels per tile side
constexpr int kStemIn = 2 * kStemTile + 1;    // input rows / columns of a tile (stride 2, 3x3, pad 1)
constexpr int kStemRow = 100;                 // bytes per staged input row (33 * 3 = 99, padded)
constexpr int kStemOutPitch = 72;             // staging row pitch in bf16 (64 + 8: conflict-free fragment stores)
constexpr int kStemWPitch = 40;               // weight row pitch in bf16 (32 + 8)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// w: packed K-major weights [>= 48][32] bf16 (27 taps in (ky,kx,c) order + 5 zeros), bias [48] fp32.
// Persistent CTAs walk the (image, tile) list; the input window of tile i+1 is fetched into registers while tile i is
// being multiplied, weights / bias fragments are loaded once per CTA.
constexpr int kStemLoads = (kStemIn * kStemRow + 255) / 256;

__global__ void __launch_bounds__(256, 4) stem_conv_kernel(const uint8_t* __restrict__ img, const __nv_bfloat16* __restrict__ w,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int S, int B,
                                                           int out_cstride, int relu) {
  __shared__ __align__(16) uint8_t tile[kStemIn * kStemRow];
  __shared__ __align__(16) __nv_bfloat16 stage[kStemTile * kStemTile * kStemOutPitch];
  __shared__ __align__(16) uint32_t wsm[48 * kStemWPitch / 2];
  __shared__ float bsm[48];
  const int Ho = S >> 1;
  const int tiles_x = Ho / kStemTile;
  const int tiles_per_img = tiles_x * tiles_x;
  const int n_tiles = tiles_per_img * B;

  auto fetch = [&](int t, uint8_t (&v)[kStemLoads]) {   // window of tile t -> registers (zero outside the image = conv padding)
    const int b = t / tiles_per_img, r_ = t - b * tiles_per_img;
    const int ty = r_ / tiles_x, tx = r_ - ty * tiles_x;
    const int iy0 = 2 * ty * kStemTile - 1, ix0 = 2 * tx * kStemTile - 1;
    const uint8_t* src = img + static_cast<size_t>(b) * S * S * 3;
#pragma unroll
    for (int q = 0; q < kStemLoads; ++q) {
      const int i = threadIdx.x + q * 256;
      const int r = i / kStemRow, c = i - r * kStemRow;
      const int iy = iy0 + r, ix = ix0 + c / 3;
      v[q] = 0;
      if (t < n_tiles && i < kStemIn * kStemRow && c < kStemIn * 3 && iy >= 0 && iy < S && ix >= 0 && ix < S)
        v[q] = __ldg(src + (static_cast<size_t>(iy) * S + ix0) * 3 + c);
    }
  };

  uint8_t win[kStemLoads];
  fetch(blockIdx.x, win);

  for (int i = threadIdx.x; i < kStemTile * kStemTile; i += blockDim.x) {   // channels 48..63 of the padded output
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 48]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 56]) = make_uint4(0, 0, 0, 0);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  // K offsets of this thread's A-fragment elements inside the staged window: k = ky*9 + kx*3 + c -> ky*row + (k % 9)
  int koff[2][4];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * s + 2 * t4 + (q & 1) + 8 * (q >> 1);
      koff[s][q] = k < 27 ? (k / 9) * kStemRow + (k % 9) : -1;
    }
  // weights / bias of the six 8-channel n-tiles live in shared memory (pitch 40: conflict-free fragment reads)
  for (int i = threadIdx.x; i < 48 * 16; i += blockDim.x)
    wsm[(i >> 4) * (kStemWPitch / 2) + (i & 15)] = __ldg(reinterpret_cast<const uint32_t*>(w) + i);
  if (threadIdx.x < 48) bsm[threadIdx.x] = __ldg(bias + threadIdx.x);

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();   // every warp is done with the previous tile's window and staging tile
#pragma unroll
    for (int q = 0; q < kStemLoads; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < kStemIn * kStemRow) tile[i] = win[q];
    }
    __syncthreads();
    fetch(t + gridDim.x, win);   // in flight while this tile is multiplied and stored

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
        const float2 bj = *reinterpret_cast<const float2*>(&bsm[8 * j + 2 * t4]);
        float d[4] = {bj.x, bj.y, bj.x, bj.y};
        const uint32_t* wr = wsm + (8 * j + g) * (kStemWPitch / 2) + t4;   // row n = 8j + g: k = 2t, 2t+8 (+16 for the second step)
        mma_bf16_16816(d, a[0], wr[0], wr[4]);
        mma_bf16_16816(d, a[1], wr[8], wr[12]);
        if (relu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]), hi = __floats2bfloat162_rn(d[2], d[3]);
        *reinterpret_cast<__nv_bfloat162*>(&stage[(oy * kStemTile + g) * kStemOutPitch + 8 * j + 2 * t4]) = lo;
        *reinterpret_cast<__nv_bfloat162*>(&stage[(oy * kStemTile + g + 8) * kStemOutPitch + 8 * j + 2 * t4]) = hi;
      }
    }
    __syncthreads();
    // 256 pixels x 128 bytes, full lines
    const int b = t / tiles_per_img, r_ = t - b * tiles_per_img;
    const int ty = r_ / tiles_x, tx = r_ - ty * tiles_x;
    const int oy0 = ty * kStemTile, ox0 = tx * kStemTile;
#pragma unroll
    for (int q = 0; q < kStemTile * kStemTile * 8 / 256; ++q) {
      const int i = threadIdx.x + q * 256;
      const int pix = i >> 3, c8 = i & 7;
      const int py = pix / kStemTile, px = pix - py * kStemTile;
      const uint4 v = *reinterpret_cast<const uint4*>(&stage[pix * kStemOutPitch + 8 * c8]);
      *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * Ho + oy0 + py) * Ho + ox0 + px) * out_cstride + 8 * c8) = v;
    }
  }
}

int stem_conv_launch(const uint8_t* img, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int B, int S, int out_cstride,
                     int relu, cudaStream_t stream) {
  const int Ho = S / 2;
  if (Ho % kStemTile || out_cstride < 64 || out_cstride % 8) return 1;
  static std::atomic<int> carveout_set[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDevices && !carveout_set[dev].exchange(1))   // 4 CTAs x 40 KB static shared memory per SM
    cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int n_tiles = (Ho / kStemTile) * (Ho / kStemTile) * B;
  const int grid = n_tiles < 4 * device_sm_count() ? n_tiles : 4 * device_sm_count();
  stem_conv_kernel<<<grid, 256, 0, stream>>>(img, w, bias, out, S, B, out_cstride, relu);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
