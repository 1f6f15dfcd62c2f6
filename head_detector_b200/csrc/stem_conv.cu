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
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "aux_kernels.cuh"
#include "device_attr.cuh"

namespace vgh {

constexpr int kStemTile = 16;                 // output pixels per tile side
constexpr int kStemIn = 2 * kStemTile + 1;    // input rows / columns of a tile (stride 2, 3x3, pad 1)
constexpr int kStemRowWords = 26;             // staged input row: 99 bytes (33 px x 3) in 25 words, pitch 26 words
constexpr int kStemOutPitch = 72;             // staging row pitch in bf16 (64 + 8: conflict-free fragment stores)
constexpr int kStemWPitch = 40;               // weight row pitch in bf16 (32 + 8: conflict-free fragment reads)
constexpr int kStemFetch = (kStemIn + 7) / 8; // window rows per warp (8 warps)

template <bool F16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// two adjacent uint8 (low 16 bits of v) -> packed 16-bit floats, exact.  bf16: 0x4B000000 | b is the float 2^23 + b;
// fp16: 0x6400 | b is the half 1024 + b.
template <bool F16>
__device__ __forceinline__ uint32_t u8x2_to_16x2(uint32_t v) {
  if constexpr (F16) {
    const uint32_t biased = __byte_perm(v, 0x64u, 0x4140);   // bytes [b0, 0x64, b1, 0x64]
    const uint32_t k1024 = 0x64006400u;
    const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&biased), *reinterpret_cast<const __half2*>(&k1024));
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const float f0 = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540)) - 8388608.f;
    const float f1 = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7541)) - 8388608.f;
    const __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
    return *reinterpret_cast<const uint32_t*>(&p);
  }
}
__device__ __forceinline__ uint32_t pack_f2(float a, float b, bool f16) {   // keep in step with ptx.cuh pack16x2 (not included here)
  if (f16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// w: packed K-major weights [>= 48][32] bf16 / fp16 (F16; the pointer types say bf16 for both) (27 taps in (ky,kx,c) order + 5 zeros), bias [48] fp32.
// Persistent CTAs walk the (image, tile) list; the input window of tile i+1 is fetched into registers while tile i is
// being multiplied.  The kernel is instruction-issue bound if written naively (profiles/r2_ncu_full_stem_v1.txt: 9400
// warp instructions per tile, 67 % issue-active), so:
//   * the window is fetched as aligned 32-bit words by (warp = row, lane = word) - no divisions - and realigned by one
//     byte with a shuffle + funnel shift (the window starts 3 bytes before a 4-byte boundary);
//   * inside the kernel the 27 taps are ordered k' = ky*10 + (kx*3 + c) (weights permuted when they are staged), so that
//     every A-fragment register is two ADJACENT, 2-byte aligned window bytes: one LDS.U16 + 2 PRMT + 2 FADD + 1 pack;
//   * the copy-out walks rows with a constant address increment.
template <bool F16>
__global__ void __launch_bounds__(256, 4) stem_conv_kernel(const uint8_t* __restrict__ img, const __nv_bfloat16* __restrict__ w,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int S, int B,
                                                           int out_cstride, int relu) {
  __shared__ __align__(16) uint32_t tile[kStemIn * kStemRowWords];
  __shared__ __align__(16) __nv_bfloat16 stage[kStemTile * kStemTile * kStemOutPitch];
  __shared__ __align__(16) uint32_t wsm[48 * kStemWPitch / 2];
  __shared__ float bsm[48];
  const int Ho = S >> 1;
  const int tiles_x = Ho / kStemTile;
  const int tiles_per_img = tiles_x * tiles_x;
  const int n_tiles = tiles_per_img * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;

  // window of tile t -> registers: warp = rows warp, warp+8, ..; lane = word of the row (25 live); zero outside the image
  auto fetch = [&](int t, uint32_t (&v)[kStemFetch]) {
    const int b = t / tiles_per_img, r_ = t - b * tiles_per_img;
    const int ty = r_ / tiles_x, tx = r_ - ty * tiles_x;
    const int iy0 = 2 * ty * kStemTile - 1;
    // byte address of window column 0 is 3 bytes before a 4-byte boundary: fetch from one byte earlier
    const uint8_t* src = img + (static_cast<size_t>(b) * S * S + 2 * tx * kStemTile) * 3 - 4 + 4 * lane;
    const bool col_ok = t < n_tiles && lane < 25 && !(tx == 0 && lane == 0);
#pragma unroll
    for (int q = 0; q < kStemFetch; ++q) {
      const int r = warp + 8 * q, iy = iy0 + r;
      v[q] = 0u;
      if (col_ok && r < kStemIn && iy >= 0) v[q] = __ldg(reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(iy) * S * 3));
    }
  };

  uint32_t win[kStemFetch];
  fetch(blockIdx.x, win);

  for (int i = threadIdx.x; i < kStemTile * kStemTile; i += blockDim.x) {   // channels 48..63 of the padded output
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 48]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&stage[i * kStemOutPitch + 56]) = make_uint4(0, 0, 0, 0);
  }
  // weights: row n, tap k = ky*9 + r9 -> k' = ky*10 + r9 (zeros at k' = 9, 19, 29, 30, 31)
  {
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(wsm);
    for (int i = threadIdx.x; i < 48 * 32; i += blockDim.x) {
      const int n = i >> 5, kp = i & 31;
      const int ky = kp / 10, r9 = kp - ky * 10;
      wb[n * kStemWPitch + kp] = (kp < 30 && r9 < 9) ? w[n * 32 + ky * 9 + r9] : __float2bfloat16_rn(0.f);
    }
    if (threadIdx.x < 48) bsm[threadIdx.x] = __ldg(bias + threadIdx.x);
  }
  // byte offsets (inside the staged window, relative to the pixel's window origin) of this thread's four A-fragment
  // byte pairs: k' = 16 s + 2 t4 (+ 8); -1 = zero (k' >= 30)
  int koff[2][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kp = 16 * s + 2 * t4 + 8 * h;
      koff[s][h] = kp < 30 ? (kp / 10) * (kStemRowWords * 4) + (kp % 10) : -1;
    }
  const uint8_t* tile8 = reinterpret_cast<const uint8_t*>(tile);

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();   // every warp is done with the previous tile's window and staging tile
#pragma unroll
    for (int q = 0; q < kStemFetch; ++q) {
      const int r = warp + 8 * q;
      const uint32_t next = __shfl_down_sync(0xffffffffu, win[q], 1);
      if (r < kStemIn && lane < 25) tile[r * kStemRowWords + lane] = __funnelshift_r(win[q], lane < 24 ? next : 0u, 8);
    }
    __syncthreads();
    fetch(t + gridDim.x, win);   // in flight while this tile is multiplied and stored

#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int oy = 2 * warp + mi;                                                // one m-tile = one output row of the tile
      const uint8_t* r0 = tile8 + (2 * oy) * (kStemRowWords * 4) + (2 * g) * 3;    // window origin of pixel (oy, g)
      uint32_t a[2][4];
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // a0a1 (row g, k 2t..), a2a3 (row g+8, k 2t..), a4a5 (row g, k 2t+8..), a6a7 (row g+8, k 2t+8..)
          const int o = koff[s][h];
          a[s][2 * h] = o >= 0 ? u8x2_to_16x2<F16>(*reinterpret_cast<const uint16_t*>(r0 + o)) : 0u;
          a[s][2 * h + 1] = o >= 0 ? u8x2_to_16x2<F16>(*reinterpret_cast<const uint16_t*>(r0 + 48 + o)) : 0u;
        }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float2 bj = *reinterpret_cast<const float2*>(&bsm[8 * j + 2 * t4]);
        float d[4] = {bj.x, bj.y, bj.x, bj.y};
        const uint32_t* wr = wsm + (8 * j + g) * (kStemWPitch / 2) + t4;   // row n = 8j + g: k' = 2t, 2t+8 (+16 for the second step)
        mma_16816<F16>(d, a[0], wr[0], wr[4]);
        mma_16816<F16>(d, a[1], wr[8], wr[12]);
        if (relu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
        }
        *reinterpret_cast<uint32_t*>(&stage[(oy * kStemTile + g) * kStemOutPitch + 8 * j + 2 * t4]) = pack_f2(d[0], d[1], F16);
        *reinterpret_cast<uint32_t*>(&stage[(oy * kStemTile + g + 8) * kStemOutPitch + 8 * j + 2 * t4]) = pack_f2(d[2], d[3], F16);
      }
    }
    __syncthreads();
    // 256 pixels x 128 bytes, full lines: thread = (16-byte chunk c8, pixel px) of rows py0, py0 + 2, ...
    {
      const int b = t / tiles_per_img, r_ = t - b * tiles_per_img;
      const int ty = r_ / tiles_x, tx = r_ - ty * tiles_x;
      const int c8 = threadIdx.x & 7, px = (threadIdx.x >> 3) & 15, py0 = threadIdx.x >> 7;
      __nv_bfloat16* dst = out + ((static_cast<size_t>(b) * Ho + ty * kStemTile + py0) * Ho + tx * kStemTile + px) * out_cstride + 8 * c8;
      const __nv_bfloat16* sp = stage + (py0 * kStemTile + px) * kStemOutPitch + 8 * c8;
      const size_t dstep = static_cast<size_t>(2) * Ho * out_cstride;
#pragma unroll
      for (int q = 0; q < kStemTile / 2; ++q)
        *reinterpret_cast<uint4*>(dst + q * dstep) = *reinterpret_cast<const uint4*>(sp + q * 2 * kStemTile * kStemOutPitch);
    }
  }
}

int stem_conv_launch(const uint8_t* img, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int B, int S, int out_cstride,
                     int relu, int f16, cudaStream_t stream) {
  const int Ho = S / 2;
  if (Ho % kStemTile || out_cstride < 64 || out_cstride % 8 || (reinterpret_cast<uintptr_t>(img) & 3)) return 1;
  static std::atomic<int> carveout_set[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDevices && !carveout_set[dev].exchange(1)) {   // 4 CTAs x 44 KB static shared memory per SM
    cudaFuncSetAttribute(stem_conv_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(stem_conv_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  const int n_tiles = (Ho / kStemTile) * (Ho / kStemTile) * B;
  const int grid = n_tiles < 4 * device_sm_count() ? n_tiles : 4 * device_sm_count();
  if (f16) stem_conv_kernel<true><<<grid, 256, 0, stream>>>(img, w, bias, out, S, B, out_cstride, relu);
  else stem_conv_kernel<false><<<grid, 256, 0, stream>>>(img, w, bias, out, S, B, out_cstride, relu);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
