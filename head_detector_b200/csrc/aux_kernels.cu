// Small CUDA-core kernels around the tensor-core convs (sm_100a): stem conv on uint8 input, SPP
// max-pools, DFL/sigmoid box decode, FLAME-row assembly (dense or for NMS survivors only).
// All HBM-bound elementwise/stencil work: coalesced 16-byte accesses, no reshaping into GEMMs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "aux_kernels.cuh"

namespace vgh {

// ---------------------------------------------------------------------------------------- stem
// 3x3 stride-2 pad-1 conv, 3 -> 48 channels, ReLU; input uint8 NHWC (the /255 of detector.py:51 is
// folded into the fp32 weights), output bf16 NHWC with 64 channels (48 real + 16 zeros so that the
// next layer's K blocks are 64 wide).  thread = one output pixel, all 48 channels in 3 passes of 16.
__global__ void __launch_bounds__(128) stem_conv_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                        int B, int S) {
  // weights transposed to [tap][48] so that 4 output channels come from one broadcast 128-bit load
  __shared__ __align__(16) float ws[27 * 48];
  __shared__ __align__(16) float bs[48];
  for (int i = threadIdx.x; i < 48 * 27; i += blockDim.x) {
    const int co = i / 27, t = i - co * 27;
    ws[t * 48 + co] = w[i];
  }
  for (int i = threadIdx.x; i < 48; i += blockDim.x) bs[i] = bias[i];
  __syncthreads();
  const int Ho = S >> 1;
  const long long total_pix = static_cast<long long>(B) * Ho * Ho;
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= total_pix) return;
  const int ow = static_cast<int>(pix % Ho);
  const int oh = static_cast<int>((pix / Ho) % Ho);
  const int b = static_cast<int>(pix / (static_cast<long long>(Ho) * Ho));
  float x[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ih = 2 * oh + ky - 1;
    const bool row_ok = ih >= 0 && ih < S;
    const uint8_t* rowp = img + (static_cast<size_t>(b) * S + (row_ok ? ih : 0)) * S * 3;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iw = 2 * ow + kx - 1;
      const bool ok = row_ok && iw >= 0 && iw < S;
      const uint8_t* p = rowp + (ok ? iw : 0) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) x[(ky * 3 + kx) * 3 + c] = ok ? static_cast<float>(__ldg(p + c)) : 0.f;
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll 1
  for (int g = 0; g < 3; ++g) {  // 16 output channels per pass
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = bs[g * 16 + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4* wp = reinterpret_cast<const float4*>(ws + t * 48 + g * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = wp[q];
        acc[4 * q] = fmaf(w4.x, x[t], acc[4 * q]);
        acc[4 * q + 1] = fmaf(w4.y, x[t], acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w4.z, x[t], acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w4.w, x[t], acc[4 * q + 3]);
      }
    }
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(acc[2 * j], 0.f), fmaxf(acc[2 * j + 1], 0.f));
      packed[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    op[2 * g] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    op[2 * g + 1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
  }
  op[6] = make_uint4(0, 0, 0, 0);  // channels 48..63: zero padding (next layer's K blocks are 64 wide)
  op[7] = make_uint4(0, 0, 0, 0);
}

int stem_conv_launch(const uint8_t* img, const float* w, const float* bias, __nv_bfloat16* out, int B, int S,
                     cudaStream_t stream) {
  const long long Ho = S / 2;
  const long long total = static_cast<long long>(B) * Ho * Ho;
  stem_conv_kernel<<<static_cast<int>((total + 127) / 128), 128, 0, stream>>>(img, w, bias, out, B, S);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- SPP
// max-pool k=5,9,13 stride 1 (pad k/2 with -inf) of channel slice [0,C) of a [B,H,W,4C] buffer into
// slices 1,2,3.
__device__ __forceinline__ void max8(uint4& m, const uint4 v) {
  __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&m);
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], b[i]);
}

// CTA = one image x CH channels; the H x W x CH slab lives in shared memory and the three pools are
// computed as cascaded separable 5-wide maxima (mp9 = mp5(mp5), mp13 = mp5(mp9); -inf padding).
template <bool kRows>
__device__ __forceinline__ void max5_pass(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int L) {
  for (int i = threadIdx.x; i < H * W * L; i += blockDim.x) {
    const int l = i % L, pix = i / L;
    const int x = pix % W, y = pix / W;
    uint4 m = src[i];
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
      if (d == 0) continue;
      const int xx = kRows ? x + d : x, yy = kRows ? y : y + d;
      if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
      max8(m, src[(yy * W + xx) * L + l]);
    }
    dst[i] = m;
  }
}

__global__ void __launch_bounds__(256) spp_pool_kernel(__nv_bfloat16* __restrict__ buf, int B, int H, int W, int C, int CH) {
  extern __shared__ __align__(16) uint8_t spp_smem[];
  const int L = CH / 8;  // uint4 lanes per pixel
  uint4* b0 = reinterpret_cast<uint4*>(spp_smem);
  uint4* b1 = b0 + H * W * L;
  uint4* b2 = b1 + H * W * L;
  const int chunks = C / CH;
  const int b = blockIdx.x / chunks, c0 = (blockIdx.x - b * chunks) * CH;
  const int CT = 4 * C;
  __nv_bfloat16* base = buf + static_cast<size_t>(b) * H * W * CT + c0;
  for (int i = threadIdx.x; i < H * W * L; i += blockDim.x)
    b0[i] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(i / L) * CT + (i % L) * 8);
  __syncthreads();
  uint4* cur = b0;
  uint4* out = b2;
  for (int k = 1; k <= 3; ++k) {
    max5_pass<true>(cur, b1, H, W, L);
    __syncthreads();
    max5_pass<false>(b1, out, H, W, L);
    __syncthreads();
    for (int i = threadIdx.x; i < H * W * L; i += blockDim.x)
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(i / L) * CT + k * C + (i % L) * 8) = out[i];
    uint4* t = cur;
    cur = out;
    out = t;
  }
}

int spp_pool_launch(__nv_bfloat16* buf, int B, int H, int W, int C, cudaStream_t stream) {
  int CH = 32;
  while (CH > 8 && static_cast<size_t>(H) * W * CH * 2 * 3 > 160 * 1024) CH >>= 1;
  if (C % CH) return 1;
  const size_t smem = static_cast<size_t>(H) * W * CH * 2 * 3;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(spp_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    configured = smem;
  }
  spp_pool_kernel<<<B * (C / CH), 256, smem, stream>>>(buf, B, H, W, C, CH);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- box decode
// yolo_head_ndfl_heads.py:143-144,164-165: DFL softmax expectation, sigmoid, distance2bbox * stride.
__global__ void __launch_bounds__(256) box_decode_kernel(const DecodeLevels lv, float* __restrict__ boxes,
                                                         float* __restrict__ scores, int B, int A) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * A) return;
  const int a = static_cast<int>(idx % A);
  const int b = static_cast<int>(idx / A);
  int l = 0;
  while (l < 2 && a >= lv.a_off[l + 1]) ++l;
  const int pix = a - lv.a_off[l];
  const int W = lv.W[l];
  const float s = lv.stride[l];
  const float ax = static_cast<float>(pix % W) + 0.5f, ay = static_cast<float>(pix / W) + 0.5f;
  const float* r = lv.reg[l] + (static_cast<size_t>(b) * lv.hw[l] + pix) * lv.reg_cstride;
  float d[4];
#pragma unroll
  for (int side = 0; side < 4; ++side) {
    float v[17];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 17; ++k) { v[k] = r[side * 17 + k]; mx = fmaxf(mx, v[k]); }
    float sum = 0.f, ex = 0.f;
#pragma unroll
    for (int k = 0; k < 17; ++k) {
      const float e = expf(v[k] - mx);
      sum += e;
      ex = fmaf(e, static_cast<float>(k), ex);
    }
    d[side] = ex / sum;
  }
  const float logit = r[68];
  float4 o = make_float4((ax - d[0]) * s, (ay - d[1]) * s, (ax + d[2]) * s, (ay + d[3]) * s);
  reinterpret_cast<float4*>(boxes)[idx] = o;
  scores[idx] = 1.f / (1.f + expf(-logit));
}

int box_decode_launch(const DecodeLevels& lv, float* boxes, float* scores, int B, int A, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * A;
  box_decode_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(lv, boxes, scores, B, A);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- FLAME rows
// One 413-float row as the traced reference model emits it (yolo_head_dfl_head.py:162-184 then
// yolo_head_ndfl_heads.py:167-172, including the from_3dmm/to_3dmm channel rotation of 400..408).
// raw row layout (208 floats): [shape128 | expr64 | rot6 | jaw3 | transl3 | scale1 | pad3].
__device__ __forceinline__ float flame_row_value(const float* __restrict__ raw, int i, float ax_s, float ay_s,
                                                 float stride) {
  if (i < 128) return 3.f * tanhf(raw[i]);
  if (i < 300) return 0.f;
  if (i < 364) return 3.f * tanhf(raw[128 + (i - 300)]);
  if (i < 400) return 0.f;
  if (i < 406) {  // o[400+k] = c[403+k]: rot[3..5], jaw[0..2]
    const int k = i - 400;
    return k < 3 ? raw[192 + 3 + k] : raw[198 + (k - 3)];
  }
  if (i < 409) return raw[192 + (i - 406)];  // o[406+k] = c[400+k] = rot[0..2]
  if (i == 409) return raw[201] + ax_s;
  if (i == 410) return raw[202] + ay_s;
  if (i == 411) return raw[203];
  return __fmul_rn(__fdiv_rn(expf(raw[204]), 0.05f), stride);
}

__device__ __forceinline__ const float* raw_row(const DecodeLevels& lv, int b, int a, float& ax_s, float& ay_s,
                                                float& stride) {
  int l = 0;
  while (l < 2 && a >= lv.a_off[l + 1]) ++l;
  const int pix = a - lv.a_off[l];
  const int W = lv.W[l];
  stride = lv.stride[l];
  ax_s = (static_cast<float>(pix % W) + 0.5f) * stride;
  ay_s = (static_cast<float>(pix / W) + 0.5f) * stride;
  return lv.flame[l] + (static_cast<size_t>(b) * lv.hw[l] + pix) * lv.flame_cstride;
}

__global__ void __launch_bounds__(128) flame_dense_kernel(const DecodeLevels lv, float* __restrict__ out, int B, int A) {
  const int a = blockIdx.x, b = blockIdx.y;
  float ax_s, ay_s, stride;
  const float* raw = raw_row(lv, b, a, ax_s, ay_s, stride);
  float* o = out + (static_cast<size_t>(b) * A + a) * 413;
  for (int i = threadIdx.x; i < 413; i += blockDim.x) o[i] = flame_row_value(raw, i, ax_s, ay_s, stride);
}

int flame_dense_launch(const DecodeLevels& lv, float* out, int B, int A, cudaStream_t stream) {
  flame_dense_kernel<<<dim3(A, B), 128, 0, stream>>>(lv, out, B, A);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// survivors only: exclusive scan of per-image counts, then packed rows (image-major, score order)
__global__ void __launch_bounds__(1024) head_offsets_kernel(const int* __restrict__ cnt, int B, int* __restrict__ offsets,
                                                            int* __restrict__ total) {
  __shared__ int s[1024];
  int run = 0;
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < B ? cnt[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) offsets[i] = run + s[threadIdx.x] - v;
    const int chunk_total = s[1023];
    __syncthreads();
    run += chunk_total;
  }
  if (threadIdx.x == 0) { offsets[B] = run; *total = run; }
}

__global__ void __launch_bounds__(128) flame_gather_kernel(const DecodeLevels lv, const int* __restrict__ keep_idx,
                                                           const int* __restrict__ keep_cnt,
                                                           const int* __restrict__ offsets, int keep_k,
                                                           const float* __restrict__ img_xform,
                                                           float* __restrict__ params, float* __restrict__ head_xform,
                                                           int* __restrict__ head_img) {
  const int j = blockIdx.x, b = blockIdx.y;
  if (j >= keep_cnt[b]) return;
  const int a = keep_idx[b * keep_k + j];
  const int dst = offsets[b] + j;
  float ax_s, ay_s, stride;
  const float* raw = raw_row(lv, b, a, ax_s, ay_s, stride);
  float* o = params + static_cast<size_t>(dst) * 413;
  for (int i = threadIdx.x; i < 413; i += blockDim.x) o[i] = flame_row_value(raw, i, ax_s, ay_s, stride);
  if (threadIdx.x < 3) head_xform[dst * 3 + threadIdx.x] = img_xform ? img_xform[b * 3 + threadIdx.x] : (threadIdx.x == 2 ? 1.f : 0.f);
  if (threadIdx.x == 0 && head_img) head_img[dst] = b;
}

int flame_gather_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, int B, int keep_k,
                        const float* img_xform, int* offsets, int* total, float* params, float* head_xform,
                        int* head_img, cudaStream_t stream) {
  head_offsets_kernel<<<1, 1024, 0, stream>>>(keep_cnt, B, offsets, total);
  flame_gather_kernel<<<dim3(keep_k, B), 128, 0, stream>>>(lv, keep_idx, keep_cnt, offsets, keep_k, img_xform, params,
                                                           head_xform, head_img);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
