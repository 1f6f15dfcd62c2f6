// Small CUDA-core kernels around the tensor-core convs (sm_100a): stem im2col of the uint8 input, SPP
// max-pools, DFL/sigmoid box decode, FLAME-row assembly (dense or for NMS survivors only).
// All HBM-bound elementwise/stencil work: coalesced 16-byte accesses, no reshaping into GEMMs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "aux_kernels.cuh"
#include "device_attr.cuh"

namespace vgh {

// ---------------------------------------------------------------------------------------- stem
// The 3x3 stride-2 stem (3 -> 48 channels) has K = 27: too small for a TMA/UMMA K block fed from the
// uint8 image directly.  This kernel expands each output pixel's 27 input bytes (ky,kx,c order, zero
// outside the image = conv padding) to a 32-wide bf16 row (5 zero columns); the stem then runs on
// the tensor cores as a 1x1 conv with Cin = 32 whose weights carry the /255 of detector.py:51.
// uint8 -> bf16 is exact.
__global__ void __launch_bounds__(256) stem_pack_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                        int B, int S, int split) {
  // thread = one output pixel: its 3x3x3 window is three runs of 9 consecutive bytes (one per image row)
  const int Ho = S >> 1;
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * Ho * Ho;
  if (pix >= total) return;
  const int ow = static_cast<int>(pix % Ho);
  const int oh = static_cast<int>((pix / Ho) % Ho);
  const int b = static_cast<int>(pix / (static_cast<long long>(Ho) * Ho));
  float x[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) x[i] = 0.f;
  const int iw0 = 2 * ow - 1;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ih = 2 * oh + ky - 1;
    const bool row_ok = ih >= 0 && ih < S;
    const uint8_t* p = img + ((static_cast<size_t>(b) * S + (row_ok ? ih : 0)) * S + iw0) * 3;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const bool ok = row_ok && (j >= 3 || iw0 >= 0) && (iw0 + j / 3 < S);
      x[ky * 9 + j] = ok ? static_cast<float>(__ldg(p + j)) : 0.f;
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + pix * (split ? 192 : 32));
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(x[u * 8 + 2 * i], x[u * 8 + 2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
    op[u] = v;
    if (split) {  // planes [h|m|h|m|h|l] of 32 channels (4 uint4) each; m and l stay zero (the buffer is zero-initialised)
      op[8 + u] = v;
      op[16 + u] = v;
    }
  }
}

int stem_pack_launch(const uint8_t* img, __nv_bfloat16* out, int B, int S, int split, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * (S / 2) * (S / 2);
  stem_pack_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(img, out, B, S, split);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- SPP
// max-pool k=5,9,13 stride 1 (pad k/2 with -inf) of channel slice [0,C) of a [B,H,W,4C] buffer into
// slices 1,2,3.
template <bool F16>
__device__ __forceinline__ void max8(uint4& m, const uint4 v) {
  if constexpr (F16) {
    __half2* a = reinterpret_cast<__half2*>(&m);
    const __half2* b = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], b[i]);
  } else {
    __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&m);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hmax2(a[i], b[i]);
  }
}

// CTA = one image x CH channels; the H x W x CH slab lives in shared memory and the three pools are
// computed as cascaded separable 5-wide maxima (mp9 = mp5(mp5), mp13 = mp5(mp9); -inf padding).
template <bool kRows, bool F16>
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
      max8<F16>(m, src[(yy * W + xx) * L + l]);
    }
    dst[i] = m;
  }
}

template <bool F16>
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
    max5_pass<true, F16>(cur, b1, H, W, L);
    __syncthreads();
    max5_pass<false, F16>(b1, out, H, W, L);
    __syncthreads();
    for (int i = threadIdx.x; i < H * W * L; i += blockDim.x)
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(i / L) * CT + k * C + (i % L) * 8) = out[i];
    uint4* t = cur;
    cur = out;
    out = t;
  }
}

int spp_pool_launch(__nv_bfloat16* buf, int B, int H, int W, int C, int f16, cudaStream_t stream) {
  int CH = 32;
  while (CH > 8 && static_cast<size_t>(H) * W * CH * 2 * 3 > 160 * 1024) CH >>= 1;
  if (C % CH) return 1;
  const size_t smem = static_cast<size_t>(H) * W * CH * 2 * 3;
  static SmemOptIn opt_in[2];
  if (f16) {
    if (ensure_dynamic_smem(spp_pool_kernel<true>, opt_in[1], smem) != cudaSuccess) return 1;
    spp_pool_kernel<true><<<B * (C / CH), 256, smem, stream>>>(buf, B, H, W, C, CH);
  } else {
    if (ensure_dynamic_smem(spp_pool_kernel<false>, opt_in[0], smem) != cudaSuccess) return 1;
    spp_pool_kernel<false><<<B * (C / CH), 256, smem, stream>>>(buf, B, H, W, C, CH);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Parity mode: activations are split bf16 terms, six planes [h|m|h|m|h|l] per 32-channel granule.  One thread per
// (image, pixel, logical channel): y = (l + m) + h is exact in fp32, the three pools are plain window maxima over y
// (small maps, parity runs only - no attempt at speed), results are re-split and written to slices 1..3.
__device__ __forceinline__ float split_load(const __nv_bfloat16* g) {
  return (__bfloat162float(g[160]) + __bfloat162float(g[32])) + __bfloat162float(g[0]);
}
__device__ __forceinline__ void split_store(__nv_bfloat16* g, float y) {
  const __nv_bfloat16 h = __float2bfloat16_rn(y);
  y -= __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(y);
  y -= __bfloat162float(m);
  g[0] = h; g[64] = h; g[128] = h;
  g[32] = m; g[96] = m;
  g[160] = __float2bfloat16_rn(y);
}
__global__ void __launch_bounds__(256) spp_pool_split_kernel(__nv_bfloat16* __restrict__ buf, int B, int H, int W, int C) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * H * W * C) return;
  const int c = static_cast<int>(idx % C);
  const int x = static_cast<int>((idx / C) % W), y = static_cast<int>((idx / C / W) % H), b = static_cast<int>(idx / C / W / H);
  const size_t CT = static_cast<size_t>(4 * C) * 6;
  const __nv_bfloat16* img = buf + static_cast<size_t>(b) * H * W * CT;
  const int goff = 192 * (c >> 5) + (c & 31);
  float m5 = -INFINITY, m9 = -INFINITY, m13 = -INFINITY;
  for (int dy = -6; dy <= 6; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -6; dx <= 6; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const float v = split_load(img + (static_cast<size_t>(yy) * W + xx) * CT + goff);
      m13 = fmaxf(m13, v);
      if (abs(dy) <= 4 && abs(dx) <= 4) m9 = fmaxf(m9, v);
      if (abs(dy) <= 2 && abs(dx) <= 2) m5 = fmaxf(m5, v);
    }
  }
  __nv_bfloat16* o = buf + (static_cast<size_t>(b) * H * W + static_cast<size_t>(y) * W + x) * CT + goff;
  split_store(o + 6 * C, m5);
  split_store(o + 12 * C, m9);
  split_store(o + 18 * C, m13);
}
int spp_pool_split_launch(__nv_bfloat16* buf, int B, int H, int W, int C, cudaStream_t stream) {
  if (C % 32) return 1;
  const long long total = static_cast<long long>(B) * H * W * C;
  spp_pool_split_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(buf, B, H, W, C);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- box decode
// yolo_head_ndfl_heads.py:143-144,164-165: DFL softmax expectation, sigmoid, distance2bbox * stride.
// CTA = 128 consecutive anchors of one image and level: the 128 x reg_cstride raw rows are staged in
// shared memory with coalesced 16-byte loads (row pitch +1 float: conflict-free per-thread reads).
__global__ void __launch_bounds__(128) box_decode_kernel(const DecodeLevels lv, float* __restrict__ boxes,
                                                         float* __restrict__ scores, int B, int A, int blocks_l0,
                                                         int blocks_l1, int blocks_per_img) {
  extern __shared__ float rows[];  // [128][cs + 1]
  const int b = blockIdx.x / blocks_per_img;
  int blk = blockIdx.x - b * blocks_per_img;
  int l = 0;
  if (blk >= blocks_l0) { blk -= blocks_l0; l = 1; if (blk >= blocks_l1) { blk -= blocks_l1; l = 2; } }
  const int cs = lv.reg_cstride, pitch = cs + 1;
  const int pix0 = blk * 128;
  const int n_here = min(128, lv.hw[l] - pix0);
  const float* src = lv.reg[l] + (static_cast<size_t>(b) * lv.hw[l] + pix0) * cs;
  const int vec_per_row = cs >> 2;
  for (int i = threadIdx.x; i < n_here * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c4 = i - r * vec_per_row;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    float* dst = rows + r * pitch + c4 * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  if (threadIdx.x >= n_here) return;
  const int pix = pix0 + threadIdx.x;
  const int a = lv.a_off[l] + pix;
  const int W = lv.W[l];
  const float s = lv.stride[l];
  const float ax = static_cast<float>(pix % W) + 0.5f, ay = static_cast<float>(pix / W) + 0.5f;
  const float* r = rows + threadIdx.x * pitch;
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
  const size_t idx = static_cast<size_t>(b) * A + a;
  reinterpret_cast<float4*>(boxes)[idx] = make_float4((ax - d[0]) * s, (ay - d[1]) * s, (ax + d[2]) * s, (ay + d[3]) * s);
  scores[idx] = 1.f / (1.f + expf(-logit));
}

int box_decode_launch(const DecodeLevels& lv, float* boxes, float* scores, int B, int A, cudaStream_t stream) {
  const int b0 = (lv.hw[0] + 127) / 128, b1 = (lv.hw[1] + 127) / 128, b2 = (lv.hw[2] + 127) / 128;
  const int per_img = b0 + b1 + b2;
  const size_t smem = 128 * (lv.reg_cstride + 1) * sizeof(float);
  box_decode_kernel<<<B * per_img, 128, smem, stream>>>(lv, boxes, scores, B, A, b0, b1, per_img);
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

// head = packed index of the survivor (sparse heads only; ignored for dense maps)
__device__ __forceinline__ const float* raw_row(const DecodeLevels& lv, int b, int a, float& ax_s, float& ay_s,
                                                float& stride, int head = -1) {
  int l = 0;
  while (l < 2 && a >= lv.a_off[l + 1]) ++l;
  const int pix = a - lv.a_off[l];
  const int W = lv.W[l];
  stride = lv.stride[l];
  ax_s = (static_cast<float>(pix % W) + 0.5f) * stride;
  ay_s = (static_cast<float>(pix / W) + 0.5f) * stride;
  if (lv.head_patch != nullptr && head >= 0) {  // centre pixel of the survivor's patch
    const size_t row = (static_cast<size_t>(lv.head_patch[head]) * kPatch + kPatchC) * kPatch + kPatchC;
    return lv.flame[l] + row * lv.flame_cstride;
  }
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
  const float* raw = raw_row(lv, b, a, ax_s, ay_s, stride, dst);
  float* o = params + static_cast<size_t>(dst) * 413;
  for (int i = threadIdx.x; i < 413; i += blockDim.x) o[i] = flame_row_value(raw, i, ax_s, ay_s, stride);
  if (threadIdx.x < 3) head_xform[dst * 3 + threadIdx.x] = img_xform ? img_xform[b * 3 + threadIdx.x] : (threadIdx.x == 2 ? 1.f : 0.f);
  if (threadIdx.x == 0 && head_img) head_img[dst] = b;
}

int head_offsets_launch(const int* keep_cnt, int B, int* offsets, int* total, cudaStream_t stream) {
  head_offsets_kernel<<<1, 1024, 0, stream>>>(keep_cnt, B, offsets, total);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int flame_gather_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, int B, int keep_k,
                        const float* img_xform, const int* offsets, float* params, float* head_xform,
                        int* head_img, cudaStream_t stream) {
  flame_gather_kernel<<<dim3(keep_k, B), 128, 0, stream>>>(lv, keep_idx, keep_cnt, offsets, keep_k, img_xform, params,
                                                           head_xform, head_img);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- sparse heads
// Patch numbers per level in image-major, slot-major order (deterministic): pass 1 counts the survivors of every
// (image, level) - one warp per image -, an exclusive scan over the images gives each image's first patch number
// per level, pass 2 hands the numbers out with ballot prefix counts.
__global__ void __launch_bounds__(1024) patch_assign_kernel(const DecodeLevels lv, const int* __restrict__ keep_idx,
                                                            const int* __restrict__ keep_cnt, const int* __restrict__ offsets, int B,
                                                            int keep_k, int cap, int* __restrict__ head_level,
                                                            int* __restrict__ head_patch, int* __restrict__ patch_src,
                                                            int* __restrict__ level_rows) {
  extern __shared__ int base_s[];  // [3][B] counts, then exclusive prefixes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  auto level_of = [&](int a, int& y, int& x) {
    int l = 0;
    while (l < 2 && a >= lv.a_off[l + 1]) ++l;
    const int pix = a - lv.a_off[l];
    y = pix / lv.W[l];
    x = pix - y * lv.W[l];
    return l;
  };
  for (int b = warp; b < B; b += n_warps) {
    const int n = min(keep_cnt[b], keep_k);
    int c[3] = {0, 0, 0};
    for (int base = 0; base < n; base += 32) {
      const int j = base + lane;
      int l = -1, y, x;
      if (j < n) l = level_of(keep_idx[b * keep_k + j], y, x);
#pragma unroll
      for (int q = 0; q < 3; ++q) c[q] += __popc(__ballot_sync(0xffffffffu, l == q));
    }
    if (lane < 3) base_s[lane * B + b] = c[lane];
  }
  __syncthreads();
  if (threadIdx.x < 3) {  // exclusive scan over the images (B is a batch size: tens to hundreds)
    int run = 0;
    int* row = base_s + threadIdx.x * B;
    for (int b = 0; b < B; ++b) {
      const int v = row[b];
      row[b] = run;
      run += v;
    }
    level_rows[threadIdx.x] = min(run, cap) * kPatch;
  }
  __syncthreads();
  for (int b = warp; b < B; b += n_warps) {
    const int n = min(keep_cnt[b], keep_k);
    int run[3] = {base_s[b], base_s[B + b], base_s[2 * B + b]};
    for (int base = 0; base < n; base += 32) {
      const int j = base + lane;
      int l = -1, y = 0, x = 0;
      if (j < n) l = level_of(keep_idx[b * keep_k + j], y, x);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, l == q);
        if (l == q) {
          const int p = run[q] + __popc(m & ((1u << lane) - 1u));
          if (p < cap) {
            head_level[offsets[b] + j] = q;
            head_patch[offsets[b] + j] = p;
            patch_src[q * cap + p] = static_cast<int>((static_cast<unsigned>(b) << 20) | (static_cast<unsigned>(y) << 10) | static_cast<unsigned>(x));
          }
        }
        run[q] += __popc(m);
      }
    }
  }
}

int patch_assign_launch(const DecodeLevels& lv, const int* keep_idx, const int* keep_cnt, const int* offsets, int B, int keep_k,
                        int cap, int* head_level, int* head_patch, int* patch_src, int* level_rows, cudaStream_t stream) {
  if (B > 4096) return 1;
  patch_assign_kernel<<<1, 1024, 3 * B * sizeof(int), stream>>>(lv, keep_idx, keep_cnt, offsets, B, keep_k, cap, head_level, head_patch,
                                                                patch_src, level_rows);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// CTA = one patch: kPatch x kPatch pixels x C channels (16-byte vectors) copied from the feature map, zeros outside it
__global__ void __launch_bounds__(256) patch_gather_kernel(const __nv_bfloat16* __restrict__ feat, int H, int W, int C_total, int coff,
                                                           int C, __nv_bfloat16* __restrict__ dst, int dst_C, int dst_coff,
                                                           const int* __restrict__ patch_src, const int* __restrict__ level_rows) {
  const int n_patches = *level_rows / kPatch;
  const int vec = C >> 3;  // uint4 = 8 bf16
  for (int p = blockIdx.x; p < n_patches; p += gridDim.x) {  // grid = a few CTAs per SM, not the patch capacity
    const int src = patch_src[p];
    const int b = static_cast<int>(static_cast<unsigned>(src) >> 20), y0 = ((src >> 10) & 1023) - kPatchC, x0 = (src & 1023) - kPatchC;
    for (int i = threadIdx.x; i < kPatch * kPatch * vec; i += blockDim.x) {
      const int pix = i / vec, v = i - pix * vec;
      const int r = pix / kPatch, c = pix - r * kPatch;
      const int yy = y0 + r, xx = x0 + c;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        val = __ldg(reinterpret_cast<const uint4*>(feat + ((static_cast<size_t>(b) * H + yy) * W + xx) * C_total + coff) + v);
      reinterpret_cast<uint4*>(dst + (static_cast<size_t>(p) * kPatch * kPatch + pix) * dst_C + dst_coff)[v] = val;
    }
  }
}

int patch_gather_launch(const __nv_bfloat16* feat, int H, int W, int C_total, int coff, int C, __nv_bfloat16* dst, int dst_C,
                        int dst_coff, const int* patch_src, const int* level_rows, int cap, cudaStream_t stream) {
  if (C % 8 || C_total % 8 || coff % 8 || dst_C % 8 || dst_coff % 8) return 1;
  patch_gather_kernel<<<cap < 592 ? cap : 592, 256, 0, stream>>>(feat, H, W, C_total, coff, C, dst, dst_C, dst_coff, patch_src, level_rows);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

__global__ void __launch_bounds__(128) patch_mask_kernel(__nv_bfloat16* __restrict__ buf, int C_total, int coff, int C, int H, int W,
                                                         const int* __restrict__ patch_src, const int* __restrict__ level_rows) {
  const int n_patches = *level_rows / kPatch;
  const int vec = C >> 3;
  for (int p = blockIdx.x; p < n_patches; p += gridDim.x) {
    const int src = patch_src[p];
    const int y0 = ((src >> 10) & 1023) - kPatchC, x0 = (src & 1023) - kPatchC;
    if (y0 >= 0 && x0 >= 0 && y0 + kPatch <= H && x0 + kPatch <= W) continue;  // interior patch: nothing lies outside
    for (int i = threadIdx.x; i < kPatch * kPatch * vec; i += blockDim.x) {
      const int pix = i / vec, v = i - pix * vec;
      const int r = pix / kPatch, c = pix - r * kPatch;
      const int yy = y0 + r, xx = x0 + c;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) continue;
      reinterpret_cast<uint4*>(buf + (static_cast<size_t>(p) * kPatch * kPatch + pix) * C_total + coff)[v] = make_uint4(0, 0, 0, 0);
    }
  }
}

int patch_mask_launch(__nv_bfloat16* buf, int C_total, int coff, int C, int H, int W, const int* patch_src,
                      const int* level_rows, int cap, cudaStream_t stream) {
  if (C % 8 || C_total % 8 || coff % 8) return 1;
  patch_mask_kernel<<<cap < 296 ? cap : 296, 128, 0, stream>>>(buf, C_total, coff, C, H, W, patch_src, level_rows);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------- result staging
// Copies the first *count_ptr rows (row_floats floats each) of src to dst; the count lives on the device.
__global__ void __launch_bounds__(256) copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        const int* __restrict__ count_ptr, int row_floats, int max_rows) {
  const int rows = min(*count_ptr, max_rows);
  const size_t n = static_cast<size_t>(rows) * row_floats;
  const size_t n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    d4[i] = s4[i];
  for (size_t i = (n4 << 2) + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

int copy_rows_launch(const float* src, float* dst, const int* count_ptr, int row_floats, int max_rows, cudaStream_t stream) {
  copy_rows_kernel<<<592, 256, 0, stream>>>(src, dst, count_ptr, row_floats, max_rows);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
