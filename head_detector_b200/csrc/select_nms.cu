// Confidence select + top-k + NMS + gather, ONE CTA per image, one launch per batch (sm_100a).
//
// Replaces head_detector/utils.py:159-194 (mask -> topk -> torchvision.ops.nms -> [:keep_top_k]), for
// every image of the batch as yolo_heads_post_prediction_callback.py:55-97 does.  Integer/compare
// work; IoU arithmetic uses non-contracted fp32 intrinsics in torchvision's operation order so the
// kept anchor ids are bit-identical to the CPU reference on identical inputs.
//
//   1. candidates = {a : score[a] >= conf_thr}; key = ordered(score_bits)<<32 | ~a  (descending key order ==
//      descending score for any sign, ties -> lower anchor id first)
//   2. if more than top_k candidates: 8-bit MSB radix select of the top_k-th key
//   3. compaction into shared memory, bitonic sort (descending)
//   4. n x n suppression bit matrix in shared memory (n <= 1024)
//   5. one warp walks the matrix serially (greedy), stops at keep_k survivors
//   6. kept anchor ids / boxes / scores written out
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "select_nms.cuh"
#include "device_attr.cuh"

namespace vgh {

constexpr int kNmsThreads = 1024;
constexpr int kMaxCand = 1024;

struct NmsArgs {
  const float* boxes;   // [B,A,4] xyxy
  const float* scores;  // [B,A]
  int A;
  float conf_thr, iou_thr;
  int top_k, keep_k;
  int* keep_idx;      // [B,keep_k] anchor ids (-1 padded)
  int* keep_cnt;      // [B]
  float* keep_boxes;  // optional [B,keep_k,4]
  float* keep_scores; // optional [B,keep_k]
};

// order-preserving float -> uint map (negative scores sort below positive ones; -0.0 < +0.0 only in the tie-break)
__device__ __forceinline__ unsigned long long make_key(float s, int a) {
  const unsigned u = __float_as_uint(s);
  const unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(ord) << 32) | static_cast<unsigned>(~static_cast<unsigned>(a));
}

__global__ void __launch_bounds__(kNmsThreads) select_nms_kernel(const NmsArgs p) {
  extern __shared__ __align__(16) uint8_t smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);      // [1024]
  float4* sbox = reinterpret_cast<float4*>(keys + kMaxCand);                    // [1024]
  float* sarea = reinterpret_cast<float*>(sbox + kMaxCand);                     // [1024]
  unsigned* mask = reinterpret_cast<unsigned*>(sarea + kMaxCand);              // [1024][32]
  __shared__ unsigned hist[256];
  __shared__ int s_count, s_n;
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* sc = p.scores + static_cast<size_t>(b) * p.A;
  const float4* bx = reinterpret_cast<const float4*>(p.boxes) + static_cast<size_t>(b) * p.A;

  // ---- 1. count candidates
  if (tid == 0) { s_count = 0; s_n = 0; }
  __syncthreads();
  int local = 0;
  for (int a = tid; a < p.A; a += kNmsThreads) local += (sc[a] >= p.conf_thr) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((tid & 31) == 0 && local) atomicAdd(&s_count, local);
  __syncthreads();
  const int total = s_count;

  // ---- 2. threshold key (radix select) when there are more than top_k candidates
  unsigned long long thr_key = 0ull;  // keep keys >= thr_key
  if (total > p.top_k) {
    if (tid == 0) { s_prefix = 0ull; s_need = p.top_k; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kNmsThreads) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
      for (int a = tid; a < p.A; a += kNmsThreads) {
        const float s = sc[a];
        if (s >= p.conf_thr) {
          const unsigned long long k = make_key(s, a);
          if ((k & hi_mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int need = s_need;
        int d = 255;
        for (; d > 0; --d) {
          const int c = static_cast<int>(hist[d]);
          if (c >= need) break;
          need -= c;
        }
        s_need = need;
        s_prefix = prefix | (static_cast<unsigned long long>(d) << shift);
      }
      __syncthreads();
    }
    thr_key = s_prefix;  // exactly top_k keys are >= this one (keys are unique)
  }

  // ---- 3. compaction + sort
  for (int a = tid; a < p.A; a += kNmsThreads) {
    const float s = sc[a];
    if (s >= p.conf_thr) {
      const unsigned long long k = make_key(s, a);
      if (k >= thr_key) {
        const int pos = atomicAdd(&s_n, 1);
        if (pos < kMaxCand) keys[pos] = k;
      }
    }
  }
  __syncthreads();
  const int n = min(s_n, kMaxCand);
  int n2 = 32;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += kNmsThreads) keys[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += kNmsThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys[i], y = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }

  // ---- 4. boxes, areas, suppression matrix
  const int nw = (n + 31) >> 5;
  for (int i = tid; i < n; i += kNmsThreads) {
    const int a = static_cast<int>(~static_cast<unsigned>(keys[i] & 0xffffffffull));
    const float4 q = bx[a];
    sbox[i] = q;
    sarea[i] = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
  }
  __syncthreads();
  for (int item = tid; item < n * nw; item += kNmsThreads) {
    const int i = item / nw, w = item - i * nw;
    unsigned bits = 0u;
    const int j0 = w << 5;
    if (j0 + 31 > i) {
      const float4 bi = sbox[i];
      const float ai = sarea[i];
      const int jend = min(32, n - j0);
      for (int t = 0; t < jend; ++t) {
        const int j = j0 + t;
        if (j <= i) continue;
        const float4 bj = sbox[j];
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float ww = fmaxf(0.f, __fsub_rn(xx2, xx1)), hh = fmaxf(0.f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(ww, hh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, sarea[j]), inter));
        if (ovr > p.iou_thr) bits |= (1u << t);
      }
    }
    mask[i * 32 + w] = bits;
  }
  __syncthreads();

  // ---- 5/6. greedy walk by warp 0; lane w owns word w of the "removed" bit vector
  if (tid < 32) {
    unsigned removed = 0u;
    int kept = 0;
    int* out_idx = p.keep_idx + static_cast<size_t>(b) * p.keep_k;
    for (int i = 0; i < n && kept < p.keep_k; ++i) {
      const unsigned word = __shfl_sync(0xffffffffu, removed, i >> 5);
      if ((word >> (i & 31)) & 1u) continue;
      if (tid < nw) removed |= mask[i * 32 + tid];
      if (tid == 0) {
        const unsigned long long k = keys[i];
        const int a = static_cast<int>(~static_cast<unsigned>(k & 0xffffffffull));
        out_idx[kept] = a;
        if (p.keep_boxes) reinterpret_cast<float4*>(p.keep_boxes)[static_cast<size_t>(b) * p.keep_k + kept] = sbox[i];
        if (p.keep_scores) p.keep_scores[static_cast<size_t>(b) * p.keep_k + kept] = sc[a];
      }
      ++kept;
    }
    for (int i = kept + tid; i < p.keep_k; i += 32) out_idx[i] = -1;
    if (tid == 0) p.keep_cnt[b] = kept;
  }
}

int select_nms_launch(const float* boxes, const float* scores, int B, int A, float conf_thr, float iou_thr, int top_k,
                      int keep_k, int* keep_idx, int* keep_cnt, float* keep_boxes, float* keep_scores,
                      cudaStream_t stream, char* err, size_t errlen) {
  if (B <= 0) return 0;
  if (top_k < 1 || top_k > kMaxCand || keep_k < 1) {
    snprintf(err, errlen, "select_nms: top_k must be in [1,%d], keep_k >= 1", kMaxCand);
    return 1;
  }
  NmsArgs p{boxes, scores, A, conf_thr, iou_thr, top_k, keep_k, keep_idx, keep_cnt, keep_boxes, keep_scores};
  const size_t smem = kMaxCand * (8 + 16 + 4 + 32 * 4);
  static SmemOptIn opt_in;
  {
    cudaError_t e = ensure_dynamic_smem(select_nms_kernel, opt_in, smem);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "select_nms smem: %s", cudaGetErrorString(e));
      return 2;
    }
  }
  select_nms_kernel<<<B, kNmsThreads, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "select_nms launch: %s", cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

}  // namespace vgh
