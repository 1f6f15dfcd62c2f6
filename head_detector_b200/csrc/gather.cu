// Multi-GPU prediction gather kernels (see gather.cuh).  The exchange of SURVEY.md 8e - every rank's predictions of a
// step end up on rank 0 - is fused into the result snapshot every step needs anyway: the pack kernel reads the live
// result buffers once and writes the packed record straight into rank 0's receive ring through the peer mapping
// (NVLink / NVSwitch stores), so there is no staging copy, no size exchange and no host synchronisation on the path.
#include "gather.cuh"

#include "../../include/vggheads_b200.h"

namespace vgh {

RecordLayout record_layout(int B, int K) {
  RecordLayout l;
  l.B = B;
  l.K = K;
  l.fixed_words = kRecordHeader + record_pad4(B) + 4L * B * K + record_pad4(static_cast<long>(B) * K);
  const long cap = static_cast<long>(B) * K;
  l.capacity_words = l.fixed_words + record_pad4(cap * VGH_NUM_PARAMS) + cap * VGH_NUM_VERTS * 3;
  return l;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one warp; lane 0 polls.  Bounded: a consumer that died must not hang the producer's GPU.
__global__ void wait_flag_kernel(const unsigned long long* flag, unsigned long long value, int* status, unsigned long long timeout_ns) {
  if (threadIdx.x != 0) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(200);
    if (global_ns() - t0 > timeout_ns) {
      if (status) atomicOr(status, 1);
      return;
    }
  }
}

__device__ __forceinline__ void copy_words(float* __restrict__ dst, const float* __restrict__ src, long n, long tid, long nthreads) {
  const long n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (long i = tid; i < n4; i += nthreads) d4[i] = s4[i];
  for (long i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) record_pack_kernel(RecordSrc src, int B, int K, long fixed_words, uint32_t seq, float* __restrict__ dst,
                                                          unsigned long long* done_flag, unsigned long long done_val, int* block_counter) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const long cap = static_cast<long>(B) * K;
  long n = *src.total;
  n = n < 0 ? 0 : (n > cap ? cap : n);
  if (tid < kRecordHeader) {
    int* h = reinterpret_cast<int*>(dst);
    h[tid] = tid == 0 ? static_cast<int>(n) : tid == 1 ? B : tid == 2 ? K : tid == 3 ? static_cast<int>(seq) : 0;
  }
  long at = kRecordHeader;
  copy_words(dst + at, reinterpret_cast<const float*>(src.keep_cnt), B, tid, nthreads);
  at += (B + 3) & ~3L;
  copy_words(dst + at, src.keep_boxes, 4 * cap, tid, nthreads);
  at += 4 * cap;
  copy_words(dst + at, src.keep_scores, cap, tid, nthreads);
  copy_words(dst + fixed_words, src.params, n * VGH_NUM_PARAMS, tid, nthreads);
  copy_words(dst + fixed_words + ((n * VGH_NUM_PARAMS + 3) & ~3L), src.verts, n * VGH_NUM_VERTS * 3, tid, nthreads);
  if (!done_flag) return;
  // publish: every block makes its stores visible system-wide, the last one to arrive raises the flag
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(block_counter, 1);
    if (arrived == static_cast<int>(gridDim.x) - 1) {
      *block_counter = 0;
      __threadfence_system();
      st_release_sys(done_flag, done_val);
    }
  }
}

int record_push_launch(const RecordSrc& src, const RecordLayout& lay, uint32_t seq, float* dst, const unsigned long long* wait_flag,
                       unsigned long long wait_val, unsigned long long* done_flag, unsigned long long done_val, int* block_counter,
                       int* status, int timeout_ms, cudaStream_t stream) {
  if (wait_flag && wait_val > 0)
    wait_flag_kernel<<<1, 32, 0, stream>>>(wait_flag, wait_val, status, static_cast<unsigned long long>(timeout_ms) * 1000000ULL);
  record_pack_kernel<<<592, 256, 0, stream>>>(src, lay.B, lay.K, lay.fixed_words, seq, dst, done_flag, done_val, block_counter);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

__global__ void gather_wait_kernel(const unsigned long long* ready, int n, int stride, unsigned long long value, const float* records,
                                   long record_stride, unsigned long long* const* ack_ptrs, unsigned long long ack_val, int* total_out,
                                   int* status, unsigned long long timeout_ns) {
  __shared__ int s_total;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long t0 = global_ns();
    bool ok = true;
    while (ld_acquire_sys(ready + static_cast<long>(i) * stride) < value) {
      __nanosleep(200);
      if (global_ns() - t0 > timeout_ns) { ok = false; break; }
    }
    if (!ok) { if (status) atomicOr(status, 2); }
    else if (records) atomicAdd(&s_total, reinterpret_cast<const int*>(records + static_cast<long>(i) * record_stride)[0]);
  }
  __syncthreads();
  if (threadIdx.x == 0 && total_out) *total_out = s_total;
  if (ack_ptrs)
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (ack_ptrs[i]) st_release_sys(ack_ptrs[i], ack_val);
}

int gather_wait_launch(const unsigned long long* ready, int n, int stride, unsigned long long value, const float* records,
                       long record_stride, unsigned long long* const* ack_ptrs_dev, unsigned long long ack_val, int* total_out,
                       int* status, int timeout_ms, cudaStream_t stream) {
  gather_wait_kernel<<<1, 32, 0, stream>>>(ready, n, stride, value, records, record_stride, ack_ptrs_dev, ack_val, total_out, status,
                                           static_cast<unsigned long long>(timeout_ms) * 1000000ULL);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace vgh
