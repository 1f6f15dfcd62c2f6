#pragma once
// Multi-GPU prediction gather: one packed result RECORD per step, written by ONE kernel straight into
// its destination - local memory (snapshot for an NCCL send) or rank 0's peer-mapped receive ring over NVLink.
#include <cuda_runtime.h>
#include <cstdint>

namespace vgh {

// Record layout in 4-byte words (B images, K = keep_top_k, n = heads of this step, known on the device only):
//   [0] n  [1] B  [2] K  [3] sequence number  [4..15] reserved
//   [16, 16+B)            keep_cnt   int32
//   [.., +4*B*K)          keep_boxes float [B,K,4]
//   [.., +B*K)            keep_scores float [B,K]          -> fixed part, `fixed_words` (multiple of 4)
//   [fixed, +pad4(413 n)) params     float [n,413]
//   [.., +15069 n)        vertices   float [n,5023,3]
struct RecordLayout {
  int B, K;
  long fixed_words, capacity_words;
};
constexpr int kRecordHeader = 16;
inline long record_pad4(long w) { return (w + 3) & ~3L; }
RecordLayout record_layout(int B, int K);

struct RecordSrc {
  const int* keep_cnt;
  const int* total;  // device: heads of this step
  const float *keep_boxes, *keep_scores, *params, *verts;
};

// (1) if wait_flag: a one-warp kernel spins until *wait_flag >= wait_val (the consumer handed the slot back), bounded by
//     timeout_ms (then *status |= 1 and the step proceeds);
// (2) pack kernel copies the record to dst (any device-accessible address, local or peer);
// (3) if done_flag: the last block to finish publishes *done_flag = done_val (system-scope release) after every
//     block's stores are visible system-wide.
// `block_counter` is a zero-initialised device int owned by the caller (reset by the kernel).
int record_push_launch(const RecordSrc& src, const RecordLayout& lay, uint32_t seq, float* dst, const unsigned long long* wait_flag,
                       unsigned long long wait_val, unsigned long long* done_flag, unsigned long long done_val,
                       int* block_counter, int* status, int timeout_ms, cudaStream_t stream);

// Consumer side (rank 0): one block waits until ready[i * stride] >= value for all i < n (bounded by timeout_ms ->
// *status |= 2), sums the head counts of the n records (record i at records + i * record_stride words) into
// *total_out, then writes ack_val to every ack_ptrs[i] (peer addresses, system-scope release).
int gather_wait_launch(const unsigned long long* ready, int n, int stride, unsigned long long value, const float* records,
                       long record_stride, unsigned long long* const* ack_ptrs_dev, unsigned long long ack_val, int* total_out,
                       int* status, int timeout_ms, cudaStream_t stream);

}  // namespace vgh
